"""BF16 mode, kernel by kernel: errors of every bf16 kernel against a float64 evaluation of the same (bf16-rounded) inputs,
printed without asserting -- one run on the GPU box tells which kernel / shape is off and by how much.
    python tools/diag_bf16.py [small|full]
"""
import os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))

from deepsphere_tf1_b200 import ops  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

dev = torch.device('cuda', 0)
BF = torch.bfloat16


def rel(a, e):
    a, e = a.double().cpu(), e.double().cpu()
    scale = max(float(e.abs().max()), 1e-30)
    d = (a - e).abs()
    return float(d.max()) / scale, float(d.pow(2).sum().sqrt() / max(float(e.pow(2).sum().sqrt()), 1e-30))


def report(name, a, e):
    try:
        m, l2 = rel(a, e)
        bad = int((~torch.isfinite(a.double())).sum())
        print('{:58s} max/max|e| {:.3e}   relL2 {:.3e}   nonfinite {}'.format(name, m, l2, bad), flush=True)
    except Exception:
        print(name, 'FAILED'); traceback.print_exc()


def bfr(t):
    return t.to(BF).to(torch.float32)


def gemm_case(R, Fin, K, Fout, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    x = bfr(torch.randn(K, R, Fin, generator=g))
    W = torch.randn(Fin * K, Fout, generator=g) * 0.2
    dy = bfr(torch.randn(R, Fout, generator=g))
    Wb = bfr(W).double().reshape(Fin, K, Fout)
    xd = x.double()
    y = torch.einsum('krf,fko->ro', xd, Wb)
    G = torch.einsum('ro,fko->krf', dy.double(), Wb)
    dW = torch.einsum('krf,ro->fko', xd, dy.double()).reshape(Fin * K, Fout)
    return x, W, dy, y, G, dW


def run_gemms(shapes):
    for (R, Fin, K, Fout) in shapes:
        tag = 'R={} Fin={} K={} Fout={}'.format(R, Fin, K, Fout)
        sup = [ops._lib_int('contract_bf16_supported', d, R, Fin, K, Fout) for d in range(3)]
        print('--', tag, 'native fwd/bwd_data/bwd_weight:', sup, flush=True)
        x, W, dy, y, G, dW = gemm_case(R, Fin, K, Fout)
        xb = x.to(dev).to(BF).contiguous()
        x0, xk = xb[0].view(1, R, Fin), (xb[1:].view(K - 1, 1, R, Fin) if K > 1 else None)
        Wd, dyb = W.to(dev), dy.to(dev).to(BF).view(1, R, Fout)
        try:
            sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
            yo = ops.contract_fwd(x0, xk, Wd, K, Fout, sums)
            torch.cuda.synchronize()
            report('  fwd  ' + tag, yo.view(R, Fout).float(), y)
            ys = yo.view(R, Fout).double()
            report('  fwd bn sums (vs stored y)', sums.cpu(), torch.cat([ys.sum(0), (ys * ys).sum(0)]).cpu())
        except Exception:
            traceback.print_exc()
        try:
            Go = ops.contract_bwd_data(dyb, Wd, K, Fin)
            torch.cuda.synchronize()
            report('  bwd_data ' + tag, Go.view(K, R, Fin).float(), G)
        except Exception:
            traceback.print_exc()
        try:
            dWo = torch.zeros(Fin * K, Fout, device=dev)
            ops.contract_bwd_weight(x0, xk, dyb, dWo, K)
            torch.cuda.synchronize()
            report('  bwd_weight ' + tag, dWo, dW)
        except Exception:
            traceback.print_exc()


def run_spmm(nside, order, Fs, B):
    from helpers import healpix_L
    L = healpix_L(nside, order)
    g = DeviceGraph(L, dev)
    M = L.shape[0]
    Ld = torch.from_numpy(L.toarray()).double() if M <= 4096 else None
    import scipy.sparse as sp
    Ls = sp.csr_matrix(L).astype(np.float64)
    for F in Fs:
        gen = torch.Generator().manual_seed(F)
        x, y, z = [bfr(torch.randn(B, M, F, generator=gen)) for _ in range(3)]

        def apply(Lm, t):
            return torch.from_numpy(np.stack([Lm @ t[b].double().numpy() for b in range(B)]))
        for dense in (True, False):
            ops.USE_DENSE_GROUPS = dense
            for (tr, nm) in ((False, 'L'), (True, 'L^T')):
                Lm = Ls.T.tocsr() if tr else Ls
                exp = 2.0 * apply(Lm, x) - 1.0 * y.double() + 0.5 * z.double()
                try:
                    pp = ops._pipe(g, 'bwd' if tr else 'fwd', B, M, F, esize=2, min_items=1)
                    out = ops.spmm(g, x.to(dev).to(BF), y.to(dev).to(BF), z.to(dev).to(BF), alpha=2.0, beta=-1.0, gamma=0.5, transpose=tr)
                    torch.cuda.synchronize()
                    report('  spmm bf16 nside={} M={} F={} {} dense={} plan={}'.format(nside, M, F, nm, dense, None if pp is None else (pp.TR, pp.HC)),
                           out.float(), exp)
                except Exception:
                    traceback.print_exc()
        ops.USE_DENSE_GROUPS = True
        K = 5
        try:
            basis = ops.cheb_basis(g, x.to(dev).to(BF), K)
            planes = [x.double()]
            planes.append(apply(Ls, planes[0]))
            for k in range(2, K):
                planes.append(2 * apply(Ls, planes[-1]) - planes[-2])
            report('  cheb_basis bf16 (rounding accumulates over K) F={}'.format(F), basis.float(), torch.stack(planes[1:]))
        except Exception:
            traceback.print_exc()
    del Ld


def run_elementwise():
    for (B, M, F, p) in ((3, 64, 16, 4), (2, 256, 32, 4), (2, 128, 64, 1), (2, 96, 6, 2)):
        gen = torch.Generator().manual_seed(F)
        y = bfr(torch.randn(B, M, F, generator=gen))
        mean = torch.randn(F, generator=gen) * 0.1
        rstd = torch.rand(F, generator=gen) + 0.5
        bias = torch.randn(F, generator=gen) * 0.1
        dout = bfr(torch.randn(B, M // p, F, generator=gen))
        args32 = (y.to(dev), mean.to(dev), rstd.to(dev), bias.to(dev))
        argsbf = (y.to(dev).to(BF), mean.to(dev), rstd.to(dev), bias.to(dev))
        for act in ('relu', 'elu'):
            try:
                o32 = ops.bn_act_pool_fwd(*args32, p, 'max', act)
                obf = ops.bn_act_pool_fwd(*argsbf, p, 'max', act)
                report('  bn_act_pool_fwd bf16 vs fp32 kernel F={} p={} {}'.format(F, p, act), obf.float(), o32)
                d32, dbf = dout.to(dev), dout.to(dev).to(BF)
                _, s32 = ops.bn_act_pool_bwd(*args32, d32, p, 'max', act, want_gz=False)
                _, sbf = ops.bn_act_pool_bwd(*argsbf, dbf, p, 'max', act, want_gz=False)
                report('  bn_act_pool_bwd sums', sbf, s32)
                g32 = ops.bn_act_pool_bwd_dy(*args32, d32, s32, float(B * M), p, 'max', act)
                gbf = ops.bn_act_pool_bwd_dy(*argsbf, dbf, s32, float(B * M), p, 'max', act)
                report('  bn_act_pool_bwd_dy', gbf.float(), g32)
                if act == 'relu' and F % 4 == 0:
                    l32 = ops.bn_relu_pool_sums(o32, d32, bias.to(dev))
                    lbf = ops.bn_relu_pool_sums(ops.to_bf16(o32), dbf, bias.to(dev))
                    report('  bn_relu_pool_sums', lbf, l32)
            except Exception:
                traceback.print_exc()
    t = torch.randn(1000003, device=dev)
    report('  cast round trip', ops.to_f32(ops.to_bf16(t)), t.to(BF).float())


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'small'
    run_elementwise()
    run_spmm(16, 0, (16, 32, 64), 3)
    run_spmm(32, 2, (16, 32), 2)
    shapes = [(1024, 16, 5, 32), (1000, 16, 5, 32), (4096, 32, 5, 64), (2048, 64, 5, 64), (512, 64, 5, 2), (2048, 128, 3, 64), (768, 16, 1, 16),
              (3072, 32, 4, 32), (4096, 16, 5, 64)]
    if mode == 'full':
        shapes += [(1048576, 16, 5, 32), (262144, 32, 5, 64), (65536, 64, 5, 64)]
    run_gemms(shapes)
