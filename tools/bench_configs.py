"""Bench lines for the BASELINE.json configurations other than the headline one (bench.py measures config 2):

  config 1  SHREC17-style classifier, HEALPix Nside 32 full sphere (12,288 px), 6 input features, K=4, F=[16,32,64,128,256],
            'CNN' head M=[55], batch 32   (experiments/shrec17/hyperparameters.py:114-135, 146)
  config 3  climate segmentation, icosahedral levels [5,5,4,3,2,1,0,0] (10,242 -> 12 vertices), 16 channels,
            F=[32,64,128,256,512,512,512], K=4, encoder-decoder with unpooling, batch 64   (experiments/climate/run_experiment_ico.py:61-86)
  config 5  GHCN-like station graph, 10,000 random stations, kNN k=10, combinatorial Laplacian passed un-rescaled,
            L=[L]*4, F=[50,100,100,1], K=5, 5 input features, dense regression, batch 128   (experiments/ghcn/GHCN_train.py:7-32)

  python tools/bench_configs.py [--configs 1,3,5] [--steps 20] [--out profiles/r02_bench_configs.jsonl]

One JSON line per configuration with the keys of bench.py: `value` (training step replayed as one CUDA graph, batch resident
in HBM), `e2e` (train_step(pinned host batch, next_batch=...), loss read back every step), `parity` (one training step against
the CPU oracle evaluated in float64 from identical weights, at the batch stated there), `cpu_baseline` (the float32 oracle on the
host cores, bounded sample), `roofline` (one recurrence step of the layer with the largest feature slab, timed alone back to
back over operand sets larger than L2, against the measured copy bandwidth; algorithmic bytes 3 X + 8 nnz (+ 4 (M + 1) for CSR)).
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import graphs, models, ops, optimizers, utils  # noqa: E402


def config(n):
    rs = np.random.RandomState(0)
    if n == 1:
        nsides = [32, 16, 8, 4, 2, 1]
        L, p = utils.build_laplacians(nsides, new=False)
        kw = dict(F=[16, 32, 64, 128, 256], K=[4] * 5, batch_norm=[True] * 5, M=[55], statistics='mean', num_feat_in=6)
        B = 32
        x = rs.randn(B, 12288, 6).astype(np.float32)
        labels = rs.randint(0, 55, size=B).astype(np.int32)
        opt = lambda lr: optimizers.AdamOptimizer(lr, epsilon=0.1)                    # noqa: E731
        return dict(name='SHREC17-style classifier, HEALPix Nside=32 full sphere (12,288 px), 6 features, K=4, '
                         'F=[16,32,64,128,256], M=[55]', L=L, p=p, kw=kw, sampling='healpix', B=B, x=x, labels=labels,
                    lr=5e-2, opt=opt, oracle_opt=('Adam', dict(epsilon=0.1)), cpu_batch=B)
    if n == 3:
        L, p = utils.build_laplacians([5, 5, 4, 3, 2, 1, 0, 0], sampling='icosahedron')
        kw = dict(F=[32, 64, 128, 256, 512, 512, 512], K=[4] * 7, batch_norm=[True] * 7, M=[], pool='average', dense=True,
                  Fseg=3, num_feat_in=16)
        B = 64
        x = rs.randn(B, 10242, 16).astype(np.float32)
        labels = rs.choice(3, size=(B, 10242), p=[0.977, 0.001, 0.022]).astype(np.int32)
        opt = lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8)                   # noqa: E731
        return dict(name='climate segmentation, icosahedral level-5 mesh (10,242 vertices), 16 channels, K=4, '
                         'F=[32,64,128,256,512,512,512], encoder-decoder, Fseg=3', L=L, p=p, kw=kw, sampling='icosahedron', B=B,
                    x=x, labels=labels, lr=1e-3, opt=opt, oracle_opt=('Adam', dict(epsilon=1e-8)), cpu_batch=8)
    if n == 5:
        # 10,000 stations: 20 von Mises-Fisher-like clusters + a uniform background
        N = 10000
        centres = rs.randn(20, 3)
        centres /= np.linalg.norm(centres, axis=1, keepdims=True)
        pts = np.concatenate([centres[rs.randint(0, 20, size=7000)] + 0.15 * rs.randn(7000, 3), rs.randn(3000, 3)])
        pts /= np.linalg.norm(pts, axis=1, keepdims=True)
        lon, lat = np.arctan2(pts[:, 1], pts[:, 0]), np.arcsin(pts[:, 2])
        Lc = graphs.sphere_knn_laplacian(lon, lat, 10)                                # combinatorial, NOT rescaled (GHCN_train.py:8)
        L = [Lc.tocoo()] * 4
        kw = dict(F=[50, 100, 100, 1], K=[5] * 4, batch_norm=[True] * 4, M=[], num_feat_in=5, regression=True, dense=True)
        B = 128
        x = rs.randn(B, N, 5).astype(np.float32)
        labels = rs.randn(B, N).astype(np.float32)
        opt = lambda lr: optimizers.RMSPropOptimizer(lr, decay=0.9, momentum=0.)      # noqa: E731
        return dict(name='GHCN-like station graph, 10,000 nodes, kNN k=10 (irregular degree, CSR), combinatorial Laplacian '
                         'un-rescaled, 5 features, K=5, F=[50,100,100,1], dense regression', L=L, p=[1] * 4, kw=kw,
                    sampling='healpix', B=B, x=x, labels=labels, lr=1e-3, opt=opt,
                    oracle_opt=('RMSProp', dict(decay=0.9, momentum=0.)), cpu_batch=16)
    raise ValueError(n)


def build_model(c, B, cuda_graph, dir_name):
    cls = models.cgcnn if c['sampling'] == 'healpix' else type('cgcnn_ico', (models.cgcnn,), {'sampling': c['sampling']})
    return cls(L=c['L'], p=c['p'], num_epochs=1, batch_size=B, eval_frequency=10 ** 9, seed=2, verbose=False,
               cuda_graph=cuda_graph, scheduler=lambda step: c['lr'], optimizer=c['opt'], dir_name=dir_name, **c['kw'])


def oracle_for(c, dtype):
    from oracle.model import OracleCgcnn
    return OracleCgcnn(c['L'], p=c['p'], sampling=c['sampling'], dtype=dtype, **c['kw'])


def parity(c, tol=1e-4):
    """One training step at a batch the float64 oracle finishes in seconds."""
    Bp = min(c['B'], c['cpu_batch'])
    m = build_model(c, Bp, False, '_cfg_parity')
    x, labels = c['x'][:Bp], c['labels'][:Bp]
    xd, ld = m._to_device(x), m._labels_to_device(labels)
    ops.fill(m._P.g, 0.0)
    m._loss.zero_()
    logits, _ = m._forward(xd, training=True, save=True)
    m._backward(m._loss_fwd_bwd(logits, ld, True))
    torch.cuda.synchronize()
    lt = torch.from_numpy(labels)
    res = {}
    for dt in (torch.float64, torch.float32):
        o = oracle_for(c, dt)
        for name, v in o.params.items():
            o.params[name] = m._P.views[name].detach().cpu().to(dt).reshape(v.shape).clone()
        res[dt] = o.loss_and_grads(torch.from_numpy(x).to(dt), lt if not c['kw'].get('regression') else lt.to(dt), True)
    total, ologits, grads, _ = res[torch.float64]
    grads32 = res[torch.float32][2]

    def rel(a, b, floor=0.0):
        a, b = a.detach().double().cpu().numpy().reshape(-1), b.detach().double().cpu().numpy().reshape(-1)
        return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30))
    gmax = max(float(g.abs().max()) for g in grads.values())
    worst, worst32, per_var = 0.0, 0.0, {}
    num = den = num32 = 0.0
    for name, g in grads.items():
        e, e32 = rel(m._P.gviews[name], g, 1e-3 * gmax), rel(grads32[name], g, 1e-3 * gmax)
        per_var[name] = {'cuda_vs_f64': e, 'cpu_f32_vs_f64': e32, 'max_abs': float(g.abs().max())}
        worst, worst32 = max(worst, e), max(worst32, e32)
        gd = g.double().reshape(-1)
        num += float(((m._P.gviews[name].detach().cpu().double().reshape(-1) - gd) ** 2).sum())
        num32 += float(((grads32[name].double().reshape(-1) - gd) ** 2).sum())
        den += float((gd ** 2).sum())
    grad_l2 = (num / den) ** 0.5
    ok = grad_l2 <= max(tol, 4.0 * (num32 / den) ** 0.5)
    out = {'loss_rel': abs(float(m._loss) - float(total)) / max(abs(float(total)), 1e-30),
           'max_logit_rel': rel(logits, ologits), 'grad_rel_l2': grad_l2,
           'grad_rel_l2_of_cpu_float32_oracle': (num32 / den) ** 0.5, 'max_grad_rel': worst,
           'max_grad_rel_of_cpu_float32_oracle': worst32,
           'batch': Bp, 'tolerance': tol, 'gradients': per_var, 'largest_gradient': gmax}
    out['ok'] = bool(ok and out['loss_rel'] <= tol and out['max_logit_rel'] <= tol)
    del m
    return out


def cpu_baseline(c, steps=2):
    from oracle import optim_tf1
    warnings.filterwarnings('ignore')
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bc = min(c['B'], c['cpu_batch'])
    o = oracle_for(c, torch.float32)
    name, kw = c['oracle_opt']
    opt = getattr(optim_tf1, name)(**kw)
    xt = torch.from_numpy(c['x'][:Bc])
    lt = torch.from_numpy(c['labels'][:Bc])
    o.train_step(xt, lt, opt, c['lr'])
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        o.train_step(xt, lt, opt, c['lr'])
        ts.append(time.perf_counter() - t)
    return {'value': Bc / float(np.mean(ts)), 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
            'sample': '{} full training steps of batch {} (of {}) after 1 warm-up, torch-CPU oracle, {} threads'.format(
                steps, Bc, c['B'], cores)}


def spmm_roofline(model, c, iters=40):
    """Recurrence step X_k = 2 L X_(k-1) - X_(k-2) of the stage with the largest feature slab B * M * Fin."""
    stages = list(model._encoder) + [d for _, d, _ in getattr(model, '_decoder', [])]
    st = max((s for s in stages if s.K > 1), key=lambda s: s.graph.M * s.Fin)
    g, Fin, B = st.graph, st.Fin, c['B']
    M = g.M
    X = 4 * B * M * Fin
    nsets = max(2, min(8, int(1.4e9 // (3 * X))))
    dev = model.device
    xs = [torch.randn(B, M, Fin, device=dev) for _ in range(nsets)]
    ys = [torch.randn(B, M, Fin, device=dev) for _ in range(nsets)]
    outs = [torch.empty(B, M, Fin, device=dev) for _ in range(nsets)]
    for i in range(nsets):
        ops.spmm(g, xs[i], ys[i], None, out=outs[i], alpha=2.0, beta=-1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.spmm(g, xs[i % nsets], ys[i % nsets], None, out=outs[i % nsets], alpha=2.0, beta=-1.0)
    e1.record()
    e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    alg = 3 * X + g.bytes_stored()
    peak, how = bench.peaks()
    return {'bound': 'hbm', 'achieved': alg / t / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': alg / t / 1e9 / peak, 'traffic': None,
            'kernel': 'ChebConv recurrence step of stage {} (B={} M={} Fin={}, {})'.format(
                st.scope, B, M, Fin, 'ELL width {}'.format(g.fwd[3]) if g.is_ell else 'CSR'),
            'algorithmic_bytes_per_launch': alg, 'avg_launch_us': t * 1e6, 'peak_source': how,
            'timing': '{} launches back to back over {} rotating operand sets ({} MB)'.format(iters, nsets, nsets * 3 * X // 10 ** 6)}


def run(n, steps, warm=3):
    c = config(n)
    par = parity(c)
    model = build_model(c, c['B'], True, '_cfg_bench')
    x_pin = torch.from_numpy(c['x']).pin_memory()
    l_pin = torch.from_numpy(c['labels']).pin_memory()
    nxt = (x_pin, l_pin)
    for _ in range(warm):
        float(model.train_step(x_pin, l_pin, next_batch=nxt)[1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = bench.e2e_loop(model, x_pin, l_pin, nxt, steps)
    e1.record()
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1) * 1e-3
    xd, ld = model._to_device(x_pin), model._labels_to_device(l_pin)
    model._gx.copy_(xd)
    model._gl.copy_(ld)
    for _ in range(warm):
        model._advance_hyper()
        model._step_graph_resident()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        model._advance_hyper()
        model._step_graph_resident()
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) * 1e-3
    from deepsphere_tf1_b200 import _lib
    n0 = _lib.kernel_launches()
    model._advance_hyper()
    model._step_device(xd, ld)
    torch.cuda.synchronize()
    per_step = _lib.kernel_launches() - n0
    roof = spmm_roofline(model, c)
    out = {'metric': 'train samples/s', 'value': c['B'] * steps / t_dev, 'unit': 'samples/s', 'n_gpus': 1, 'steps': steps,
           'warmup': warm, 'ms_per_step': t_dev / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': 'config {}: {}'.format(n, c['name']), 'global_batch': c['B'], 'cuda_graph': True},
           'e2e': {'value': c['B'] * steps / t_e2e, 'unit': 'samples/s', 'ms_per_step': t_e2e / steps * 1e3,
                   'h2d_bytes_per_step': int(x_pin.numel() * 4 + l_pin.numel() * 4 + 32), 'd2h_bytes_per_step': 4},
           'gpu_launches': int(per_step * steps), 'roofline': roof, 'parity': par, 'final_loss': last}
    del model
    torch.cuda.empty_cache()
    out['cpu_baseline'] = cpu_baseline(c)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='1,3,5')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    lines = []
    for n in [int(v) for v in args.configs.split(',')]:
        line = json.dumps(run(n, args.steps))
        print(line, flush=True)
        lines.append(line)
    if args.out:
        open(args.out, 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
