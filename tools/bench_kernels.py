#!/usr/bin/env python
"""Per-kernel timings of the libdsb200 launches at the shapes of the cosmo configuration
(BASELINE config 2, batch 16): CUDA events on the launching stream, L2 flushed (512 MB rewritten)
before every timed launch, mean of `--iters` launches.  Prints algorithmic bytes (SURVEY 8d
conventions), achieved GB/s and the fraction of the measured HBM copy bandwidth.
  python tools/bench_kernels.py [--batch 16] [--iters 10] [--layers 1,2,3] [--json out.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--layers', default='1,2,3,4,5,6')
ap.add_argument('--json', default=None)
ap.add_argument('--no-tc', action='store_true')
args = ap.parse_args()

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
if args.no_tc:
    ops.set_tensor_cores(False)
L, p = bench.build_laplacians()
peak, _ = bench.peaks()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
B, K = args.batch, bench.KCHEB
Fins = [1] + bench.F[:-1]
rows = []


def timeit(name, layer, nbytes, fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(args.iters):
        ops.fill(flush, 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.mean(ts))
    gbs = nbytes / t / 1e9
    rows.append({'layer': layer, 'kernel': name, 'us': t * 1e6, 'MB': nbytes / 1e6, 'GBps': gbs, 'frac': gbs / peak})
    print('L{} {:<28s} {:9.1f} us {:9.1f} MB {:8.1f} GB/s {:5.1f}%'.format(layer, name, t * 1e6, nbytes / 1e6, gbs, 100 * gbs / peak))


for li in [int(v) for v in args.layers.split(',')]:
    i = li - 1
    g = DeviceGraph(L[i], dev)
    M, Fin, Fout, pool = g.M, Fins[i], bench.F[i], p[i]
    R = B * M
    X = 4 * R * Fin
    Y = 4 * R * Fout
    x = torch.randn(B, M, Fin, device=dev)
    x1 = torch.randn(B, M, Fin, device=dev)
    out = torch.empty_like(x)
    W = torch.randn(Fin * K, Fout, device=dev) * 0.1
    timeit('spmm step (k>=2)', li, 3 * X + 8 * g.nnz, lambda: ops.spmm(g, x1, x, None, out=out, alpha=2.0, beta=-1.0))
    xi, x1i, outi = x.view(1, M, B * Fin), x1.view(1, M, B * Fin), out.view(1, M, B * Fin)
    timeit('spmm step, interleaved', li, 3 * X + 8 * g.nnz, lambda: ops.spmm(g, x1i, xi, None, out=outi, alpha=2.0, beta=-1.0))
    basis = ops.cheb_basis(g, x, K)
    timeit('cheb_basis (4 steps)', li, (2 + 3 * 3) * X + 4 * 8 * g.nnz, lambda: ops.cheb_basis(g, x, K))
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y = ops.contract_fwd(x, basis, W, K, Fout, sums)
    timeit('contract_fwd (+bn sums)', li, K * X + Y, lambda: ops.contract_fwd(x, basis, W, K, Fout, sums))
    dy = torch.randn(B, M, Fout, device=dev)
    dW = torch.zeros(Fin * K, Fout, device=dev)
    timeit('contract_bwd_weight', li, K * X + Y, lambda: ops.contract_bwd_weight(x, basis, dy, dW, K))
    if li > 1:
        G = ops.contract_bwd_data(dy, W, K, Fin)
        timeit('contract_bwd_data', li, K * X + Y, lambda: ops.contract_bwd_data(dy, W, K, Fin))
        timeit('cheb_basis_bwd (4 steps)', li, (4 * 3 + 3) * X + 4 * 8 * g.nnz, lambda: ops.cheb_basis_bwd(g, G))
    if li < 6:
        mean, rstd = torch.zeros(Fout, device=dev), torch.ones(Fout, device=dev)
        bias = torch.zeros(Fout, device=dev)
        o = ops.bn_act_pool_fwd(y, mean, rstd, bias, pool, 'max', 'relu')
        timeit('bn_act_pool_fwd', li, Y + Y // pool, lambda: ops.bn_act_pool_fwd(y, mean, rstd, bias, pool, 'max', 'relu'))
        do = torch.randn_like(o)
        gz, s2 = ops.bn_act_pool_bwd(y, mean, rstd, bias, do, pool, 'max', 'relu')
        timeit('bn_act_pool_bwd', li, 2 * Y + Y // pool, lambda: ops.bn_act_pool_bwd(y, mean, rstd, bias, do, pool, 'max', 'relu'))
        timeit('bn_bwd_apply', li, 3 * Y, lambda: ops.bn_bwd_apply(y, mean, rstd, s2, float(R), gz))
    del x, x1, out, basis, y, dy
    torch.cuda.empty_cache()

print('total {:.1f} us'.format(sum(r['us'] for r in rows if not r['kernel'].startswith('spmm step'))))
if args.json:
    json.dump(rows, open(args.json, 'w'), indent=1)
