#!/usr/bin/env python
"""Recurrence-step timings at the cosmo layer shapes: gather kernel (packed records) vs the bulk-asynchronous pipelined
kernel (csrc/sparse_pipe.cu), forward form (X_k = 2 L X_(k-1) - X_(k-2)) and in-place adjoint form
(A_k = G_k + 2 L^T A_(k+1) - A_(k+2)), next to a plain 3-stream kernel of the same size.  L2 flushed, CUDA events.
usage: bench_spmm.py [layers, e.g. 1,2,3,4,5]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
L, p = bench.build_laplacians()
peak, _ = bench.peaks()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
Fins = [1] + bench.F[:-1]


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        ops.fill(flush, 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


for li in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '1,2,3,4,5').split(',')]:
    g = DeviceGraph(L[li - 1], dev)
    M, Fin, B = g.M, Fins[li - 1], 16
    if li == 1:
        Fin, B = 16, 1                                 # the input layer runs sample-interleaved: [1, M, 16]
    X = 4 * B * M * Fin
    x, x1, x2 = [torch.randn(B, M, Fin, device=dev) for _ in range(3)]
    out = torch.empty_like(x)
    print('layer {}: B={} M={} Fin={}'.format(li, B, M, Fin))
    t = timeit(lambda: torch.add(x1, x, alpha=-1.0, out=out))
    print('  plain 3-stream kernel (torch add)  : {:7.1f} us ({:4.1f}% of peak)'.format(t * 1e6, 100 * 3 * X / t / 1e9 / peak))
    for form in ('fwd', 'adj'):
        nbytes = (3 if form == 'fwd' else 4) * X + 8 * g.nnz
        res = {}
        for label, pipe, dense in (('gather', False, False), ('pipelined', True, False), ('pipe+dense', True, True)):
            ops.set_pipe(pipe, dense=dense)
            if pipe and ops._pipe(g, 'fwd', B, M, Fin) is None:
                continue
            if form == 'fwd':
                fn = lambda: ops.spmm(g, x1, x, None, out=out, alpha=2.0, beta=-1.0)
            else:
                fn = lambda: ops.spmm(g, x1, out, x2, out=out, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True)
            out.copy_(x)
            fn()
            res[label] = out.clone()
            t = timeit(fn)
            print('  {} {:<10s}: {:7.1f} us  {:6.1f} MB  {:6.0f} GB/s ({:4.1f}% of peak; roofline {:.1f} us)'.format(
                form, label, t * 1e6, nbytes / 1e6, nbytes / t / 1e9, 100 * nbytes / t / 1e9 / peak, nbytes / peak / 1e3))
        for k in ('pipelined', 'pipe+dense'):
            if k in res and not torch.equal(res['gather'], res[k]):
                print('  {} differs from gather: max abs {:.2e}'.format(k, (res['gather'] - res[k]).abs().max().item()))
ops.set_pipe(True, dense=True)
