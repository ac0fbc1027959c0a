#!/usr/bin/env python
"""Column-slab / row-tile sweep of the gather SpMM kernel (env DSB200_SPMM_CB / DSB200_SPMM_TR) in the
sample-major [B, M, F] and the vertex-major ("interleaved", [1, M, B*F]) layouts.  L2 flushed, CUDA events."""
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
B = 16
Fins = [1] + bench.F[:-1]


def timeit(fn, iters=6):
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
    return float(np.mean(ts))


for li in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '2,3').split(',')]:
    g = DeviceGraph(L[li - 1], dev, tiled=False)
    M, Fin = g.M, Fins[li - 1]
    nbytes = 3 * 4 * B * M * Fin + 8 * g.nnz
    x, x1 = torch.randn(B, M, Fin, device=dev), torch.randn(B, M, Fin, device=dev)
    out = torch.empty_like(x)
    xi, x1i, outi = x.view(1, M, B * Fin), x1.view(1, M, B * Fin), out.view(1, M, B * Fin)
    print('layer {}: M={} Fin={}  {:.1f} MB, roofline {:.1f} us'.format(li, M, Fin, nbytes / 1e6, nbytes / peak / 1e3))
    for name, (a, b, o) in (('sample-major', (x1, x, out)), ('vertex-major', (x1i, xi, outi))):
        for cb in (0, 4, 8, 16, 32):
            for tr in (0, 32, 64, 128, 256, 512):
                os.environ['DSB200_SPMM_CB'] = str(cb)
                os.environ['DSB200_SPMM_TR'] = str(tr)
                t = timeit(lambda: ops.spmm(g, a, b, None, out=o, alpha=2.0, beta=-1.0))
                print('  {} CB={:2d} TR={:3d}: {:7.1f} us ({:4.1f}%)'.format(name, cb, tr, t * 1e6, 100 * nbytes / t / 1e9 / peak))
