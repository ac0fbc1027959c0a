#!/usr/bin/env python
"""Packed-record SpMM (spmm_rec_kernel) vs the row-major gather kernel at the cosmo layer shapes: samples per tile
(DSB200_SPMM_SB) x row tile (DSB200_SPMM_TR).  L2 flushed, CUDA events.  usage: sweep_spmm_rec.py [layers, e.g. 2,3,4]"""
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


def timeit(fn, iters=8):
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


for li in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '2,3,4').split(',')]:
    g = DeviceGraph(L[li - 1], dev)
    M, Fin = g.M, Fins[li - 1]
    if li == 1:
        M, Fin, Bs = g.M, 16, 1                       # the input layer runs sample-interleaved: [1, M, 16]
    else:
        Bs = B
    nbytes = 3 * 4 * Bs * M * Fin + 8 * g.nnz
    x, x1 = torch.randn(Bs, M, Fin, device=dev), torch.randn(Bs, M, Fin, device=dev)
    out = torch.empty_like(x)
    print('layer {}: B={} M={} Fin={}  {:.1f} MB, roofline {:.1f} us'.format(li, Bs, M, Fin, nbytes / 1e6, nbytes / peak / 1e3))
    os.environ.pop('DSB200_SPMM_TR', None)
    ops.set_records(False)
    t = timeit(lambda: ops.spmm(g, x1, x, None, out=out, alpha=2.0, beta=-1.0))
    print('  row-major gather kernel     : {:7.1f} us ({:4.1f}%)'.format(t * 1e6, 100 * nbytes / t / 1e9 / peak))
    ref = out.clone()
    ops.set_records(True)
    variants = [int(v) for v in os.environ.get('SWEEP_VARIANTS', '0,3,4,14,15,16,28').split(',')]
    for var in variants:
        for sb in ((1, 2, 4) if var in (0, 14, 15, 16, 28) else (1,)):
            for tr in ((128,) if var == 0 else (64, 128, 256)):
                os.environ['DSB200_SPMM_VARIANT'] = str(var)
                os.environ['DSB200_SPMM_SB'] = str(sb)
                os.environ['DSB200_SPMM_TR'] = str(tr)
                out.zero_()
                t = timeit(lambda: ops.spmm(g, x1, x, None, out=out, alpha=2.0, beta=-1.0))
                ok = torch.equal(out, ref)
                print('  records var={:2d} SB={} TR={:3d}  : {:7.1f} us ({:4.1f}%) {}'.format(var, sb, tr, t * 1e6, 100 * nbytes / t / 1e9 / peak,
                                                                                       '' if ok else 'MISMATCH'))
