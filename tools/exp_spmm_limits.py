#!/usr/bin/env python
"""What bounds the gather SpMM?  Same kernel, same sizes (cosmo layer shapes), different column patterns:
real HEALPix neighbours / every slot = the row itself (pure L1 hits) / a band around the diagonal / random columns,
next to a plain streaming kernel with the same three streams (2 reads + 1 write, no gather)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

ops.set_pipe(False)          # this experiment is about the gather kernels

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


def repack(g, col):
    _, _, val, width = g.fwd
    col = col.to(torch.int32).contiguous()
    g.fwd = (None, col, val, width)
    rec = torch.empty_like(g.fwd_rec)
    from deepsphere_tf1_b200._lib import call
    call('ell_pack', col, val, width, g.M, rec)
    g.fwd_rec = rec


for li in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '2,3').split(',')]:
    g = DeviceGraph(L[li - 1], dev)
    M, Fin = g.M, Fins[li - 1]
    nbytes = 3 * 4 * B * M * Fin + 8 * g.nnz
    x, x1 = torch.randn(B, M, Fin, device=dev), torch.randn(B, M, Fin, device=dev)
    out = torch.empty_like(x)
    print('layer {}: B={} M={} Fin={}  {:.1f} MB, roofline {:.1f} us'.format(li, B, M, Fin, nbytes / 1e6, nbytes / peak / 1e3))
    t = timeit(lambda: torch.add(x1, x, alpha=-1.0, out=out))
    print('  torch add (2 reads + 1 write, no gather)   : {:7.1f} us ({:4.1f}% of the 3-stream bytes)'.format(
        t * 1e6, 100 * (nbytes - 8 * g.nnz) / t / 1e9 / peak))
    real = g.fwd[1].clone()
    rows = torch.arange(M, device=dev).view(M, 1)
    pats = {'healpix': real,
            'self (col = row)': rows.expand(M, 9),
            'band (row-4..row+4)': (rows + torch.arange(-4, 5, device=dev).view(1, 9)).clamp(0, M - 1),
            'random': torch.randint(0, M, (M, 9), device=dev)}
    for name, col in pats.items():
        repack(g, col)
        for label, rec_on, sb in (('row-major', False, 1), ('records', True, 1), ('records SB4', True, 4)):
            ops.set_records(rec_on)
            os.environ['DSB200_SPMM_SB'] = str(sb)
            t = timeit(lambda: ops.spmm(g, x1, x, None, out=out, alpha=2.0, beta=-1.0))
            print('  {:<22s} {:<10s}: {:7.1f} us ({:4.1f}%)'.format(name, label, t * 1e6, 100 * nbytes / t / 1e9 / peak))
