"""Fused tile recurrence (csrc/cheb_tile.cu) against the per-step pipelined kernels at the cosmo layer shapes: all K - 1 = 4
sparse products of a layer, forward (basis) and backward (adjoint), CUDA events, launches back to back over rotating operand
sets larger than L2.  Prints us per layer-recurrence and the achieved algorithmic GB/s of the fused form
(forward: X read + 4 X written + 8 nnz; backward: 5 X read + X written + 8 nnz)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops, utils  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

ops.TILE_BWD_MAX_ITEMS = ops.TILE_FWD_MAX_ITEMS = 1 << 30
ops.TILE_MIN_ITEMS = 1
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
L, p = bench.build_laplacians()
idx = utils.nside2indexes(bench.NSIDES, bench.ORDER)
K = 5
shapes = [(0, 1, 16), (1, 16, 16), (2, 16, 32), (3, 16, 64), (4, 16, 64), (5, 16, 64)]     # (level, B, F); level 0 = interleaved input
rows = []
for level, B, F in shapes:
    g = DeviceGraph(L[level], dev, nested=(bench.NSIDES[level], idx[level]))
    M = g.M
    X = 4 * B * M * F
    nsets = max(2, min(8, int(1.3e9 // (10 * X)) + 1))
    xs = [torch.randn(B, M, F, device=dev) for _ in range(nsets)]
    Gs = [torch.randn(K, B, M, F, device=dev) for _ in range(nsets)]
    for mode in ('tile', 'steps'):
        ops.set_tile(mode == 'tile')
        if mode == 'tile' and ops._tile(g, 'fwd', K, B, F) is None:
            continue
        for direction in ('fwd', 'bwd'):
            fn = (lambda i: ops.cheb_basis(g, xs[i % nsets], K)) if direction == 'fwd' else \
                 (lambda i: ops.cheb_basis_bwd(g, Gs[i % nsets].clone() if mode == 'steps' else Gs[i % nsets]))
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            iters = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i)
            e1.record()
            e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            alg = (5 * X if direction == 'fwd' else 6 * X) + 8 * g.nnz
            rows.append({'level': level + 1, 'B': B, 'M': M, 'F': F, 'mode': mode, 'dir': direction, 'us': us,
                         'fused_algorithmic_GBps': alg / us / 1e3})
            print('layer {} B={} M={} F={} {:5s} {}: {:8.1f} us   ({:.0f} GB/s of the fused algorithmic bytes{})'.format(
                level + 1, B, M, F, mode, direction, us, alg / us / 1e3,
                ', incl. a clone of G' if (mode == 'steps' and direction == 'bwd') else ''), flush=True)
    del xs, Gs
    torch.cuda.empty_cache()
ops.set_tile(True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], 'w'), indent=1)
