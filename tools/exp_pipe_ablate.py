#!/usr/bin/env python
"""Ablations of the pipelined sparse product (csrc/sparse_pipe.cu) at cosmo layers 2 / 3: which part of the work costs
the time?  DSB200_PIPE_ABLATE bits: 1 = one gather instead of W, 2 = no halo copies, 4 = records loaded once,
8 = no stores, 16 = no arithmetic (consumers store the staged x row).  Results with ablations are of course wrong; this is a timing experiment."""
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
rbuf = torch.ones(512 * 1024 * 1024 // 4, device=dev)
Fins = [1] + bench.F[:-1]


def timeit(fn, nsets=6, reps=30):
    """Back to back over rotating operand sets (together several times L2): steady state, no flush kernel; CUDA event
    resolution on this box is ~2 us, too coarse for single launches."""
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nsets)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


for li in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '2,3').split(',')]:
    M, Fin, B = L[li - 1].shape[0], Fins[li - 1], 16
    xs = [torch.randn(B, M, Fin, device=dev) for _ in range(6)]
    ys = [torch.randn(B, M, Fin, device=dev) for _ in range(6)]
    outs = [torch.empty(B, M, Fin, device=dev) for _ in range(6)]
    print('layer {}: B={} M={} Fin={}'.format(li, B, M, Fin))
    for tb in (8192,):
        os.environ['DSB200_PIPE_TILE_BYTES'] = str(tb)
        g = DeviceGraph(L[li - 1], dev)
        nbytes = 3 * 4 * B * M * Fin + 8 * g.nnz
        if ops._pipe(g, 'fwd', B, M, Fin) is None:
            print('  tile {} B: no plan'.format(tb))
            continue
        for dense in (True,):
            ops.set_pipe(True, dense=dense)
            for order in (0, 1, 2):
                for stages in (3,):
                    for ab in (0, 2, 8, 16):
                        os.environ['DSB200_PIPE_ABLATE'] = str(ab)
                        os.environ['DSB200_PIPE_ORDER'] = str(order)
                        os.environ['DSB200_PIPE_STAGES'] = str(stages)
                        t = timeit(lambda i: ops.spmm(g, xs[i], ys[i], None, out=outs[i], alpha=2.0, beta=-1.0))
                        print('  dense {}  order {}  stages {}  ablate {:2d}: {:7.1f} us ({:4.1f}%)'.format(
                            int(dense), order, stages, ab, t * 1e6, 100 * nbytes / t / 1e9 / peak))
