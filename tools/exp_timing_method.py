#!/usr/bin/env python
"""How much of a per-kernel timing is the flush?  Same kernels (plain 3-stream add, gather SpMM, pipelined SpMM at cosmo
layer 2 / 3), three methods: (a) 512 MB rewritten before every launch (dirty L2: its write-backs are charged to the timed
kernel), (b) the same followed by a 512 MB read (clean L2), (c) N launches back to back over rotating operand sets that
together exceed L2 several times (steady state, no flush kernel at all)."""
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


def per_launch(fn, clean, iters=10):
    ts = []
    for _ in range(iters + 2):
        ops.fill(flush, 0.0)
        if clean:
            rbuf.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(0)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts[2:]))


def back_to_back(fn, nsets, reps=40):
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
    g = DeviceGraph(L[li - 1], dev)
    M, Fin, B = g.M, Fins[li - 1], 16
    nsets = 6
    xs = [torch.randn(B, M, Fin, device=dev) for _ in range(nsets)]
    ys = [torch.randn(B, M, Fin, device=dev) for _ in range(nsets)]
    os_ = [torch.empty(B, M, Fin, device=dev) for _ in range(nsets)]
    nbytes = 3 * 4 * B * M * Fin + 8 * g.nnz
    print('layer {}: B={} M={} Fin={}  {:.1f} MB per launch, roofline {:.1f} us; {} operand sets = {:.0f} MB'.format(
        li, B, M, Fin, nbytes / 1e6, nbytes / peak / 1e3, nsets, nsets * nbytes / 1e6))
    kernels = [('plain add', lambda i: torch.add(xs[i], ys[i], alpha=-1.0, out=os_[i]), None)]
    for label, pipe in (('gather spmm', (False, False)), ('pipelined spmm', (True, False)), ('pipe+dense spmm', (True, True))):
        kernels.append((label, lambda i: ops.spmm(g, xs[i], ys[i], None, out=os_[i], alpha=2.0, beta=-1.0), pipe))
    for label, fn, pipe in kernels:
        if pipe is not None:
            ops.set_pipe(pipe[0], dense=pipe[1])
        a = per_launch(fn, False)
        b = per_launch(fn, True)
        c = back_to_back(fn, nsets)
        print('  {:<16s}: dirty flush {:6.1f} us ({:4.1f}%)   clean flush {:6.1f} us ({:4.1f}%)   back to back {:6.1f} us ({:4.1f}%)'.format(
            label, a * 1e6, 100 * nbytes / a / 1e9 / peak, b * 1e6, 100 * nbytes / b / 1e9 / peak, c * 1e6, 100 * nbytes / c / 1e9 / peak))
ops.set_pipe(True, dense=True)
