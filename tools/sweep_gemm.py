#!/usr/bin/env python
"""Tuning sweep of the tensor-core contraction kernels at the cosmo layer shapes (batch 16): CTAs per SM,
stage size and row packing are read from the environment by the launchers (DSB200_RG_CTAS, DSB200_WG_CTAS,
DSB200_WG_STAGE_KB, DSB200_WG_NOPACK).  L2 flushed before every timed launch, CUDA events.
  python tools/sweep_gemm.py [--layers 2,3,4,5]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--iters', type=int, default=8)
ap.add_argument('--layers', default='2,3,4,5')
args = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
peak, _ = bench.peaks()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
B, K = args.batch, bench.KCHEB
Fins = [1] + bench.F[:-1]
Ms = [262144 // 4 ** i for i in range(6)]


def timeit(fn):
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
    return float(np.mean(ts))


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


for li in [int(v) for v in args.layers.split(',')]:
    i = li - 1
    M, Fin, Fout = Ms[i], Fins[i], bench.F[i]
    R = B * M
    nbytes = 4 * R * (K * Fin + Fout)
    x = torch.randn(B, M, Fin, device=dev)
    basis = torch.randn(K - 1, B, M, Fin, device=dev)
    W = torch.randn(Fin * K, Fout, device=dev) * 0.1
    dy = torch.randn(B, M, Fout, device=dev)
    dW = torch.zeros(Fin * K, Fout, device=dev)
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    print('layer {}: R={} Fin={} Fout={}  {:.1f} MB  roofline {:.1f} us'.format(li, R, Fin, Fout, nbytes / 1e6, nbytes / peak / 1e3))
    for c in (1, 2):
        setenv(DSB200_RG_CTAS=c)
        t = timeit(lambda: ops.contract_fwd(x, basis, W, K, Fout, sums))
        t2 = timeit(lambda: ops.contract_bwd_data(dy, W, K, Fin))
        print('  rowgemm ctas={}: fwd {:7.1f} us ({:4.1f}%)   bwd_data {:7.1f} us ({:4.1f}%)'.format(
            c, t * 1e6, 100 * nbytes / t / 1e9 / peak, t2 * 1e6, 100 * nbytes / t2 / 1e9 / peak))
    setenv(DSB200_RG_CTAS=None)
    for nopack in (1, 0):
        if nopack == 0 and Fin >= 32:
            continue
        for c in (1, 2):
            for kb in (24, 32, 48, 64):
                if c == 2 and kb > 48:
                    continue
                setenv(DSB200_WG_CTAS=c, DSB200_WG_STAGE_KB=kb, DSB200_WG_NOPACK=nopack)
                try:
                    t = timeit(lambda: ops.contract_bwd_weight(x, basis, dy, dW, K))
                    print('  wgrad pack={} ctas={} stage={}KB: {:7.1f} us ({:4.1f}%)'.format(1 - nopack, c, kb, t * 1e6, 100 * nbytes / t / 1e9 / peak))
                except RuntimeError as e:
                    print('  wgrad pack={} ctas={} stage={}KB: {}'.format(1 - nopack, c, kb, e))
    setenv(DSB200_WG_CTAS=None, DSB200_WG_STAGE_KB=None, DSB200_WG_NOPACK=None)
    if li < 6:
        mean, rstd = torch.zeros(Fout, device=dev), torch.ones(Fout, device=dev)
        bias = torch.zeros(Fout, device=dev)
        y = torch.randn(B, M, Fout, device=dev)
        o = ops.bn_act_pool_fwd(y, mean, rstd, bias, 4, 'max', 'relu')
        do = torch.randn_like(o)
        Y = 4 * R * Fout
        t = timeit(lambda: ops.bn_act_pool_bwd(y, mean, rstd, bias, do, 4, 'max', 'relu', want_gz=False))
        print('  bn bwd pass 1 (from y):       {:7.1f} us'.format(t * 1e6))
        t = timeit(lambda: ops.bn_relu_pool_sums(o, do, bias))
        print('  bn bwd pass 1 (pooled only):  {:7.1f} us ({:4.1f}% of {:.1f} MB)'.format(t * 1e6, 100 * (Y // 2) / t / 1e9 / peak, Y / 2e6))
        s = ops.bn_relu_pool_sums(o, do, bias)
        t = timeit(lambda: ops.bn_act_pool_bwd_dy(y, mean, rstd, bias, do, s, float(R), 4, 'max', 'relu'))
        print('  bn bwd pass 2 (dy):           {:7.1f} us ({:4.1f}% of {:.1f} MB)'.format(t * 1e6, 100 * (2 * Y + Y // 4) / t / 1e9 / peak, (2 * Y + Y // 4) / 1e6))
        t = timeit(lambda: ops.bn_act_pool_fwd(y, mean, rstd, bias, 4, 'max', 'relu'))
        print('  bn_act_pool_fwd:              {:7.1f} us ({:4.1f}% of {:.1f} MB)'.format(t * 1e6, 100 * (Y + Y // 4) / t / 1e9 / peak, (Y + Y // 4) / 1e6))
    del x, basis, dy
    torch.cuda.empty_cache()
