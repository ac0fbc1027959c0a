#!/usr/bin/env python
"""A few ChebConv recurrence SpMM launches at one cosmo layer shape, for ncu.
  python tools/profile_spmm.py --layer 2 [--interleave]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--layer', type=int, default=2)
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--interleave', action='store_true')
args = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
L, p = bench.build_laplacians()
g = DeviceGraph(L[args.layer - 1], dev)
Fin = ([1] + bench.F)[args.layer - 1]
B, M = args.batch, g.M
shape = (1, M, B * Fin) if args.interleave else (B, M, Fin)
x1, x0 = torch.randn(shape, device=dev), torch.randn(shape, device=dev)
out = torch.empty_like(x1)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
out2 = torch.empty_like(x1)
for _ in range(3):
    ops.fill(flush, 0.0)
    ops.spmm(g, x1, x0, None, out=out, alpha=2.0, beta=-1.0)
torch.cuda.synchronize()
print('ok', shape)
