"""One forward and one backward launch of the fused tile recurrence (csrc/cheb_tile.cu) at a cosmo layer shape, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:cheb_tile --launch-skip 2 --launch-count 2 \
      -o gpurun_out/tile python tools/profile_tile.py [layer]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops, utils  # noqa: E402
from deepsphere_tf1_b200.graph import DeviceGraph  # noqa: E402

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ops.TILE_BWD_MAX_ITEMS = ops.TILE_FWD_MAX_ITEMS = 1 << 30
ops.TILE_MIN_ITEMS = 1
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
L, p = bench.build_laplacians()
idx = utils.nside2indexes(bench.NSIDES, bench.ORDER)
level = layer - 1
B, F = [(1, 16), (16, 16), (16, 32), (16, 64), (16, 64), (16, 64)][level]
g = DeviceGraph(L[level], dev, nested=(bench.NSIDES[level], idx[level]))
K = 5
x = torch.randn(B, g.M, F, device=dev)
G = torch.randn(K, B, g.M, F, device=dev)
for _ in range(2):
    ops.cheb_basis(g, x, K)
    ops.cheb_basis_bwd(g, G)
torch.cuda.synchronize()
print('ok')
