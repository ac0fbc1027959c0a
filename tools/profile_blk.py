import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deepsphere_tf1_b200 import ops
from deepsphere_tf1_b200.graph import DeviceGraph
dev = torch.device('cuda', 0)
L, p = bench.build_laplacians()
import sys as _s
g = DeviceGraph(L[1], dev, blocks='--blocks' in _s.argv, staged='--blocks' not in _s.argv)
ops.set_blocks('--blocks' in _s.argv)
ops.set_staged('--blocks' not in _s.argv)
B, M, Fin = 16, g.M, 16
x1, x0 = torch.randn(B, M, Fin, device=dev), torch.randn(B, M, Fin, device=dev)
out = torch.empty_like(x1)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for _ in range(3):
    ops.fill(flush, 0.0)
    ops.spmm(g, x1, x0, None, out=out, alpha=2.0, beta=-1.0)
torch.cuda.synchronize()
print('ok')
