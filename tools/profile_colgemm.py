import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsphere_tf1_b200 import ops
dev = torch.device('cuda', 0)
B, K, M, Fin, Fout = 16, 5, 65536, 16, 32
x = torch.randn(B, M, Fin, device=dev); basis = torch.randn(K - 1, B, M, Fin, device=dev)
W = torch.randn(Fin * K, Fout, device=dev) * 0.1
sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
for _ in range(3):
    ops.contract_fwd(x, basis, W, K, Fout, sums)
torch.cuda.synchronize()
print('ok')
