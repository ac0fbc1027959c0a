"""Per-stage timeline of CTA 0 of the row-GEMM kernel (debug timestamps, DSB200_RG_DBG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepsphere_tf1_b200 import ops
dev = torch.device('cuda', 0)
B, K = 16, 5
for (M, Fin, Fout) in ((65536, 16, 32), (16384, 32, 64)):
    x = torch.randn(B, M, Fin, device=dev); basis = torch.randn(K - 1, B, M, Fin, device=dev)
    W = torch.randn(Fin * K, Fout, device=dev) * 0.1
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    for _ in range(2):
        ops.contract_fwd(x, basis, W, K, Fout, sums)
    dbg = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
    os.environ['DSB200_RG_DBG'] = str(dbg.data_ptr())
    ops.contract_fwd(x, basis, W, K, Fout, sums)
    torch.cuda.synchronize()
    os.environ.pop('DSB200_RG_DBG')
    d = dbg.cpu().numpy().reshape(256, 8)
    t0 = d[0, 0]
    print('fwd M={} Fin={} Fout={}: stage | tma: wait-empty begin, end | split: data seen, done | mma: operands ready, committed  (ns from start)'.format(M, Fin, Fout))
    for i in list(range(0, 24)) + list(range(100, 112)):
        r = d[i]
        print('{:4d} | {:7d} {:7d} | {:7d} {:7d} | {:7d} {:7d}'.format(i, *[int(v - t0) if v else -1 for v in r[:6]]))
