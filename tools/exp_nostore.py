import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepsphere_tf1_b200 import ops
dev = torch.device('cuda', 0)
peak, _ = bench.peaks()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
B, K = 16, 5
def timeit(fn, iters=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        ops.fill(flush, 0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.mean(ts))
for (M, Fin, Fout) in ((65536, 16, 32), (16384, 32, 64), (4096, 64, 64)):
    x = torch.randn(B, M, Fin, device=dev); basis = torch.randn(K - 1, B, M, Fin, device=dev)
    W = torch.randn(Fin * K, Fout, device=dev) * 0.1; dy = torch.randn(B, M, Fout, device=dev)
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    for ns in (0, 1):
        os.environ['DSB200_RG_NOSTORE'] = str(ns)
        t = timeit(lambda: ops.contract_fwd(x, basis, W, K, Fout, sums))
        t2 = timeit(lambda: ops.contract_bwd_data(dy, W, K, Fin))
        print('M={} Fin={} Fout={} nostore={}: fwd {:6.1f} us  bwd_data {:6.1f} us'.format(M, Fin, Fout, ns, t * 1e6, t2 * 1e6))
