#!/usr/bin/env python
"""What do the contraction kernels spend their time on?  Forward / data-gradient / weight-gradient contraction at the cosmo
layer 2 / 3 shapes, with and without the epilogue stores (DSB200_RG_NOSTORE: wrong results, timing only) and the fused
batch-norm sums.  Back to back over rotating operand sets (several times L2), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
peak, _ = bench.peaks()
B, K = 16, 5


def b2b(fn, nsets, reps=24):
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


for (M, Fin, Fout) in ((65536, 16, 32), (16384, 32, 64)):
    ns = 3
    xs = [torch.randn(B, M, Fin, device=dev) for _ in range(ns)]
    bs = [torch.randn(K - 1, B, M, Fin, device=dev) for _ in range(ns)]
    dys = [torch.randn(B, M, Fout, device=dev) for _ in range(ns)]
    W = torch.randn(Fin * K, Fout, device=dev) * 0.1
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    nbytes = 4 * B * M * (K * Fin + Fout)
    print('M={} Fin={} Fout={}: {:.0f} MB, roofline {:.1f} us'.format(M, Fin, Fout, nbytes / 1e6, nbytes / peak / 1e3))
    for nostore in (0, 1):
        os.environ['DSB200_RG_NOSTORE'] = str(nostore)
        t1 = b2b(lambda i: ops.contract_fwd(xs[i], bs[i], W, K, Fout, sums), ns)
        t2 = b2b(lambda i: ops.contract_fwd(xs[i], bs[i], W, K, Fout, None), ns)
        t3 = b2b(lambda i: ops.contract_bwd_data(dys[i], W, K, Fin), ns)
        print('  nostore={}: fwd+sums {:6.1f} us   fwd {:6.1f} us   bwd_data {:6.1f} us'.format(nostore, t1 * 1e6, t2 * 1e6, t3 * 1e6))
    os.environ['DSB200_RG_NOSTORE'] = '0'
    dW = torch.zeros(Fin * K, Fout, device=dev)
    t4 = b2b(lambda i: ops.contract_bwd_weight(xs[i], bs[i], dys[i], dW, K), ns)
    print('  wgrad {:6.1f} us'.format(t4 * 1e6))
