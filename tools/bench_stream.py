#!/usr/bin/env python
"""Streaming-bandwidth sanity check of the simple libdsb200 kernels against torch's copy (the driver's
MEASURED_PEAKS method): fill (write only), axpy (2 reads + 1 write), torch copy_ (1 read + 1 write)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsphere_tf1_b200 import ops  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
n = 256 * 1024 * 1024          # 1 GiB of fp32
a = torch.randn(n, device=dev)
b = torch.randn(n, device=dev)


def t(fn, nbytes, name, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    print('{:<28s} {:8.1f} us  {:8.1f} GB/s (best {:8.1f})'.format(name, np.mean(ts) * 1e6, nbytes / np.mean(ts) / 1e9, nbytes / min(ts) / 1e9))


t(lambda: b.copy_(a), 8 * n, 'torch copy_ 1 GiB fp32')
t(lambda: ops.fill(a, 1.0), 4 * n, 'dsb200 fill')
t(lambda: ops.axpy(a, b, 0.5), 12 * n, 'dsb200 axpy')
t(lambda: ops.scale(a, 0.99), 8 * n, 'dsb200 scale')
x = a[:16 * 65536 * 16].view(16, 65536, 16)
t(lambda: ops.vertex_resize(x, 65536, 0.0), 2 * x.numel() * 4, 'vertex_resize copy 67 MB')
