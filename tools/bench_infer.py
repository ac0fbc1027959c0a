"""Inference (`predict` path: `_forward(training=False)`) of the cosmo benchmark configuration, batch 16 resident in HBM:
samples/s in float32 and in the BF16 mode, with the max-pool fused into the contraction's epilogue (default) and without.
    python tools/bench_infer.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import models, ops, optimizers, utils  # noqa: E402

torch.cuda.set_device(0)
B = 16
x, labels = bench.synthetic_batch(B, 0)
out = []
for dtype in ('float32', 'bfloat16'):
    m = models.deepsphere(
        nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
        F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1, dtype=dtype,
        batch_size=B, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False,
        scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_infer')
    xd = m._to_device(x)
    ref = None
    for pool4 in (True, False):
        ops.USE_POOL4 = pool4
        for _ in range(3):
            logits, _ = m._forward(xd, training=False, save=False)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            logits, _ = m._forward(xd, training=False, save=False)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        same = None
        if ref is None:
            ref = logits.clone()
        else:
            same = bool(torch.equal(ref, logits))
        out.append({'dtype': dtype, 'pool_fused_into_contraction': pool4, 'ms_per_batch': ms, 'samples_per_s': B / ms * 1e3,
                    'logits_bit_identical_to_fused': same})
        print(json.dumps(out[-1]), flush=True)
    ops.USE_POOL4 = True
    del m
