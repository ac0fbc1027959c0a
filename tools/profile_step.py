#!/usr/bin/env python
"""One eager (no CUDA graph) cosmo training step for ncu: two warm-up steps, then the profiled
step between cudaProfilerStart/Stop (run ncu with --profile-from-start off).
  python tools/profile_step.py [--batch 16] [--infer]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import _lib, models, optimizers, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--infer', action='store_true')
ap.add_argument('--dtype', default='float32', choices=['float32', 'bfloat16'])
args = ap.parse_args()
torch.cuda.set_device(0)
model = models.deepsphere(
    nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
    F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
    batch_size=args.batch, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False, dtype=args.dtype,
    scheduler=lambda step: optimizers.exponential_decay(2e-4, step, 1, 0.999),
    optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_profile')
x, labels = bench.synthetic_batch(args.batch, 100)
xd, ld = model._to_device(x), model._labels_to_device(labels)
for _ in range(2):
    model._advance_hyper()
    model._step_device(xd, ld)
torch.cuda.synchronize()
n0 = _lib.kernel_launches()
torch.cuda.profiler.start()
if args.infer:
    model._forward(xd, training=False, save=False)
else:
    model._advance_hyper()
    model._step_device(xd, ld)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled launches:', _lib.kernel_launches() - n0, 'loss', float(model._loss))
