#!/usr/bin/env python
"""Robustness check: many eager cosmo training steps back to back with a device synchronisation and an error check
every step (tcgen05 / mbarrier protocol bugs show up as launch failures or NaN losses).
  python tools/stress_step.py [--steps 2000] [--flush]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsphere_tf1_b200 import models, ops, optimizers, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=2000)
ap.add_argument('--flush', action='store_true', help='rewrite a 512 MB buffer between steps (cold L2, like under ncu)')
ap.add_argument('--dtype', default='float32', choices=['float32', 'bfloat16'])
args = ap.parse_args()
torch.cuda.set_device(0)
model = models.deepsphere(
    nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
    F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
    batch_size=16, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False, dtype=args.dtype,
    scheduler=lambda step: optimizers.exponential_decay(2e-4, step, 1, 0.999),
    optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_stress')
x, labels = bench.synthetic_batch(16, 100)
xd, ld = model._to_device(x), model._labels_to_device(labels)
flush = torch.empty(512 * 1024 * 1024 // 4, device=model.device) if args.flush else None
bad = 0
for i in range(args.steps):
    model._advance_hyper()
    model._step_device(xd, ld)
    if flush is not None:
        ops.fill(flush, 0.0)
    if i % 50 == 49 or i == args.steps - 1:
        torch.cuda.synchronize()
        loss = float(model._loss)
        if not (loss == loss) or loss > 10:
            bad += 1
            print('step', i, 'suspicious loss', loss)
print('done: {} steps, {} suspicious, final loss {:.6f}'.format(args.steps, bad, float(model._loss)))
