"""Zigzag recurrence (ops.ZIGZAG): the cosmo training step (config 2, batch 16, one CUDA graph) with the sparse products of a
layer sweeping memory all start -> end (0) or in alternating directions (1, 2: the two parities), same process, same model,
CUDA events over `--steps` replays, interleaved rounds so that clock drift hits every setting alike."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import models, ops, optimizers, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=40)
ap.add_argument('--rounds', type=int, default=3)
ap.add_argument('--dtype', default='float32')
args = ap.parse_args()
torch.cuda.set_device(0)
model = models.deepsphere(
    nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
    F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
    batch_size=16, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=True, dtype=args.dtype,
    scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_zigzag')
x, labels = bench.synthetic_batch(16, 0)
xd, ld = model._to_device(x), model._labels_to_device(labels)
res = {0: [], 1: [], 2: []}
loss = {}
for rnd in range(args.rounds):
    for z in (0, 1, 2):
        ops.ZIGZAG = z
        model._graph_exec = None                       # capture the step again with this setting
        for _ in range(5):
            model._advance_hyper()
            model._step_graphed(xd, ld)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            model._advance_hyper()
            model._step_graph_resident()
        e1.record()
        torch.cuda.synchronize()
        res[z].append(e0.elapsed_time(e1) / args.steps)
        loss[z] = float(model._loss)
print(json.dumps({'what': 'ms per training step, cosmo config 2, batch 16, ' + args.dtype, 'steps': args.steps,
                  'zigzag_0': res[0], 'zigzag_1': res[1], 'zigzag_2': res[2], 'last_loss': loss}))
