#!/usr/bin/env python
"""Vertex-sharded training step on a FULL HEALPix sphere (BASELINE config 4 in shape: single-map regression,
cosmo stack, statistics='mean', batch 1), N ranks = N contiguous nested vertex ranges with a 1-ring halo exchange
before every sparse product.  Default Nside=256 (786,432 px; the graph of Nside=1024 takes minutes to build on the
host).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                 tools/bench_sharded.py [--nside 256] [--steps 10]
Prints one JSON line on rank 0: steps/s, ms per step (max over ranks, CUDA events), halo bytes per step."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsphere_tf1_b200 import models, optimizers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nside', type=int, default=256)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--lmax', action='store_true', help='skip the eigenvalue solves (1.57 for every level): start-up only')
ap.add_argument('--graph', action='store_true', help='replay the step (kernels + NCCL send/recv) as one CUDA graph')
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl')
nsides = [args.nside // 2 ** i for i in range(6)] + [args.nside // 32]
t0 = time.time()
m = models.deepsphere(nsides=nsides, new=False, F=[16, 32, 64, 64, 64, 1], K=[5] * 6, batch_norm=[True] * 6, M=[],
                      statistics='mean', regression=True, num_epochs=1, batch_size=1, eval_frequency=10 ** 9, seed=0,
                      verbose=False, vertex_shard=world > 1, fused_small=world == 1, cuda_graph=args.graph,
                      scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8),
                      dir_name='_bench_sharded', lmax=[1.57] * 6 if args.lmax else None)
t_build = time.time() - t0
npix = 12 * args.nside ** 2
rs = np.random.RandomState(0)
x = m._to_device(rs.standard_normal((1, npix, 1)).astype(np.float32))
y = m._labels_to_device(rs.standard_normal(1).astype(np.float32))


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if args.graph:
    m._capture(x, y)
    step = m._step_graph_resident
else:
    step = lambda: m._step_device(x, y)      # noqa: E731
for _ in range(args.warmup):
    m._advance_hyper()
    step()
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    m._advance_hyper()
    step()
e1.record()
sync()
t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=m.device, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
halo = 0
if world > 1:
    for i, st in enumerate(m._encoder):
        if st.sg is not None and st.K > 1:
            rows = sum(len(v) for v in st.sg.send.values())
            halo += rows * st.Fin * 4 * 2 * (st.K - 1)          # forward + adjoint, one exchange per sparse product
if rank == 0:
    ms = float(t[0]) / args.steps * 1e3
    print(json.dumps({'workload': 'full-sphere HEALPix Nside={} ({} px), cosmo stack, regression, batch 1'.format(args.nside, npix),
                      'n_gpus': world, 'cuda_graph': bool(args.graph), 'mode': 'vertex-sharded' if world > 1 else 'single device',
                      'ms_per_step': ms, 'maps_per_s': 1e3 / ms, 'halo_bytes_sent_per_rank_per_step': halo,
                      'graph_build_s': t_build, 'loss': float(m._loss)}))
if world > 1:
    m._graph_exec = None
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
