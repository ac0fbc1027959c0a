"""Per-kernel time of one eager vertex-sharded training step on rank 0 (torch.profiler), under torchrun."""
import collections
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsphere_tf1_b200 import models, optimizers  # noqa: E402

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl')
nsides = [nside // 2 ** i for i in range(6)] + [nside // 32]
m = models.deepsphere(nsides=nsides, new=False, F=[16, 32, 64, 64, 64, 1], K=[5] * 6, batch_norm=[True] * 6, M=[],
                      statistics='mean', regression=True, num_epochs=1, batch_size=1, eval_frequency=10 ** 9, seed=0,
                      verbose=False, vertex_shard=world > 1, fused_small=world == 1, cuda_graph=False,
                      scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8),
                      dir_name='_trace_sharded')
rs = np.random.RandomState(0)
x = m._to_device(rs.standard_normal((1, 12 * nside ** 2, 1)).astype(np.float32))
y = m._labels_to_device(rs.standard_normal(1).astype(np.float32))
for _ in range(3):
    m._advance_hyper()
    m._step_device(x, y)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        m._advance_hyper()
        m._step_device(x, y)
    torch.cuda.synchronize()
if rank == 0:
    tot = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != 'CUDA' or not ev.name:
            continue
        name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '').replace('dsb200::', '').split('(')[0]
        d = tot.setdefault(name, [0.0, 0])
        d[0] += ev.device_time
        d[1] += 1
    total = sum(v[0] for v in tot.values()) / 3
    for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
        print('{:10.1f} {:6.1f} {:5.1f}%  {}'.format(us / 3, n / 3, 100 * us / 3 / total, name[:100]))
    print('{:10.1f} total kernel time per step, {} ranks'.format(total, world))
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)
