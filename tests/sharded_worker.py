"""torchrun worker of tests/test_vertex_sharded.py: the vertex-sharded model on N GPUs against the single-device
model (same weights, same full-sphere input): logits, loss, every gradient and three optimizer steps."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from deepsphere_tf1_b200 import models, optimizers  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl')

nsides = [16, 8, 4, 4]
kw = dict(nsides=nsides, new=False, F=[8, 16, 1], K=[5, 4, 3], batch_norm=[True, True, False], M=[], statistics='mean',
          num_feat_in=2, regression=True, num_epochs=1, batch_size=1, eval_frequency=10 ** 9, seed=5, verbose=False,
          regularization=1e-3, scheduler=lambda step: 1e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8),
          dir_name='_sharded_test')
rs = np.random.RandomState(1)
x = rs.randn(1, 12 * 16 * 16, 2).astype(np.float32)
y = rs.randn(1).astype(np.float32)

# single-device reference: a group of its own so that it does not all-reduce with the other ranks
solo = dist.new_group([rank]) if world > 1 else None
groups = [dist.new_group([r]) for r in range(world)]
ref = models.deepsphere(process_group=groups[rank], **kw)
shd = models.deepsphere(vertex_shard=True, **kw)
assert shd.get_nbr_var() == ref.get_nbr_var()
shd._P.w.copy_(ref._P.w)


def close(a, b, name, rtol=1e-4):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, '{}: {:.3e} (scale {:.3e})'.format(name, err, scale)
    return err


for step in range(3):
    for m in (ref, shd):
        m._advance_hyper()
        xd, yd = m._to_device(x), m._labels_to_device(y)
        m._step_device(xd, yd)
    e1 = close(shd._loss, ref._loss, 'loss step {}'.format(step))
    e2 = close(shd._P.g, ref._P.g, 'gradients step {}'.format(step))
    e3 = close(shd._P.w, ref._P.w, 'weights step {}'.format(step))
    for k in ref._state:
        close(shd._state[k], ref._state[k], k)
    if rank == 0:
        print('step {}: loss {:.6f} err loss {:.1e} grad {:.1e} weights {:.1e}'.format(step, float(ref._loss), e1, e2, e3))
pr = ref.predict(x)
ps = shd.predict(x)
close(torch.as_tensor(ps), torch.as_tensor(pr), 'predict')
dist.barrier()
if rank == 0:
    print('SHARDED PARITY OK on {} GPUs'.format(world))
dist.destroy_process_group()
