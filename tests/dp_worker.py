"""torchrun worker of tests/test_dp_parity.py: data-parallel product code on N GPUs (each rank feeds B/N samples:
SyncBN sums, the Gram all-reduce of the fused input layer, the 1/world pre-division of the gradients that SyncBN already
made global, one flat gradient all-reduce) against the SAME product code on one GPU at the global batch B -- loss, every
gradient, weights after three TF1-Adam steps, batch-norm moving statistics -- and, on rank 0, against the CPU oracle.
Covers the cosmo stack in miniature (fused input layer + pipelined sparse products + max pooling) and a classifier with a
fully connected head and regularisation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from deepsphere_tf1_b200 import models, optimizers, utils  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl')
solo = [dist.new_group([r]) for r in range(world)][rank]           # a group of one: the single-GPU reference

B = 8 * world if world <= 2 else 2 * world                           # global batch, B / world per rank


def close(a, b, name, rtol=1e-4, floor=0.0):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    scale = max(np.abs(b).max(), floor, 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, '{}: {:.3e} (scale {:.3e})'.format(name, err, scale)
    return err


def run(tag, sphere, kw, x, labels, cuda_graph):
    common = dict(num_epochs=1, eval_frequency=10 ** 9, seed=3, verbose=False, new=False,
                  scheduler=lambda step: 1e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8))
    ref = models.deepsphere(process_group=solo, batch_size=B, dir_name='_dp_ref', **sphere, **kw, **common)
    dp = models.deepsphere(batch_size=B // world, dir_name='_dp', cuda_graph=cuda_graph, **sphere, **kw, **common)
    dp._P.w.copy_(ref._P.w)
    assert ref.world_size == 1 and dp.world_size == world
    Bl = B // world
    xs, ls = x[rank * Bl:(rank + 1) * Bl], labels[rank * Bl:(rank + 1) * Bl]
    oracle = None
    if rank == 0:
        from oracle.model import OracleCgcnn
        from oracle import optim_tf1
        okw = {k: v for k, v in kw.items()}
        oracle = OracleCgcnn(ref.L, p=ref.p, **okw)
        for name, v in oracle.params.items():
            oracle.params[name] = ref._P.views[name].detach().cpu().reshape(v.shape).clone()
        oopt = optim_tf1.Adam(epsilon=1e-8)
    for step in range(3):
        ref._advance_hyper()
        ref._step_device(ref._to_device(x), ref._labels_to_device(labels))
        dp.train_step(xs, ls)
        e1 = close(dp._loss, ref._loss, tag + ' loss step {}'.format(step))
        gmax = float(ref._P.g.abs().max())
        e2 = 0.0
        if not cuda_graph or step == 0:
            for name in ref._P.gviews:
                e2 = max(e2, close(dp._P.gviews[name], ref._P.gviews[name], tag + ' grad ' + name, floor=1e-3 * gmax))
        e3 = close(dp._P.w, ref._P.w, tag + ' weights step {}'.format(step), rtol=2e-4)
        for k in ref._state:
            close(dp._state[k], ref._state[k], tag + ' ' + k)
        if oracle is not None:
            lo, _ = oracle.train_step(torch.from_numpy(x), torch.from_numpy(labels), oopt, 1e-2)
            assert abs(lo - float(dp._loss)) <= 1e-4 * max(1.0, abs(lo)), (tag, step, lo, float(dp._loss))
            print('{} step {}: loss {:.6f} (oracle {:.6f}) err loss {:.1e} grad {:.1e} weights {:.1e}'.format(
                tag, step, float(dp._loss), lo, e1, e2, e3), flush=True)
    if oracle is not None:
        for name, v in oracle.params.items():
            close(dp._P.views[name].reshape(v.shape), v, tag + ' oracle weights ' + name, rtol=5e-4)
    dp._graph_exec = None
    torch.cuda.synchronize()
    dist.barrier()


rs = np.random.RandomState(4)
# cosmo stack in miniature: Fin = 1 (fused input layer, Gram batch norm), K = 5, max pool, mean statistic
nsides = [64, 32, 16, 16]
sphere = dict(nsides=nsides, indexes=utils.nside2indexes(nsides, 2))
kw = dict(F=[16, 32, 2], K=[5, 5, 5], batch_norm=[True, True, True], M=[], statistics='mean')
x = (rs.randn(B, 1024, 1) * 3.6).astype(np.float32)
labels = rs.randint(0, 2, size=B).astype(np.int32)
from deepsphere_tf1_b200 import ops  # noqa: E402
ops.PIPE_MIN_ITEMS = 1               # the pipelined sparse product (the bench's kernel) also at this small size
run('cosmo-mini eager', sphere, kw, x, labels, False)
run('cosmo-mini graph', sphere, kw, x, labels, True)
# classifier: several input features, average pooling, fully connected head, regularisation
sphere = dict(nsides=[8, 4, 2])
kw = dict(F=[8, 16], K=[4, 4], batch_norm=[True, True], M=[12, 5], pool='average', num_feat_in=3, regularization=0.05)
x = rs.randn(B, 768, 3).astype(np.float32)
labels = rs.randint(0, 5, size=B).astype(np.int32)
run('classifier eager', sphere, kw, x, labels, False)
if rank == 0:
    print('DP PARITY OK on {} GPUs'.format(world), flush=True)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
