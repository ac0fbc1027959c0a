"""Data-parallel arithmetic on the CPU with a real 2-rank gloo group: each rank runs the oracle on
its half of the batch with the SAME scheme deepsphere_tf1_b200.models uses on the GPUs --
batch-norm sums all-reduced over the ranks (SyncBN, global count), local loss scaled by 1/world,
gradients summed over the ranks -- and must reproduce the single-process full-batch gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def _model():
    from helpers import healpix_L
    from oracle.model import OracleCgcnn
    L = [healpix_L(2), healpix_L(1)]
    return OracleCgcnn(L, F=[6, 3], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='mean', num_feat_in=2,
                       dtype=torch.float64, regularization=0.0)


def _data():
    rs = np.random.RandomState(0)
    return torch.from_numpy(rs.randn(8, 48, 2)), torch.from_numpy(rs.randint(0, 3, size=8))


def _worker(rank, world, port, out):
    import torch.distributed.nn.functional as dfn
    from oracle import ops as oops
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def sync_bn(x, moving_mean, moving_var, training, momentum=0.9, eps=1e-5):
        # sums over the local rows, all-reduced (differentiably) over the ranks, global count
        n = x.shape[0] * x.shape[1] * world
        s1 = dfn.all_reduce(x.sum(dim=(0, 1)))
        s2 = dfn.all_reduce((x * x).sum(dim=(0, 1)))
        mean = s1 / n
        var = s2 / n - mean * mean
        return (x - mean) * torch.rsqrt(var + eps), moving_mean, moving_var

    oops.batch_normalization = sync_bn
    m = _model()
    x, labels = _data()
    per = x.shape[0] // world
    xs, ls = x[rank * per:(rank + 1) * per], labels[rank * per:(rank + 1) * per]
    P = {k: v.clone().requires_grad_(True) for k, v in m.params.items()}
    logits, _ = m.forward(xs, P, True, {})
    total, _ = m.loss(logits, ls, P)
    (total / world).backward()                       # gradient of the global-batch mean loss
    flat = torch.cat([p.grad.reshape(-1) for p in P.values()])
    dist.all_reduce(flat)                            # the flat gradient all-reduce (sum)
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_two_rank_syncbn_gradient_equals_full_batch(tmp_path):
    out = str(tmp_path / 'grad.pt')
    mp.spawn(_worker, args=(2, 29000 + os.getpid() % 2000, out), nprocs=2, join=True)
    got = torch.load(out)
    m = _model()
    x, labels = _data()
    _, _, grads, _ = m.loss_and_grads(x, labels)
    ref = torch.cat([grads[k].reshape(-1) for k in m.params])
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-12)
