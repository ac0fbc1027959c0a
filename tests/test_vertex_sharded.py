"""Vertex-sharded mode: contiguous nested-index vertex ranges, 1-ring halo exchange.

* CPU / gloo, 2 and 4 ranks: the partition, the ghost ordering and the send / receive plan of `ShardedGraph`
  together with `halo.exchange` reproduce the Chebyshev recurrence (and its adjoint with L^T) of the whole
  graph on every rank's own vertices -- the host logic of the mode, no CUDA kernel involved.
* GPU (needs >= 2 devices, `gpurun --gpus 2`): the sharded model against the single-device model on the same
  weights and input: logits, loss and every gradient (tests/sharded_worker.py under torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import free_port  # noqa: E402
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def _graph(nonsym):
    from helpers import healpix_L
    L = sparse.csr_matrix(healpix_L(8), dtype=np.float32)
    if nonsym:                                    # non-symmetric values: the adjoint really needs the rows of L^T
        L = sparse.csr_matrix(L.multiply(sparse.csr_matrix(1.0 + 0.2 * np.random.RandomState(3).rand(*L.shape))), dtype=np.float32)
    return L


def _worker(rank, world, port, nonsym, out):
    from deepsphere_tf1_b200 import halo
    from deepsphere_tf1_b200.graph import ShardedGraph
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    L = _graph(nonsym)
    sg = ShardedGraph(L, rank, world, None)
    M, Ml, K, F = L.shape[0], sg.Ml, 5, 3
    x = torch.from_numpy(np.random.RandomState(0).randn(M, F))
    lo = rank * Ml
    gather = lambda rows, idx: rows[torch.from_numpy(idx.astype(np.int64))].contiguous()

    def product(A_local, v_owned):
        e = torch.zeros(sg.Me, F, dtype=torch.float64)
        e[:Ml] = v_owned
        halo.exchange(sg, e, gather=gather)
        return torch.from_numpy(A_local.astype(np.float64)[:Ml] @ e.numpy())

    planes = [x[lo:lo + Ml].clone()]
    planes.append(product(sg.L_local, planes[0]))
    for k in range(2, K):
        planes.append(2 * product(sg.L_local, planes[-1]) - planes[-2])
    adj = product(sg.LT_local, planes[0])                      # L^T x on the owned rows
    torch.save((torch.stack(planes), adj, torch.from_numpy(sg.ghost_ids)), '{}.{}'.format(out, rank))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,nonsym', [(2, False), (4, False), (2, True)])
def test_halo_exchange_reproduces_the_global_recurrence(tmp_path, world, nonsym):
    out = str(tmp_path / 'planes')
    mp.spawn(_worker, args=(world, 29000 + (os.getpid() * 7 + world) % 2000, nonsym, out), nprocs=world, join=True)
    L = _graph(nonsym).astype(np.float64)
    M, K, F = L.shape[0], 5, 3
    x = np.random.RandomState(0).randn(M, F)
    ref = [x, L @ x]
    for k in range(2, K):
        ref.append(2 * (L @ ref[-1]) - ref[-2])
    ref = np.stack(ref)
    adj = L.T @ x
    Ml = M // world
    for r in range(world):
        planes, a, ghosts = torch.load('{}.{}'.format(out, r))
        np.testing.assert_allclose(planes.numpy(), ref[:, r * Ml:(r + 1) * Ml], rtol=0, atol=1e-9)
        np.testing.assert_allclose(a.numpy(), adj[r * Ml:(r + 1) * Ml], rtol=0, atol=1e-9)
        g = ghosts.numpy()
        assert np.all(np.diff(g) > 0) and not np.any((g >= r * Ml) & (g < (r + 1) * Ml))


def _wide_worker(rank, world, port, nonsym, out):
    """(K - 1)-ring halo: ONE exchange per recurrence, every product on all local rows (graph.ShardedGraph(rings=K-1))."""
    from deepsphere_tf1_b200 import halo
    from deepsphere_tf1_b200.graph import ShardedGraph
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    L = _graph(nonsym)
    K, F = 5, 3
    sg = ShardedGraph(L, rank, world, None, rings=K - 1)
    M, Ml = L.shape[0], sg.Ml
    lo = rank * Ml
    gather = lambda rows, idx: rows[torch.from_numpy(idx.astype(np.int64))].contiguous()
    x = np.random.RandomState(0).randn(M, F)
    G = np.random.RandomState(1).randn(K, M, F)
    A, AT = sg.L_local.astype(np.float64), sg.LT_local.astype(np.float64)
    e = torch.zeros(sg.Me, F, dtype=torch.float64)
    e[:Ml] = torch.from_numpy(x[lo:lo + Ml])
    halo.exchange(sg, e, gather=gather)                                   # the only exchange of the forward recurrence
    planes = [e.numpy(), A @ e.numpy()]
    for k in range(2, K):
        planes.append(2 * (A @ planes[-1]) - planes[-2])
    g = torch.zeros(K, sg.Me, F, dtype=torch.float64)
    g[:, :Ml] = torch.from_numpy(G[:, lo:lo + Ml])
    for k in range(K):
        halo.exchange(sg, g[k], gather=gather)                            # (the model exchanges dy once and derives the G_k locally)
    a = [None] * K
    a[K - 1] = g[K - 1].numpy()
    for k in range(K - 2, -1, -1):
        a[k] = g[k].numpy() + (2.0 if k >= 1 else 1.0) * (AT @ a[k + 1]) - (a[k + 2] if k + 2 <= K - 1 else 0.0)
    torch.save((torch.from_numpy(np.stack([p[:Ml] for p in planes])), torch.from_numpy(a[0][:Ml]),
                torch.from_numpy(sg.ghost_ids), torch.from_numpy(sg.ghost_ring)), '{}.{}'.format(out, rank))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,nonsym', [(2, False), (4, True)])
def test_wide_halo_needs_one_exchange_per_recurrence(tmp_path, world, nonsym):
    out = str(tmp_path / 'wide')
    mp.spawn(_wide_worker, args=(world, 31000 + (os.getpid() * 11 + world) % 2000, nonsym, out), nprocs=world, join=True)
    L = _graph(nonsym).astype(np.float64)
    M, K, F = L.shape[0], 5, 3
    x = np.random.RandomState(0).randn(M, F)
    G = np.random.RandomState(1).randn(K, M, F)
    ref = [x, L @ x]
    for k in range(2, K):
        ref.append(2 * (L @ ref[-1]) - ref[-2])
    ref = np.stack(ref)
    A = [None] * K
    A[K - 1] = G[K - 1]
    for k in range(K - 2, -1, -1):
        A[k] = G[k] + (2.0 if k >= 1 else 1.0) * (L.T @ A[k + 1]) - (A[k + 2] if k + 2 <= K - 1 else 0.0)
    Ml = M // world
    for r in range(world):
        planes, a0, ghosts, ring = torch.load('{}.{}'.format(out, r))
        np.testing.assert_allclose(planes.numpy(), ref[:, r * Ml:(r + 1) * Ml], rtol=0, atol=1e-9)
        np.testing.assert_allclose(a0.numpy(), A[0][r * Ml:(r + 1) * Ml], rtol=0, atol=1e-9)
        assert np.all(np.diff(ghosts.numpy()) > 0) and ring.numpy().min() == 1 and ring.numpy().max() <= K - 1


def test_partition_plan_is_consistent_across_ranks():
    from deepsphere_tf1_b200.graph import ShardedGraph
    L = _graph(False)
    world = 4
    sgs = [ShardedGraph(L, r, world, None) for r in range(world)]
    for r, a in enumerate(sgs):
        assert a.Ml * world == L.shape[0] and a.Me == a.Ml + a.H
        for q, idx in a.send.items():
            off, n = sgs[q].recv[r]
            assert n == len(idx)
            # what r sends to q is exactly the ghost segment of q that r owns, in the same order
            np.testing.assert_array_equal(sgs[q].ghost_ids[off:off + n], idx + r * a.Ml)
        assert a.L_local.shape == (a.Me, a.Me) and a.L_local[a.Ml:].nnz == 0
    # wide halos: the same plan invariants, rows for the inner rings only
    sgs = [ShardedGraph(L, r, world, None, rings=3) for r in range(world)]
    for r, a in enumerate(sgs):
        for q, idx in a.send.items():
            off, n = sgs[q].recv[r]
            np.testing.assert_array_equal(sgs[q].ghost_ids[off:off + n], idx + r * a.Ml)
        rows_with_entries = np.unique(a.L_local[a.Ml:].tocoo().row)
        assert set(rows_with_entries) <= set(np.nonzero(a.ghost_ring < 3)[0])
    with pytest.raises(ValueError):
        ShardedGraph(L, 0, 5, None)


@pytest.mark.gpu
def test_sharded_model_matches_single_device_model():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs (gpurun --gpus 2)')
    n = min(4, torch.cuda.device_count())
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()),
           os.path.join(ROOT, 'tests', 'sharded_worker.py')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert 'SHARDED PARITY OK' in r.stdout, r.stdout[-4000:]


@pytest.mark.gpu
def test_peer_memory_allreduce_matches_nccl():
    """csrc/peer_allreduce.cu: eager calls and CUDA-graph replays against NCCL (tests/peer_worker.py under torchrun)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs (gpurun --gpus 2)')
    n = 4 if torch.cuda.device_count() >= 4 else 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()),
           os.path.join(ROOT, 'tests', 'peer_worker.py')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert 'PEER ALLREDUCE OK' in r.stdout, r.stdout[-4000:]
