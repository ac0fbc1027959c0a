"""GPU parity at the BENCHMARK's own sizes (cosmo config 2, batch 16): the persistent kernels loop thousands of times
per CTA there, which the miniature cases never exercise (a pipeline phase hazard of round 1 showed only on long loops).
Sparse products at M = 65,536 (layer 2) in every y / z form the step uses, the three contraction directions at
R = 1,048,576 rows (layer 2) and R = 262,144 (layer 3), the fused tile recurrence at layers 1-3.  Oracle = scipy CSR /
torch-CPU float64 on the same seeded inputs; tolerance 1e-4 of the largest expected magnitude (fp32 bar)."""
import numpy as np
import pytest
import torch
from scipy import sparse

from helpers import assert_close, healpix_L

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope='module')
def dev():
    from deepsphere_tf1_b200 import _lib
    _lib.LIB.load()
    return torch.device('cuda', 0)


@pytest.fixture(scope='module')
def layer2_graph(dev):
    from deepsphere_tf1_b200.graph import DeviceGraph
    L = healpix_L(512, order=2)                                  # the 65,536-vertex patch of cosmo layer 2
    assert L.shape[0] == 65536
    return L, DeviceGraph(L, dev)


def _apply(L64, x):
    """L x for x [B, M, F] in float64 on the CPU."""
    B, M, F = x.shape
    x0 = x.double().permute(1, 2, 0).reshape(M, F * B).numpy()
    return torch.from_numpy(np.asarray(L64 @ x0)).reshape(M, F, B).permute(2, 0, 1)


def test_sparse_product_forms_at_layer2_size(dev, layer2_graph):
    from deepsphere_tf1_b200 import ops
    L, g = layer2_graph
    L64 = sparse.csr_matrix(L, dtype=np.float64)
    B, F = 16, 16
    assert ops._pipe(g, 'fwd', B, g.M, F) is not None, 'the pipelined kernel must be the one that runs at this size'
    rs = np.random.RandomState(0)
    x, y, z = [torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32)) for _ in range(3)]
    xd, yd, zd = x.to(dev), y.to(dev), z.to(dev)
    Lx = _apply(L64, x)
    LTx = _apply(L64.T.tocsr(), x)
    assert_close(ops.spmm(g, xd), Lx, RTOL, 'X1 = L x')
    assert_close(ops.spmm(g, xd, yd, None, alpha=2.0, beta=-1.0), 2 * Lx - y.double(), RTOL, 'Xk = 2 L x - y')
    assert_close(ops.spmm(g, xd, yd, zd, alpha=0.5, beta=-1.5, gamma=0.25), 0.5 * Lx - 1.5 * y.double() + 0.25 * z.double(),
                 RTOL, 'general form')
    yt = yd.clone()
    ops.spmm(g, xd, yt, zd, out=yt, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True)
    assert_close(yt, 2 * LTx + y.double() - z.double(), RTOL, 'in-place adjoint form')
    # many launches back to back on the same operands: every launch must give the same bits (races show as noise)
    first = ops.spmm(g, xd, yd, None, alpha=2.0, beta=-1.0)
    for _ in range(50):
        assert torch.equal(ops.spmm(g, xd, yd, None, alpha=2.0, beta=-1.0), first)


def test_recurrence_and_adjoint_at_layer2_size(dev, layer2_graph):
    from deepsphere_tf1_b200 import ops
    L, g = layer2_graph
    L64 = sparse.csr_matrix(L, dtype=np.float64)
    B, F, K = 16, 16, 5
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32))
    planes = [x.double(), _apply(L64, x)]
    for k in range(2, K):
        planes.append(2 * _apply(L64, planes[-1]) - planes[-2])
    basis = ops.cheb_basis(g, x.to(dev), K)
    assert_close(basis, torch.stack(planes[1:]), RTOL, 'basis')
    # adjoint: <G, T(x)> = <T^T G, x> for the linear map x -> (T_0 x .. T_{K-1} x)
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32))
    LT = L64.T.tocsr()
    a2, a1 = None, G[K - 1].double()
    for k in range(K - 2, -1, -1):
        a = G[k].double() + (2.0 if k >= 1 else 1.0) * _apply(LT, a1) - (a2 if a2 is not None else 0)
        a2, a1 = a1, a
    dx = ops.cheb_basis_bwd(g, G.to(dev).clone())
    assert_close(dx, a1, RTOL, 'adjoint recurrence')


@pytest.mark.parametrize('shape', [(16, 65536, 16, 32, 5), (16, 16384, 32, 64, 5), (16, 4096, 64, 64, 5)])
def test_contraction_directions_at_benchmark_size(dev, shape):
    """Layer 2 (R = 1,048,576), layer 3 and layer 4 of the cosmo stack: forward with batch-norm sums, data gradient,
    weight gradient, against float64 on the CPU."""
    from deepsphere_tf1_b200 import ops
    B, M, Fin, Fout, K = shape
    R = B * M
    rs = np.random.RandomState(Fin + Fout)
    planes = torch.from_numpy(rs.standard_normal((K, B, M, Fin)).astype(np.float32))
    W = torch.from_numpy((rs.randn(Fin * K, Fout) / np.sqrt(Fin * K)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, M, Fout)).astype(np.float32))
    W3 = W.double().reshape(Fin, K, Fout)
    y_ref = torch.zeros(R, Fout, dtype=torch.float64)
    dW_ref = torch.zeros(Fin, K, Fout, dtype=torch.float64)
    dX_ref = torch.empty(K, R, Fin, dtype=torch.float64)
    dy2 = dy.reshape(R, Fout).double()
    for k in range(K):                                            # plane by plane: no [R, Fin*K] float64 temporary
        pk = planes[k].reshape(R, Fin).double()
        y_ref += pk @ W3[:, k, :]
        dW_ref[:, k, :] = pk.t() @ dy2
        dX_ref[k] = dy2 @ W3[:, k, :].t()
    x0 = planes[0].contiguous().to(dev)
    xk = planes[1:].contiguous().to(dev)
    Wd = W.to(dev)
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y = ops.contract_fwd(x0, xk, Wd, K, Fout, sums)
    assert_close(y.reshape(R, Fout), y_ref, RTOL, 'Y')
    assert_close(sums[:Fout], y_ref.sum(0), RTOL, 'sum y', floor=float(y_ref.abs().sum(0).max()))
    assert_close(sums[Fout:], (y_ref ** 2).sum(0), RTOL, 'sum y^2')
    G = ops.contract_bwd_data(dy.to(dev), Wd, K, Fin)
    assert_close(G.reshape(K, R, Fin), dX_ref, RTOL, 'dX_k')
    dW = torch.zeros(Fin * K, Fout, device=dev)
    ops.contract_bwd_weight(x0, xk, dy.to(dev), dW, K)
    assert_close(dW, dW_ref.reshape(Fin * K, Fout), RTOL, 'dW', floor=float(dW_ref.abs().max()))
    # determinism of the forward over repeated launches (pipeline races show as run-to-run differences)
    y2 = ops.contract_fwd(x0, xk, Wd, K, Fout, None)
    for _ in range(10):
        assert torch.equal(ops.contract_fwd(x0, xk, Wd, K, Fout, None), y2)


@pytest.mark.parametrize('level', [0, 1, 2])
def test_fused_tile_recurrence_at_benchmark_size(dev, level, monkeypatch):
    """csrc/cheb_tile.cu at cosmo layers 1 (sample-interleaved [1, 262144, 16]), 2 ([16, 65536, 16]) and 3 ([16, 16384, 32]):
    forward basis and adjoint against float64 on the CPU, and bit-stable over repeated launches."""
    from deepsphere_tf1_b200 import ops, utils
    from deepsphere_tf1_b200.graph import DeviceGraph
    nside = [1024, 512, 256][level]
    B, F = [(1, 16), (16, 16), (16, 32)][level]
    K = 5
    L = healpix_L(nside, order=2)
    ids = utils.nside2indexes([nside], 2)[0]
    g = DeviceGraph(L, dev, nested=(nside, ids))
    monkeypatch.setattr(ops, 'TILE_BWD_MAX_ITEMS', 1 << 30)
    monkeypatch.setattr(ops, 'TILE_FWD_MAX_ITEMS', 1 << 30)
    assert ops._tile(g, 'fwd', K, B, F) is not None, 'the fused tile kernel must be the one that runs at this size'
    L64 = sparse.csr_matrix(L, dtype=np.float64)
    rs = np.random.RandomState(level)
    x = torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32))
    planes = [x.double(), _apply(L64, x)]
    for k in range(2, K):
        planes.append(2 * _apply(L64, planes[-1]) - planes[-2])
    xd = x.to(dev)
    basis = ops.cheb_basis(g, xd, K)
    assert_close(basis, torch.stack(planes[1:]), RTOL, 'basis')
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32))
    LT = L64.T.tocsr()
    a2, a1 = None, G[K - 1].double()
    for k in range(K - 2, -1, -1):
        a = G[k].double() + (2.0 if k >= 1 else 1.0) * _apply(LT, a1) - (a2 if a2 is not None else 0)
        a2, a1 = a1, a
    Gd = G.to(dev)
    dx = ops.cheb_basis_bwd(g, Gd)
    assert_close(dx, a1, RTOL, 'adjoint recurrence')
    for _ in range(20):
        assert torch.equal(ops.cheb_basis(g, xd, K), basis)
        assert torch.equal(ops.cheb_basis_bwd(g, Gd), dx)
