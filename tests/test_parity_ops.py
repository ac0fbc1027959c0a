"""GPU parity: every libdsb200 kernel (called through the C ABI via ctypes) against the CPU
oracle on the same seeded inputs.  Tolerance (BASELINE.json north_star): fp32, max error
<= 1e-4 of the largest expected magnitude; index / arg-max results exact."""
import numpy as np
import pytest
import torch

from helpers import assert_close, healpix_L, random_knn_L, random_L

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope='module')
def dev():
    from deepsphere_tf1_b200 import _lib
    _lib.LIB.load()
    return torch.device('cuda', 0)


def _graph(L, dev, force_csr=False):
    from deepsphere_tf1_b200.graph import DeviceGraph
    return DeviceGraph(L, dev, force_csr=force_csr)


def _oracle_basis(L, x, K):
    """X_k planes [K, B, M, F] with the reference's recurrence (models.py:1049-1061)."""
    from oracle import ops as oops
    Lt = oops.sparse_to_torch(L)
    B, M, F = x.shape
    x0 = x.permute(1, 2, 0).reshape(M, F * B)
    planes = [x0]
    if K > 1:
        planes.append(torch.sparse.mm(Lt, x0))
    for k in range(2, K):
        planes.append(2 * torch.sparse.mm(Lt, planes[-1]) - planes[-2])
    return torch.stack([p.reshape(M, F, B).permute(2, 0, 1) for p in planes])


@pytest.mark.parametrize('F', [1, 2, 3, 4, 6, 16, 33, 64])
@pytest.mark.parametrize('kind', ['ell', 'csr'])
def test_cheb_basis_matches_oracle(dev, F, kind):
    from deepsphere_tf1_b200 import ops
    L = healpix_L(8) if kind == 'ell' else random_knn_L(500, 7)
    g = _graph(L, dev, force_csr=(kind == 'csr'))
    assert g.is_ell == (kind == 'ell')
    rs = np.random.RandomState(F)
    x = torch.from_numpy(rs.randn(3, L.shape[0], F).astype(np.float32))
    K = 5
    basis = ops.cheb_basis(g, x.to(dev), K)
    assert_close(basis, _oracle_basis(L, x, K)[1:], RTOL, 'basis')


@pytest.mark.parametrize('F', [3, 16])
def test_monomial_basis_and_its_adjoint(dev, F):
    """ops.cheb_basis(monomial=True): X_k = L X_(k-1) (models.py:1094-1103); its adjoint against the dense transpose."""
    from deepsphere_tf1_b200 import ops
    L = healpix_L(8)
    g = _graph(L, dev)
    K = 4
    rs = np.random.RandomState(F)
    x = torch.from_numpy(rs.randn(3, L.shape[0], F).astype(np.float32))
    Ld = torch.from_numpy(L.toarray()).double()
    planes = [x.double()]
    for k in range(1, K):
        planes.append(torch.einsum('ij,bjf->bif', Ld, planes[-1]))
    basis = ops.cheb_basis(g, x.to(dev), K, monomial=True)
    assert_close(basis, torch.stack(planes[1:]), RTOL, 'monomial basis')
    G = torch.from_numpy(rs.randn(K, 3, L.shape[0], F).astype(np.float32))
    A = G[K - 1].double()
    for k in range(K - 2, -1, -1):
        A = G[k].double() + torch.einsum('ji,bjf->bif', Ld, A)
    assert_close(ops.cheb_basis_bwd(g, G.to(dev).clone(), monomial=True), A, RTOL, 'monomial adjoint')


def test_spmm_general_form_and_aliasing(dev):
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    L = random_L(300, 5)
    g = _graph(L, dev)
    rs = np.random.RandomState(0)
    x, y, z = [torch.from_numpy(rs.randn(2, 300, 8).astype(np.float32)) for _ in range(3)]
    Ld = torch.from_numpy(L.toarray())
    LT = Ld.t()
    exp = 0.5 * torch.einsum('ij,bjf->bif', Ld, x) - 1.5 * y + 0.25 * z
    out = ops.spmm(g, x.to(dev), y.to(dev), z.to(dev), alpha=0.5, beta=-1.5, gamma=0.25)
    assert_close(out, exp, RTOL, 'spmm')
    # transposed matrix, result written over y
    yd = y.to(dev).clone()
    ops.spmm(g, x.to(dev), yd, None, out=yd, alpha=2.0, beta=1.0, transpose=True)
    assert_close(yd, 2 * torch.einsum('ij,bjf->bif', LT, x) + y, RTOL, 'spmm^T in place')
    assert oops is not None


@pytest.mark.parametrize('sb', [1, 2, 4])
@pytest.mark.parametrize('F', [4, 16, 32, 64, 132, 260])
@pytest.mark.parametrize('graph', ['healpix', 'knn', 'nonsym'])
def test_packed_record_spmm_is_bit_identical_to_row_major(dev, F, graph, sb, monkeypatch):
    """Packed slot-major ELL records (spmm_rec_kernel, the default path) vs the row-major col[] / val[] kernel: same
    products in the same order, so the results are bit-identical; dsb200_ell_pack reproduces the host packing."""
    from deepsphere_tf1_b200 import ops
    from deepsphere_tf1_b200._lib import call
    L = {'healpix': lambda: healpix_L(16), 'knn': lambda: random_knn_L(700, 6), 'nonsym': lambda: random_L(1500, 9)}[graph]()
    g = _graph(L, dev)
    assert g.is_ell and g.fwd_rec is not None
    if graph == 'healpix':
        assert g.fwd[3] == 9
    _, col, val, width = g.fwd
    rec = torch.empty_like(g.fwd_rec)
    call('ell_pack', col, val, width, g.M, rec)
    assert torch.equal(rec, g.fwd_rec)
    monkeypatch.setenv('DSB200_SPMM_SB', str(sb))
    B, K = 5, 4
    rs = np.random.RandomState(F + sb)
    x, y, z = [torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32)).to(dev) for _ in range(3)]
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32)).to(dev)
    try:
        ops.set_records(True)
        out = ops.spmm(g, x, y, z, alpha=0.5, beta=-1.5, gamma=0.25)
        outT = ops.spmm(g, x, y.clone(), None, alpha=2.0, beta=1.0, transpose=True)
        basis = ops.cheb_basis(g, x, K)
        dx = ops.cheb_basis_bwd(g, G.clone())
        ops.set_records(False)
        assert torch.equal(out, ops.spmm(g, x, y, z, alpha=0.5, beta=-1.5, gamma=0.25))
        assert torch.equal(outT, ops.spmm(g, x, y.clone(), None, alpha=2.0, beta=1.0, transpose=True))
        assert torch.equal(basis, ops.cheb_basis(g, x, K))
        assert torch.equal(dx, ops.cheb_basis_bwd(g, G.clone()))
    finally:
        ops.set_records(True)
    assert_close(basis, _oracle_basis(L, x.cpu(), K)[1:], RTOL, 'basis vs oracle')


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('dense', [False, True])
@pytest.mark.parametrize('graph', ['patch64', 'ragged'])
def test_pipelined_spmm_walking_memory_backwards_gives_the_same_bits(dev, graph, dense, bf16, monkeypatch):
    """`reverse=True` (bit 1 of the `dense` argument of dsb200_spmm_pipe; the zigzag recurrence of ops.ZIGZAG) only changes the
    order in which the (row tile, sample) items are taken: every form of the operator, the in-place adjoint form included,
    and the whole recurrence / adjoint under both zigzag parities are bit-identical to the forward walk."""
    from scipy import sparse
    from deepsphere_tf1_b200 import ops
    L = {'patch64': lambda: healpix_L(64, order=2),
         'ragged': lambda: sparse.coo_matrix(sparse.csr_matrix(healpix_L(16))[:3001, :3001])}[graph]()
    g = _graph(L, dev)
    monkeypatch.setattr(ops, 'PIPE_MIN_ITEMS', 1)
    monkeypatch.setattr(ops, 'USE_DENSE_GROUPS', dense)
    monkeypatch.setattr(ops, 'USE_TILE', False)
    B, F, K = 7, 32, 5
    if ops._pipe(g, 'fwd', B, g.M, F, esize=2 if bf16 else 4, min_items=1) is None:
        pytest.skip('no tile plan for this shape')
    rs = np.random.RandomState(3)
    dt = torch.bfloat16 if bf16 else torch.float32
    x, y, z = [torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32)).to(dev).to(dt) for _ in range(3)]
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32)).to(dev).to(dt)
    for kw in (dict(), dict(y=y, alpha=2.0, beta=-1.0), dict(y=y, z=z, alpha=0.5, beta=-1.5, gamma=0.25)):
        assert torch.equal(ops.spmm(g, x, **kw), ops.spmm(g, x, reverse=True, **kw)), sorted(kw)
    a, b = y.clone(), y.clone()
    ops.spmm(g, x, a, z, out=a, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True)
    ops.spmm(g, x, b, z, out=b, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True, reverse=True)
    assert torch.equal(a, b)
    monkeypatch.setattr(ops, 'ZIGZAG', 0)
    basis, dx = ops.cheb_basis(g, x, K).clone(), ops.cheb_basis_bwd(g, G.clone()).clone()
    for z_ in (1, 2):
        monkeypatch.setattr(ops, 'ZIGZAG', z_)
        assert ops._rev_fwd(K - 1, K) == (z_ == 1) and ops._rev_fwd(K - 2, K) == (z_ == 2)
        assert torch.equal(ops.cheb_basis(g, x, K), basis) and torch.equal(ops.cheb_basis_bwd(g, G.clone()), dx)


@pytest.mark.parametrize('dense', [False, True])
@pytest.mark.parametrize('F', [16, 32, 64])
@pytest.mark.parametrize('B', [1, 5, 16])
@pytest.mark.parametrize('graph', ['healpix16', 'patch64', 'ragged'])
def test_pipelined_spmm_equals_the_gather_kernel(dev, F, B, graph, dense, monkeypatch):
    """Bulk-asynchronous pipelined sparse product (csrc/sparse_pipe.cu, the default where a tile plan exists) vs the
    gather kernel -- general form, in-place adjoint form, whole recurrence and its adjoint.  Record form: identical
    arithmetic order, identical bits.  Dense-group form (4 rows x their <= 16 distinct columns): sums in column order,
    equal to rounding.  Both within the fp32 bar of the oracle."""
    from scipy import sparse
    from deepsphere_tf1_b200 import ops
    L = {'healpix16': lambda: healpix_L(16), 'patch64': lambda: healpix_L(64, order=2),
         'ragged': lambda: sparse.coo_matrix(sparse.csr_matrix(healpix_L(16))[:3001, :3001])}[graph]()
    g = _graph(L, dev)
    monkeypatch.setattr(ops, 'PIPE_MIN_ITEMS', 1)
    plan = ops._pipe(g, 'fwd', B, g.M, F)
    assert plan is not None and plan.dense is not None

    def same(a, b, name):
        if dense:
            assert_close(a, b, 2e-6, name)
        else:
            assert torch.equal(a, b), name

    K = 4
    rs = np.random.RandomState(F + B)
    x, y, z = [torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32)).to(dev) for _ in range(3)]
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32)).to(dev)
    try:
        ops.set_pipe(True, dense=dense)
        out = ops.spmm(g, x, y, z, alpha=0.5, beta=-1.5, gamma=0.25)
        out1 = ops.spmm(g, x)
        yt = y.clone()
        ops.spmm(g, x, yt, z, out=yt, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True)
        basis = ops.cheb_basis(g, x, K)
        dx = ops.cheb_basis_bwd(g, G.clone())
        ops.set_pipe(False)
        same(out, ops.spmm(g, x, y, z, alpha=0.5, beta=-1.5, gamma=0.25), 'general form')
        same(out1, ops.spmm(g, x), 'plain product')
        yr = y.clone()
        ops.spmm(g, x, yr, z, out=yr, alpha=2.0, beta=1.0, gamma=-1.0, transpose=True)
        same(yt, yr, 'in-place adjoint form')
        same(basis, ops.cheb_basis(g, x, K), 'basis')
        same(dx, ops.cheb_basis_bwd(g, G.clone()), 'adjoint recurrence')
    finally:
        ops.set_pipe(True, dense=True)
    assert_close(basis, _oracle_basis(L, x.cpu(), K)[1:], RTOL, 'basis vs oracle')


@pytest.mark.parametrize('tensor_cores', [True, False])
@pytest.mark.parametrize('shape', [(2, 192, 1, 16, 5), (2, 192, 16, 32, 5), (3, 100, 6, 16, 4), (1, 77, 3, 5, 3),
                                   (2, 64, 64, 64, 5), (2, 48, 32, 2, 5), (4, 33, 7, 9, 1), (2, 50, 96, 40, 4),
                                   (3, 1000, 16, 32, 5), (2, 700, 32, 64, 5), (2, 300, 64, 64, 5), (2, 300, 8, 16, 3),
                                   (2, 300, 24, 40, 2), (1, 4100, 64, 8, 5), (5, 1111, 16, 16, 1), (2, 333, 128, 128, 2),
                                   (1, 20000, 32, 64, 5)])
def test_contraction_forward_backward(dev, shape, tensor_cores):
    """Y = Xcat W with the reference's row order fin*K + k (models.py:1062-1078) and both
    gradients of tf.matmul, plus the batch-norm sums of the epilogue -- on the tcgen05 path
    (3xTF32, where the shape allows) and on the FP32 FFMA path."""
    from deepsphere_tf1_b200 import ops
    prev = ops.set_tensor_cores(tensor_cores)
    try:
        _contraction_case(dev, shape)
    finally:
        ops.set_tensor_cores(prev)


def _contraction_case(dev, shape):
    from deepsphere_tf1_b200 import ops
    B, M, Fin, Fout, K = shape
    rs = np.random.RandomState(Fin * 100 + Fout)
    planes = torch.from_numpy(rs.randn(K, B, M, Fin).astype(np.float32))
    W = torch.from_numpy((rs.randn(Fin * K, Fout) / np.sqrt(Fin * K)).astype(np.float32))
    dy = torch.from_numpy(rs.randn(B, M, Fout).astype(np.float32))
    xcat = planes.permute(1, 2, 3, 0).reshape(B * M, Fin * K).double()        # N x M x Fin x K (1063-1064)
    y_ref = (xcat @ W.double()).reshape(B, M, Fout)
    dW_ref = xcat.t() @ dy.reshape(B * M, Fout).double()
    dX_ref = (dy.reshape(B * M, Fout).double() @ W.double().t()).reshape(B, M, Fin, K).permute(3, 0, 1, 2)

    x0 = planes[0].contiguous().to(dev)
    xk = planes[1:].contiguous().to(dev) if K > 1 else None
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y = ops.contract_fwd(x0, xk, W.to(dev), K, Fout, sums)
    assert_close(y, y_ref, RTOL, 'Y')
    assert_close(sums[:Fout], y_ref.sum((0, 1)), RTOL, 'sum y')
    assert_close(sums[Fout:], (y_ref ** 2).sum((0, 1)), RTOL, 'sum y^2')
    y2 = ops.contract_fwd(x0, xk, W.to(dev), K, Fout, None)
    assert_close(y2, y, 1e-6, 'Y without the batch-norm epilogue')
    G = ops.contract_bwd_data(dy.to(dev), W.to(dev), K, Fin)
    assert_close(G, dX_ref, RTOL, 'dX_k')
    dW = torch.zeros(Fin * K, Fout, device=dev)
    ops.contract_bwd_weight(x0, xk, dy.to(dev), dW, K)
    assert_close(dW, dW_ref, RTOL, 'dW')


@pytest.mark.parametrize('kind', ['ell', 'csr', 'nonsym'])
def test_chebconv_gradients_match_oracle_autograd(dev, kind):
    """Whole filter (recurrence + contraction), forward and both gradients, vs torch-CPU autograd
    of oracle.ops.chebyshev5."""
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    L = {'ell': lambda: healpix_L(4), 'csr': lambda: random_knn_L(200, 6), 'nonsym': lambda: random_L(150, 4)}[kind]()
    g = _graph(L, dev, force_csr=(kind == 'csr'))
    M = L.shape[0]
    B, Fin, Fout, K = 3, 5, 7, 4
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.randn(B, M, Fin).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy((rs.randn(Fin * K, Fout) * 0.3).astype(np.float32)).requires_grad_(True)
    dy = torch.from_numpy(rs.randn(B, M, Fout).astype(np.float32))
    y_ref = oops.chebyshev5(x, oops.sparse_to_torch(L), W, K)
    y_ref.backward(dy)

    xd, Wd = x.detach().to(dev), W.detach().to(dev)
    basis = ops.cheb_basis(g, xd, K)
    y = ops.contract_fwd(xd, basis, Wd, K, Fout)
    assert_close(y, y_ref, RTOL, 'y')
    dW = torch.zeros_like(Wd)
    ops.contract_bwd_weight(xd, basis, dy.to(dev), dW, K)
    assert_close(dW, W.grad, RTOL, 'dW')
    G = ops.contract_bwd_data(dy.to(dev), Wd, K, Fin)
    dx = ops.cheb_basis_bwd(g, G)
    assert_close(dx, x.grad, RTOL, 'dx')


@pytest.mark.parametrize('F', [3, 16, 64])
@pytest.mark.parametrize('p,pool', [(1, 'max'), (4, 'max'), (4, 'average'), (16, 'max')])
@pytest.mark.parametrize('act', ['relu', 'elu', 'identity'])
@pytest.mark.parametrize('bn', [True, False])
def test_bn_bias_act_pool_chain(dev, F, p, pool, act, bn):
    """normalise -> bias -> activation -> pool (models.py:1235-1243) forward and backward,
    including the batch-norm backward through the batch statistics."""
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    B, M = 3, 64
    rs = np.random.RandomState(F + p)
    y = torch.from_numpy((rs.randn(B, M, F) * 2 + 0.5).astype(np.float32)).requires_grad_(True)
    b = torch.from_numpy(rs.randn(F).astype(np.float32) * 0.3).requires_grad_(True)
    z = y
    if bn:
        z, nm, nv = oops.batch_normalization(z, torch.zeros(F), torch.ones(F), True)
    z = oops.bias(z, b)
    z = oops.ACTIVATIONS[act](z) if act != 'identity' else z
    out_ref = oops.pool(z, p, pool, 'healpix')
    dout = torch.from_numpy(rs.randn(*out_ref.shape).astype(np.float32))
    out_ref.backward(dout)

    yd, bd = y.detach().to(dev), b.detach().to(dev)
    mean = rstd = None
    count = float(B * M)
    if bn:
        # statistics exactly as the contraction epilogue + finalize produce them
        sums = torch.cat([yd.double().sum((0, 1)), (yd.double() ** 2).sum((0, 1))])
        mean, rstd = torch.empty(F, device=dev), torch.empty(F, device=dev)
        mm, mv = torch.zeros(F, device=dev), torch.ones(F, device=dev)
        ops.bn_finalize(sums, count, F, mean, rstd, mm, mv)
        assert_close(mm, nm, RTOL, 'moving mean')
        assert_close(mv, nv, RTOL, 'moving variance')
    out = ops.bn_act_pool_fwd(yd, mean, rstd, bd, p, pool, act)
    assert_close(out, out_ref, RTOL, 'out')
    gz, s = ops.bn_act_pool_bwd(yd, mean, rstd, bd, dout.to(dev), p, pool, act)
    assert_close(s[:F], b.grad, 2 * RTOL, 'dbias')
    if bn:
        ops.bn_bwd_apply(yd, mean, rstd, s, count, gz)
    assert_close(gz, y.grad, 2 * RTOL, 'dy')
    if bn:
        # two-pass variant used by the model: sums only, then dy straight from y and dout
        none, s2 = ops.bn_act_pool_bwd(yd, mean, rstd, bd, dout.to(dev), p, pool, act, want_gz=False)
        assert none is None
        assert_close(s2, s, 1e-6, 'sums (pass 1)')
        dy = ops.bn_act_pool_bwd_dy(yd, mean, rstd, bd, dout.to(dev), s2, count, p, pool, act)
        assert_close(dy, y.grad, 2 * RTOL, 'dy (pass 2)')
        if ops.bn_relu_pool_sums_supported(F, p, pool, act):
            # lean pass 1: the same sums from the pooled output and its gradient only
            s3 = ops.bn_relu_pool_sums(out, dout.to(dev), bd)
            assert_close(s3, s, 1e-5, 'sums (pass 1, pooled tensors)')


@pytest.mark.parametrize('kind', ['mean', 'var', 'meanvar', 'max'])
def test_statistics_head(dev, kind):
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.randn(4, 100, 6).astype(np.float32)).requires_grad_(True)
    ref = oops.statistics(x, kind, False)
    d = torch.from_numpy(rs.randn(*ref.shape).astype(np.float32))
    ref.backward(d)
    out = ops.stat_fwd(x.detach().to(dev), kind)
    assert_close(out, ref, RTOL, 'stat')
    dx = ops.stat_bwd(x.detach().to(dev), d.to(dev), kind)
    assert_close(dx, x.grad, RTOL, 'dstat')


def test_resize_unpool_concat(dev):
    from deepsphere_tf1_b200 import ops
    from deepsphere_tf1_b200._consts import UNPOOL_REPEAT, UNPOOL_ZEROFILL
    from oracle import ops as oops
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.randn(2, 12, 5).astype(np.float32))
    xd = x.to(dev)
    assert torch.equal(ops.vertex_resize(xd, 42, 1.0).cpu(), oops.unpool(x, 42, 'average', 'icosahedron'))
    assert torch.equal(ops.vertex_resize(xd, 7, 0.0).cpu(), oops.pool(x, 7, 'average', 'icosahedron'))
    assert torch.equal(ops.unpool_fwd(xd, 4, UNPOOL_REPEAT).cpu(), oops.unpool(x, 4, 'average', 'healpix'))
    assert torch.equal(ops.unpool_fwd(xd, 4, UNPOOL_ZEROFILL).cpu(), oops.unpool(x, 4, 'max', 'healpix'))
    d = torch.from_numpy(rs.randn(2, 48, 5).astype(np.float32))
    assert_close(ops.unpool_bwd(d.to(dev), 4, UNPOOL_REPEAT), d.reshape(2, 12, 4, 5).sum(2), 1e-6, 'unpool bwd')
    assert torch.equal(ops.unpool_bwd(d.to(dev), 4, UNPOOL_ZEROFILL).cpu(), d.reshape(2, 12, 4, 5)[:, :, 0])
    b = torch.from_numpy(rs.randn(2, 12, 3).astype(np.float32))
    cat = ops.concat_f(xd, b.to(dev))
    assert torch.equal(cat.cpu(), torch.cat([x, b], -1))
    da, db = ops.split_f(cat, 5)
    assert torch.equal(da.cpu(), x) and torch.equal(db.cpu(), b)
    da, db2 = ops.split_f(cat, 5, db)
    assert torch.equal(db2.cpu(), 2 * b)


def test_fc_losses_regulariser(dev):
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.randn(8, 20).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy(rs.randn(20, 7).astype(np.float32) * 0.2).requires_grad_(True)
    b = torch.from_numpy(rs.randn(7).astype(np.float32)).requires_grad_(True)
    labels = torch.from_numpy(rs.randint(0, 7, size=8).astype(np.int32))
    h = torch.relu(oops.fc(x, W, b))
    total, data_loss = oops.loss(h, labels, False, [(W, 0.5)], 0.3)
    total.backward()
    hd = ops.fc_fwd(x.detach().to(dev), W.detach().to(dev), b.detach().to(dev), 'relu')
    assert_close(hd, h, RTOL, 'fc')
    loss = torch.zeros(1, device=dev)
    dl = ops.softmax_ce(hd, labels.to(dev), loss)
    assert_close(loss, data_loss.reshape(1), RTOL, 'ce')
    dW, db = torch.zeros(20, 7, device=dev), torch.zeros(7, device=dev)
    dx = ops.fc_bwd(x.detach().to(dev), W.detach().to(dev), hd, dl, dW, db, 'relu')
    coef = torch.full((140,), 0.3 / 0.25 / 140, device=dev)
    ops.l2_reg(W.detach().to(dev).reshape(-1), coef, dW.reshape(-1), loss)
    assert_close(loss, total.reshape(1), RTOL, 'total loss')
    assert_close(dx, x.grad, RTOL, 'dx')
    assert_close(dW, W.grad, RTOL, 'dW')
    assert_close(db, b.grad, RTOL, 'db')
    probs, pred = ops.softmax_argmax(hd, True)
    assert_close(probs, torch.softmax(h.detach(), -1), RTOL, 'probs')
    assert torch.equal(pred.cpu().long(), h.detach().argmax(-1))
    # MSE (tf.losses.mean_squared_error, models.py:788)
    t = torch.from_numpy(rs.randn(8, 7).astype(np.float32))
    loss.zero_()
    dp = ops.mse(hd.reshape(-1), t.to(dev).reshape(-1), loss)
    assert_close(loss, ((h.detach() - t) ** 2).mean().reshape(1), RTOL, 'mse')
    assert_close(dp.reshape(8, 7), 2 * (h.detach() - t) / 56, RTOL, 'dmse')


@pytest.mark.parametrize('shape', [(3, 50, 7), (2, 256, 64), (1, 9, 40)])
def test_learned_histogram_forward_backward(dev, shape):
    """dsb200_hist_fwd / bwd vs torch autograd of the restated cgcnn.learned_histogram (models.py:1166-1192)."""
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    B, M, F = shape
    rs = np.random.RandomState(M)
    x = torch.from_numpy((rs.randn(B, M, F) * 1.5).astype(np.float32)).requires_grad_(True)
    c0, w0 = oops.histogram_init(F)
    c = (c0 + torch.from_numpy(rs.randn(1, 1, F, 20).astype(np.float32)) * 0.05).requires_grad_(True)
    w = (w0 * torch.from_numpy(rs.choice([-1.0, 1.0], size=(1, 1, F, 20)).astype(np.float32))
         * (1 + 0.1 * torch.from_numpy(rs.randn(1, 1, F, 20).astype(np.float32)))).requires_grad_(True)
    ref = oops.learned_histogram(x, c, w)
    dh = torch.from_numpy(rs.randn(B, F * 20).astype(np.float32))
    ref.backward(dh)
    cd, wd = c.detach().reshape(F, 20).to(dev).contiguous(), w.detach().reshape(F, 20).to(dev).contiguous()
    out = ops.hist_fwd(x.detach().to(dev), cd, wd)
    assert_close(out, ref.detach(), RTOL, 'histogram')
    dc, dw = torch.zeros_like(cd), torch.zeros_like(wd)
    dx = ops.hist_bwd(x.detach().to(dev), cd, wd, dh.to(dev), dc, dw)
    assert_close(dx, x.grad, RTOL, 'dx')
    assert_close(dc, c.grad.reshape(F, 20), RTOL, 'dcenters')
    assert_close(dw, w.grad.reshape(F, 20), RTOL, 'dwidths')


def test_weighted_cross_entropy(dev):
    from deepsphere_tf1_b200 import ops
    from oracle import ops as oops
    rs = np.random.RandomState(0)
    logits = torch.from_numpy(rs.randn(2, 50, 3).astype(np.float32)).requires_grad_(True)
    labels = torch.from_numpy(rs.randint(0, 3, size=(2, 50)).astype(np.int32))
    ref, _ = oops.loss(logits, labels, False, [(torch.zeros(1), 1.0)], 0.0, weighted=True)
    ref.backward()
    loss = torch.zeros(1, device=dev)
    cw = torch.tensor(oops.CLASS_WEIGHTS, dtype=torch.float32, device=dev)
    d = ops.softmax_ce(logits.detach().to(dev), labels.to(dev), loss, 1.0, True, cw)
    assert_close(loss, ref.detach().reshape(1), RTOL, 'weighted CE')
    assert_close(d, logits.grad, RTOL, 'weighted CE gradient')


@pytest.mark.parametrize('name', ['Adam', 'Momentum', 'SGD', 'RMSProp'])
def test_optimizers_follow_tf1_update_rules(dev, name):
    from deepsphere_tf1_b200 import ops, optimizers
    from oracle import optim_tf1
    rs = np.random.RandomState(7)
    w = torch.from_numpy(rs.randn(1000).astype(np.float32))
    prod = {'Adam': optimizers.AdamOptimizer(1e-2, epsilon=0.1), 'Momentum': optimizers.MomentumOptimizer(1e-2, 0.9),
            'SGD': optimizers.GradientDescentOptimizer(1e-2), 'RMSProp': optimizers.RMSPropOptimizer(1e-2, momentum=0.5)}[name]
    orc = {'Adam': optim_tf1.Adam(epsilon=0.1), 'Momentum': optim_tf1.Momentum(0.9), 'SGD': optim_tf1.SGD(),
           'RMSProp': optim_tf1.RMSProp(momentum=0.5)}[name]
    wd = w.to(dev).clone()
    slots = [torch.full_like(wd, prod.slot1_init if i == 0 else 0.0) for i in range(prod.slots)]
    params = {'w': w.clone()}
    for t in range(1, 6):
        g = torch.from_numpy(rs.randn(1000).astype(np.float32))
        lr = 1e-2 * 0.9 ** t
        params = orc.step(params, {'w': g}, lr)
        hyper = torch.tensor(prod.hyper(lr, t), dtype=torch.float32, device=dev)
        ops.optimizer_step(prod.kind, wd, g.to(dev), slots[0] if slots else None, slots[1] if len(slots) > 1 else None, hyper)
    assert_close(wd, params['w'], RTOL, name)


@pytest.mark.parametrize('monomial', [False, True])
@pytest.mark.parametrize('K', [2, 4, 5])
@pytest.mark.parametrize('B,F', [(1, 16), (3, 32), (16, 16), (2, 64)])
@pytest.mark.parametrize('graph', [(64, 2), (128, 2), (64, 1)])
def test_fused_tile_recurrence(dev, graph, B, F, K, monomial, monkeypatch):
    """csrc/cheb_tile.cu (all K - 1 sparse products of a layer in one launch, 16 x 16 pixel tiles with a (K - 1)-ring halo in
    shared memory): forward basis and adjoint against the per-step kernels and the oracle recurrence."""
    from deepsphere_tf1_b200 import ops, utils
    from deepsphere_tf1_b200.graph import DeviceGraph
    nside, order = graph
    L = healpix_L(nside, order)
    ids = utils.nside2indexes([nside], order)[0]
    g = DeviceGraph(L, dev, nested=(nside, ids))
    monkeypatch.setattr(ops, 'TILE_MIN_ITEMS', 1)
    monkeypatch.setattr(ops, 'TILE_BWD_MAX_ITEMS', 1 << 30)
    monkeypatch.setattr(ops, 'TILE_FWD_MAX_ITEMS', 1 << 30)
    assert ops._tile(g, 'fwd', K, B, F) is not None and ops._tile(g, 'bwd', K, B, F) is not None
    rs = np.random.RandomState(K + F)
    x = torch.from_numpy(rs.randn(B, g.M, F).astype(np.float32)).to(dev)
    G = torch.from_numpy(rs.randn(K, B, g.M, F).astype(np.float32)).to(dev)
    basis = ops.cheb_basis(g, x, K, monomial)
    dx = ops.cheb_basis_bwd(g, G.clone(), monomial)
    try:
        ops.set_tile(False)
        basis_ref = ops.cheb_basis(g, x, K, monomial)
        dx_ref = ops.cheb_basis_bwd(g, G.clone(), monomial)
    finally:
        ops.set_tile(True)
    assert_close(basis, basis_ref, 2e-6, 'basis vs per-step kernels')
    assert_close(dx, dx_ref, 2e-6, 'adjoint vs per-step kernels')
    if not monomial:
        assert_close(basis, _oracle_basis(L, x.cpu(), K)[1:], RTOL, 'basis vs oracle')
    # the adjoint identity <G, T x> = <T^T G, x> over all K planes, in float64
    planes = torch.cat([x[None], basis]).double()
    lhs = float((G.double() * planes).sum())
    rhs = float((dx.double() * x.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), float(planes.abs().max()) * float(G.abs().max())), (lhs, rhs)


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('B,M,Fin,K,Fout', [(2, 1024, 16, 5, 32), (3, 512, 32, 5, 64), (2, 256, 64, 5, 64), (1, 4096, 16, 3, 32),
                                            (2, 1028, 32, 4, 32), (16, 65536, 16, 5, 32)])
def test_contraction_with_the_max_pool_fused_into_the_epilogue(dev, B, M, Fin, K, Fout, bf16):
    """dsb200_cheb_contract_fwd_pool4(_bf16): Y, the batch-norm sums and the max over every 4 consecutive vertices in one launch
     -- bit-identical to the contraction followed by the pooling, with and without the un-pooled output."""
    from deepsphere_tf1_b200 import ops
    gen = torch.Generator().manual_seed(M + Fin)
    dt = torch.bfloat16 if bf16 else torch.float32
    x = torch.randn(K, B, M, Fin, generator=gen).to(dev).to(dt)
    W = (torch.randn(Fin * K, Fout, generator=gen) * 0.2).to(dev)
    x0, xk = x[0], x[1:]
    if not ops.contract_pool4_supported(x0, K, Fout):
        pytest.skip('no tensor-core kernel with a pooled epilogue for this shape')
    s_ref = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y_ref = ops.contract_fwd(x0, xk, W, K, Fout, s_ref)
    p_ref = y_ref.view(B, M // 4, 4, Fout).max(dim=2).values
    s1 = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y, yp = ops.contract_fwd_pool4(x0, xk, W, K, Fout, s1, want_y=True)
    assert torch.equal(y, y_ref) and torch.equal(yp, p_ref)
    assert_close(s1, s_ref, 1e-6, 'sums')                             # float partial sums, added in launch-dependent order
    y2, yp2 = ops.contract_fwd_pool4(x0, xk, W, K, Fout, None, want_y=False)
    assert y2 is None and torch.equal(yp2, p_ref)
    # pooling first == pooling last through batch norm + bias + relu (what models.py does with the pooled tensor)
    mean = (torch.randn(Fout, generator=gen) * 0.1).to(dev)
    rstd = (torch.rand(Fout, generator=gen) + 0.5).to(dev)
    bias = (torch.randn(Fout, generator=gen) * 0.1).to(dev)
    assert torch.equal(ops.bn_act_pool_fwd(yp, mean, rstd, bias, 1, 'max', 'relu'), ops.bn_act_pool_fwd(y_ref, mean, rstd, bias, 4, 'max', 'relu'))
