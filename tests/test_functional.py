"""torch.autograd / torch.library face of the kernels (deepsphere_tf1_b200/functional.py) against the oracle's autograd."""
import numpy as np
import pytest
import torch

from helpers import assert_close, healpix_L, random_knn_L


def test_registered_operators_have_no_cpu_kernel():
    from deepsphere_tf1_b200 import functional
    functional.register_ops()
    assert hasattr(torch.ops.dsb200, 'spmm_ell') and hasattr(torch.ops.dsb200, 'cheb_contract_fwd')
    x = torch.zeros(1, 4, 2)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.dsb200.spmm_ell(torch.zeros(4, 1, dtype=torch.int32), torch.zeros(4, 1), x, None, None, 1.0, 0.0, 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['healpix', 'knn'])
def test_chebconv_module_gradients_match_oracle_autograd(dev, kind):
    from deepsphere_tf1_b200 import functional
    from oracle import ops as oops
    L = healpix_L(8) if kind == 'healpix' else random_knn_L(400, 7)
    Fin, Fout, K, B = 6, 10, 4, 3
    torch.manual_seed(0)
    conv = functional.ChebConv(L, Fin, Fout, K, device=dev)
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.randn(B, L.shape[0], Fin).astype(np.float32))
    t = torch.from_numpy(rs.randn(B, L.shape[0] // 4, Fout).astype(np.float32))
    xd = x.to(dev).requires_grad_(True)
    y = functional.pool(torch.relu(conv(xd)), 4, 'max')
    loss = ((y - t.to(dev)) ** 2).mean()
    loss.backward()
    # oracle: the same composition on the CPU
    xo = x.clone().requires_grad_(True)
    Wo = conv.weight.detach().cpu().clone().requires_grad_(True)
    yo = oops.pool(torch.relu(oops.chebyshev5(xo, oops.sparse_to_torch(L), Wo, K)), 4, 'max', 'healpix')
    lo = ((yo - t) ** 2).mean()
    lo.backward()
    assert float(loss) == pytest.approx(float(lo), rel=1e-5)
    assert_close(y, yo, 1e-4, 'y')
    assert_close(conv.weight.grad, Wo.grad, 1e-4, 'dW')
    assert_close(xd.grad, xo.grad, 1e-4, 'dx')


@pytest.mark.gpu
def test_registered_operators_equal_the_wrappers(dev):
    from deepsphere_tf1_b200 import functional, ops
    from deepsphere_tf1_b200.graph import DeviceGraph
    functional.register_ops()
    L = healpix_L(4)
    g = DeviceGraph(L, dev)
    _, col, val, width = g.fwd
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.randn(2, g.M, 8).astype(np.float32)).to(dev)
    y = torch.from_numpy(rs.randn(2, g.M, 8).astype(np.float32)).to(dev)
    out = torch.ops.dsb200.spmm_ell(col, val, x, y, None, 2.0, -1.0, 0.0)
    ops.set_records(False)
    try:
        assert torch.equal(out, ops.spmm(g, x, y, None, alpha=2.0, beta=-1.0))
    finally:
        ops.set_records(True)
    K = 3
    basis = torch.ops.dsb200.cheb_basis_ell(col, val, x, K)
    assert_close(basis, ops.cheb_basis(g, x, K), 1e-6, 'basis')
    W = torch.from_numpy(rs.randn(8 * K, 5).astype(np.float32)).to(dev)
    yy = torch.ops.dsb200.cheb_contract_fwd(x, basis, W, K)
    assert_close(yy, ops.contract_fwd(x, basis, W, K, 5), 1e-6, 'Y')
    dy = torch.randn_like(yy)
    assert torch.ops.dsb200.cheb_contract_bwd_data(dy, W, K, 8).shape == (K, 2, g.M, 8)
    assert torch.ops.dsb200.cheb_contract_bwd_weight(x, basis, dy, K).shape == (8 * K, 5)
