"""Device-side HEALPix graph / Laplacian builder (graph_build.py, csrc/healpix_build.cu) against the numpy path of utils.py
(itself pinned bit for bit to the reference's own code, tests/test_graph_golden.py): same vertices, same sparsity pattern, the same
float32 bits for the same lmax; lmax from the device Lanczos within 1e-5 of ARPACK's."""
import time

import numpy as np
import pytest
import torch
from scipy import sparse

from deepsphere_tf1_b200 import utils

pytestmark = pytest.mark.gpu


def _host(nside, indexes, lap_type, lmax, std=None):
    W = utils.healpix_weightmatrix(nside=nside, indexes=indexes, dtype=np.float32, std=std)
    L = utils.build_laplacian(W, lap_type)
    return utils.rescale_L(sparse.csr_matrix(L), lmax=lmax).tocoo()


@pytest.mark.parametrize('case', [(2, None), (4, None), (8, None), (16, None), (32, None), (64, None), (64, 'order2'),
                                  (32, 'order1'), (8, 'scattered')])
@pytest.mark.parametrize('lap_type', ['normalized', 'combinatorial'])
def test_device_builder_equals_the_numpy_path_bit_for_bit(dev, case, lap_type):
    from deepsphere_tf1_b200 import graph_build
    from deepsphere_tf1_b200.graph import DeviceGraph
    nside, kind = case
    indexes = None
    if kind == 'order2':
        indexes = utils.nside2indexes([nside], 2)[0]
    elif kind == 'order1':
        indexes = utils.nside2indexes([nside], 1)[0]
    elif kind == 'scattered':
        indexes = np.array([3, 4, 5, 6, 7, 40, 41, 42, 43, 200, 201, 203, 500, 501, 502, 503, 767])
    lmax = np.float32(1.93)
    g, L, _ = graph_build.healpix_level(nside, indexes, dev, lap_type=lap_type, lmax=lmax)
    Lh = _host(nside, indexes, lap_type, lmax)
    ref = DeviceGraph(Lh, dev)
    assert g.M == ref.M and g.nnz == ref.nnz and g.fwd[3] == ref.fwd[3]
    assert torch.equal(g.fwd[1], ref.fwd[1]), 'columns'
    assert torch.equal(g.fwd[2], ref.fwd[2]), 'values'
    assert torch.equal(g.bwd[2], ref.bwd[2]), 'transposed values'
    assert torch.equal(g.fwd_rec, ref.fwd_rec)
    A, B = sparse.csr_matrix(L), sparse.csr_matrix(Lh)
    A.sort_indices(), B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)


def test_explicit_kernel_width(dev):
    from deepsphere_tf1_b200 import graph_build
    from deepsphere_tf1_b200.graph import DeviceGraph
    g, L, _ = graph_build.healpix_level(4, None, dev, std=0.05, lmax=np.float32(1.5))
    ref = DeviceGraph(_host(4, None, 'normalized', np.float32(1.5), std=0.05), dev)
    assert torch.equal(g.fwd[2], ref.fwd[2])


@pytest.mark.parametrize('nside,order', [(16, 0), (64, 2), (256, 2)])
def test_lanczos_lmax_agrees_with_arpack(dev, nside, order):
    from deepsphere_tf1_b200 import graph_build
    indexes = utils.nside2indexes([nside], order)[0] if order else None
    W = utils.healpix_weightmatrix(nside=nside, indexes=indexes, dtype=np.float32)
    Lu = utils.build_laplacian(W, 'normalized')
    expect = utils.estimate_lmax(Lu)
    g, _, lmax = graph_build.healpix_level(nside, indexes, dev, with_scipy=False)
    # the Ritz value approaches the largest eigenvalue from below
    assert float(expect) * (1 - 2e-5) <= float(lmax) <= float(expect) * (1 + 2e-6)


def test_build_laplacians_through_the_model_front_end(dev):
    """deepsphere(new=False, device_build=True): the same model as with the numpy builder (same lmax), and the device graphs are
    the ones the layers use."""
    from deepsphere_tf1_b200 import models, optimizers
    nsides = [32, 16, 8, 8]
    idx = utils.nside2indexes(nsides, 1)
    Lh, ph = utils.build_laplacians(nsides, indexes=idx, new=False)
    lmax = [float(utils.estimate_lmax(utils.healpix_laplacian(nside=n, indexes=i, new=False))) for n, i in zip(nsides[:3], idx[:3])]
    kw = dict(F=[8, 16, 2], K=[4, 4, 4], batch_norm=[True] * 3, M=[], statistics='mean', num_epochs=1, batch_size=2, verbose=False,
              scheduler=lambda s: 1e-3, optimizer=lambda lr: optimizers.AdamOptimizer(lr), seed=0)
    a = models.deepsphere(nsides=nsides, indexes=idx, new=False, lmax=lmax, device_build=True, dir_name='_gb_a', **kw)
    b = models.deepsphere(nsides=nsides, indexes=idx, new=False, lmax=lmax, device_build=False, dir_name='_gb_b', **kw)
    assert a.p == b.p == ph
    for La, Lb in zip(a.L, b.L):
        assert (sparse.csr_matrix(La) != sparse.csr_matrix(Lb)).nnz == 0
    assert a._graphs[0] is a._prebuilt_graphs[0]
    x = np.random.RandomState(0).randn(2, 1024, 1).astype(np.float32)
    lab = np.array([0, 1], np.int32)
    la = float(a.train_step(x, lab)[1])
    lb = float(b.train_step(x, lab)[1])
    assert la == pytest.approx(lb, rel=1e-6)


def test_full_sphere_nside_512_builds_in_seconds(dev):
    from deepsphere_tf1_b200 import graph_build
    torch.cuda.synchronize()
    t = time.time()
    g, L, lmax = graph_build.healpix_level(512, None, dev, with_scipy=False)
    torch.cuda.synchronize()
    dt = time.time() - t
    print('nside 512 full sphere ({} vertices): {:.2f} s, lmax {:.6f}'.format(g.M, dt, float(lmax)))
    assert g.M == 12 * 512 * 512 and g.nnz == 9 * g.M - 24 and dt < 20
    assert 1.5 < float(lmax) < 1.6
