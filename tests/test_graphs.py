"""Icosahedral / kNN graph builders (deepsphere_tf1_b200/graphs.py, the graphs the reference takes from a PyGSP fork).
The icosahedral MESH is pinned bit for bit to the reference's own class source (tests/golden/make_ico_golden.py executes
notebooks/sphere_to_graph.ipynb's `SphereIcosahedron`); the kNN weights (PyGSP's NNGraph, absent) through invariants."""
import os

import numpy as np
import pytest
from scipy import sparse

from deepsphere_tf1_b200 import graphs, utils

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ico_golden.npz'))


@pytest.mark.parametrize('level', [0, 1, 2, 3, 4])
def test_icosahedron_mesh_equals_the_reference_class(level):
    coords, faces = graphs.icosahedron_mesh(level)
    assert np.array_equal(coords, GOLD['coords%d' % level])
    assert np.array_equal(faces, GOLD['faces%d' % level])
    assert int(GOLD['k%d' % level]) == (5 if level == 0 else 6)             # the k the class hands to NNGraph
    assert coords.shape == (10 * 4 ** level + 2, 3) and faces.shape == (20 * 4 ** level, 3)
    np.testing.assert_allclose(np.linalg.norm(coords, axis=1), 1.0, atol=1e-14)


def test_coarser_level_is_a_prefix():
    """`divide()` appends the edge mid-points: what the icosahedral pooling x[:, :p, :] relies on (models.py:1137-1138)."""
    fine, _ = graphs.icosahedron_mesh(3)
    for level in (0, 1, 2):
        coarse, _ = graphs.icosahedron_mesh(level)
        np.testing.assert_allclose(fine[:len(coarse)], coarse, atol=1e-12)


def test_nn_graph_invariants():
    rs = np.random.RandomState(0)
    pts = rs.randn(400, 3)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    k = 7
    W = graphs.nn_graph(pts, k=k, center=False, rescale=False, dtype=np.float64)
    assert abs(W - W.T).max() == 0 and W.diagonal().max() == 0
    deg = np.diff(W.indptr)
    assert deg.min() >= k                                                    # symmetrisation only adds neighbours
    # every weight is exp(-d^2 / sigma) or half of it (edge found from one side only), sigma = mean neighbour distance
    d = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
    sigma = np.sort(d, axis=1)[:, 1:k + 1].mean()
    Wc = W.tocoo()
    full = np.exp(-d[Wc.row, Wc.col] ** 2 / sigma)
    assert np.all(np.isclose(Wc.data, full) | np.isclose(Wc.data, full / 2))
    # PyGSP's centring / rescaling changes the kernel width only through the distances
    W2 = graphs.nn_graph(pts, k=k, dtype=np.float64)
    assert (W2 != 0).nnz and np.array_equal((W2 != 0).toarray(), (W != 0).toarray())


def test_icosahedral_laplacians_through_build_laplacians():
    """utils.build_laplacians(sampling='icosahedron') (utils.py:388-431): one Laplacian per level but the last, p = vertex
    counts of the following levels, spectrum in [-1, 1] after the rescale."""
    L, p = utils.build_laplacians([3, 3, 2, 1, 0, 0], sampling='icosahedron')
    assert [l.shape[0] for l in L] == [642, 642, 162, 42, 12] and p == [642, 162, 42, 12, 12]
    for l in L:
        assert l.dtype == np.float32 and sparse.isspmatrix_coo(l)
        ev = np.linalg.eigvalsh(l.toarray().astype(np.float64))
        assert ev.min() >= -1 - 1e-5 and ev.max() <= 1.0
        assert abs(l - l.T).max() < 1e-6
    assert np.diff(L[2].tocsr().indptr).max() <= 8                           # 6 or 7 neighbours + the diagonal


def test_healpix_knn_graph_is_the_upstream_default():
    """new=True (reference default, utils.py:329-334): kNN graph on the pixel centres, kernel width from the notebook's table."""
    L, p = utils.build_laplacians([8, 4, 2])
    assert p == [4, 4] and [l.shape[0] for l in L] == [768, 192]
    W = graphs.healpix_nn_weights(8, n_neighbors=8, dtype=np.float64)
    assert abs(W - W.T).max() == 0 and np.diff(W.indptr).min() >= 8
    # partial sphere and the wide neighbourhoods of the fork (20 / 40 neighbours)
    W20 = graphs.healpix_nn_weights(8, indexes=np.arange(192), n_neighbors=20)
    assert W20.shape == (192, 192) and np.diff(W20.indptr).min() >= 20
    Lold, _ = utils.build_laplacians([8, 4, 2], new=False)
    assert (Lold[0] != L[0]).nnz > 0


def test_station_graph_of_the_ghcn_experiment():
    rs = np.random.RandomState(1)
    lon, lat = rs.uniform(-np.pi, np.pi, 300), np.arcsin(rs.uniform(-1, 1, 300))
    W, coords = graphs.sphere_knn_weights(lon, lat, 10)
    np.testing.assert_allclose(np.linalg.norm(coords, axis=1), 1.0, atol=1e-12)
    assert W.shape == (300, 300) and abs(W - W.T).max() == 0
    L = graphs.sphere_knn_laplacian(lon, lat, 10)
    np.testing.assert_allclose(np.asarray(L.sum(axis=1)).ravel(), 0, atol=1e-4)      # combinatorial: rows sum to zero


@pytest.mark.parametrize('rows,cols', [(8, 8), (8, 12), (6, 9), (24, 36), (4, 32)])
def test_equiangular_block_order_makes_2d_windows_consecutive(rows, cols):
    """The product pools p CONSECUTIVE vertices; on an equiangular grid the reference pools sqrt(p) x sqrt(p) windows of the
    [rows, cols] reshape (models.py:1131-1136).  utils.equiangular_block_order is the vertex order that makes the two the same,
    level after level: checked here against the oracle's 2-D pooling / unpooling for every depth the grid allows."""
    import torch
    from deepsphere_tf1_b200 import utils
    from oracle import ops as oops
    perm = utils.equiangular_block_order(rows, cols)
    assert sorted(perm.tolist()) == list(range(rows * cols))
    ratio = cols / rows
    rs = np.random.RandomState(rows * 100 + cols)
    x = torch.from_numpy(rs.randn(2, rows * cols, 3).astype(np.float32))
    r, c, xz = rows, cols, x[:, perm]
    while r % 2 == 0 and c % 2 == 0:
        for kind in ('max', 'average'):
            want = oops.pool(x, 4, kind, 'equiangular', ratio)                      # row-major in, row-major out
            got = xz.reshape(2, -1, 4, 3)
            got = got.max(dim=2).values if kind == 'max' else got.mean(dim=2)        # what the HEALPix kernels do
            coarse = utils.equiangular_block_order(r // 2, c // 2)
            assert torch.allclose(got, want[:, coarse], atol=1e-6), (r, c, kind)
        # unpooling: repeat each vertex 4 times in the block order == repeat_elements on both axes (models.py:1348-1355)
        up = oops.unpool(want, 4, 'average', 'equiangular', ratio)
        assert torch.equal(want[:, coarse].repeat_interleave(4, dim=1), up[:, utils.equiangular_block_order(r, c)])
        x, xz, r, c = want, want[:, coarse], r // 2, c // 2
    with pytest.raises(ValueError):
        utils.equiangular_block_order(rows, cols, depth=10)


def test_rectangular_equiangular_graph():
    """Pair of bandwidths (utils.py:401-403): the in-repo 8-neighbour construction on the 2 bw[0] x 2 bw[1] grid; equal bandwidths
    give the square graph bit for bit (which the golden vectors pin to the reference's own function)."""
    from deepsphere_tf1_b200 import utils
    W = utils.equiangular_weightmatrix((4, 6))
    assert W.shape == (96, 96) and (np.diff(W.indptr) == 8).all() and np.isfinite(W.data).all()
    assert abs(W - W.T).max() < 1e-5 * W.max()
    assert (utils.equiangular_weightmatrix(4) != utils.equiangular_weightmatrix((4, 4))).nnz == 0
    L, p = utils.build_laplacians([(8, 12), (4, 6), (2, 3), (2, 3)], sampling='equiangular')
    assert p == [4, 4, 1] and [l.shape[0] for l in L] == [384, 96, 24]
