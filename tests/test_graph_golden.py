"""Graph construction: oracle and product vs fixtures produced by the reference's own utils.py
(tests/golden/make_golden.py), and the two HEALPix geometry implementations vs each other."""
import os

import numpy as np
import pytest
from scipy import sparse

from conftest import GOLDEN
from oracle import graph as ograph
from oracle import healpix as ohp
from deepsphere_tf1_b200 import healpix as php
from deepsphere_tf1_b200 import utils as putils

G = np.load(os.path.join(GOLDEN, 'graph_golden.npz'))


def parts(A):
    A = sparse.csr_matrix(A).copy()
    A.sum_duplicates()
    A.sort_indices()
    return A.data, A.indices.astype(np.int64), A.indptr.astype(np.int64)


def assert_same(A, prefix, exact=True, rtol=0.0):
    data, ind, ptr = parts(A)
    assert tuple(G[prefix + '_shape']) == A.shape
    np.testing.assert_array_equal(ptr, G[prefix + '_indptr'])
    np.testing.assert_array_equal(ind, G[prefix + '_indices'])
    assert data.dtype == G[prefix + '_data'].dtype == np.float32
    if exact:
        np.testing.assert_array_equal(data, G[prefix + '_data'])      # bit-exact
    else:
        np.testing.assert_allclose(data, G[prefix + '_data'], rtol=rtol, atol=rtol)


IMPLS = [pytest.param(ograph, id='oracle'), pytest.param(putils, id='product')]


@pytest.mark.parametrize('mod', IMPLS)
@pytest.mark.parametrize('nside', [1, 2, 4, 8])
def test_full_sphere_weights_and_laplacians_bit_exact(mod, nside):
    W = mod.healpix_weightmatrix(nside=nside, nest=True, dtype=np.float32)
    assert_same(W, 'full%d_W' % nside)
    assert_same(mod.build_laplacian(W, 'normalized'), 'full%d_Lnorm' % nside)
    assert_same(mod.build_laplacian(W, 'combinatorial'), 'full%d_Lcomb' % nside)


@pytest.mark.parametrize('mod', IMPLS)
def test_partial_sphere_fast_path(mod):
    idx = mod.nside2indexes([32, 16, 8, 4, 4], 2)
    for i in range(5):
        np.testing.assert_array_equal(idx[i], G['n2i_order2_%d' % i])
    np.testing.assert_array_equal(mod.nside2indexes([32], 1)[0], G['n2i_order1_0'])
    np.testing.assert_array_equal(mod.nside2indexes([32, 16, 8, 4, 4], 0)[4], G['n2i_order0_4'])
    W = mod.healpix_weightmatrix(nside=32, nest=True, indexes=idx[0], dtype=np.float32)
    assert_same(W, 'part32o2_W')
    assert_same(mod.build_laplacian(W, 'normalized'), 'part32o2_Lnorm')


@pytest.mark.parametrize('mod', IMPLS)
def test_partial_sphere_arbitrary_indexes(mod):
    W = mod.healpix_weightmatrix(nside=8, nest=True, indexes=G['sel8_indexes'], dtype=np.float32)
    assert_same(W, 'sel8_W')


@pytest.mark.parametrize('mod', IMPLS)
def test_explicit_kernel_width(mod):
    assert_same(mod.healpix_weightmatrix(nside=4, nest=True, dtype=np.float32, std=0.05), 'full4std_W')


@pytest.mark.parametrize('mod', IMPLS)
def test_rescale_bit_exact(mod):
    W = mod.healpix_weightmatrix(nside=4, nest=True, dtype=np.float32)
    L = mod.build_laplacian(W, 'normalized')
    assert_same(mod.rescale_L(sparse.csr_matrix(L).copy(), lmax=np.float32(1.93)), 'rescale4')


@pytest.mark.parametrize('mod', IMPLS)
def test_ds_index(mod):
    ds = mod.ds_index(np.arange(512, 1024), [16, 8, 4])
    for i, a in enumerate(ds):
        np.testing.assert_array_equal(np.asarray(a), G['ds_index_%d' % i])


@pytest.mark.parametrize('mod', IMPLS)
def test_build_laplacians_pipeline(mod):
    """Whole pipeline incl. the ARPACK lmax (start-vector dependent -> 1e-5 tolerance)."""
    nsides = [32, 16, 8, 4, 4]
    Ls, p = mod.build_laplacians(nsides, indexes=mod.nside2indexes(nsides, 2), new=False)
    np.testing.assert_array_equal(np.array(p), G['bl_p'])
    assert len(Ls) == int(G['bl_n'])
    for i, L in enumerate(Ls):
        assert L.format == 'coo' and L.dtype == np.float32
        assert_same(L, 'bl_L%d' % i, exact=False, rtol=2e-5)
    Ls, p = mod.build_laplacians([4, 2, 1], new=False)
    np.testing.assert_array_equal(np.array(p), G['blfull_p'])
    for i, L in enumerate(Ls):
        assert_same(L, 'blfull_L%d' % i, exact=False, rtol=2e-5)


def test_build_laplacians_product_equals_oracle_with_same_lmax():
    nsides = [16, 8, 4, 4]
    idx = putils.nside2indexes(nsides, 1)
    lmax = [np.float32(1.9), np.float32(1.85), np.float32(1.8)]
    La, pa = ograph.build_laplacians(nsides, indexes=idx, lmax=lmax)
    Lb, pb = putils.build_laplacians(nsides, indexes=idx, lmax=lmax, new=False)
    assert pa == pb == [4, 4, 1]
    for A, B in zip(La, Lb):
        for u, v in zip(parts(A), parts(B)):
            np.testing.assert_array_equal(u, v)


# ---- HEALPix geometry: scalar oracle vs vectorised product, plus SURVEY Appendix C invariants

def test_healpy_docstring_example():
    want = [11, 7, 3, -1, 0, 5, 8, -1]
    assert list(ohp.get_all_neighbours(1, 4)) == want
    assert list(php.get_all_neighbours(1, 4)) == want
    # second example of the same docstring: hp.get_all_neighbours(1, np.pi / 2, np.pi / 2) -- the pixel at (theta, phi) =
    # (pi / 2, pi / 2) is base pixel 5
    want = [8, 4, 0, -1, 1, 6, 9, -1]
    assert list(ohp.get_all_neighbours(1, 5)) == want
    assert list(php.get_all_neighbours(1, 5)) == want


def test_published_pixel_centres():
    """Known answers that do not come from this repository (healpy itself is absent):
    (1) the 12 base pixels (Gorski et al. 2005, fig. 4: z = 2/3, 0, -2/3; phi = pi/4 + k pi/2 in the caps, k pi/2 on the equator;
        at Nside = 1 the NESTED and RING numbers coincide);
    (2) Nside = 2, children of base pixel 0 in NESTED order (south, east, west, north) from the ring formulas of the same
        paper: z = 1 - i^2 / (3 Nside^2) in the cap, 4/3 - 2 i / (3 Nside) in the belt;
    (3) the (theta, phi) pairs of healpy's own `pix2ang` docstring (`hp.pix2ang(16, [1440, 427, 1520, 0, 3068])`, RING
        numbers): the pixel SET is the same in both schemes, so each pair must be the centre of exactly one nested pixel."""
    for mod in (ohp, php):
        v = np.array(mod.pix2vec(1, np.arange(12) if mod is php else range(12)), dtype=np.float64)
        z = np.repeat([2 / 3, 0.0, -2 / 3], 4)
        phi = np.concatenate([np.pi / 4 + np.arange(4) * np.pi / 2, np.arange(4) * np.pi / 2, np.pi / 4 + np.arange(4) * np.pi / 2])
        np.testing.assert_allclose(v[2], z, atol=1e-15)
        np.testing.assert_allclose(np.arctan2(v[1], v[0]) % (2 * np.pi), phi, atol=1e-15)
        v = np.array(mod.pix2vec(2, np.arange(4) if mod is php else range(4)), dtype=np.float64)
        np.testing.assert_allclose(v[2], [1 / 3, 2 / 3, 2 / 3, 11 / 12], atol=1e-15)
        np.testing.assert_allclose(np.arctan2(v[1], v[0]), [np.pi / 4, 3 * np.pi / 8, np.pi / 8, np.pi / 4], atol=1e-15)
    v = np.array(php.pix2vec(16, np.arange(3072)))
    theta, phi = np.arccos(v[2]), np.arctan2(v[1], v[0]) % (2 * np.pi)
    for t, f in zip([1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608], [0., 0.78539816, 1.61988371, 0.78539816, 5.49778714]):
        hit = (np.abs(theta - t) < 2e-8) & (np.abs((phi - f + np.pi) % (2 * np.pi) - np.pi) < 2e-8)
        assert hit.sum() == 1, (t, f)


@pytest.mark.parametrize('nside', [1, 2, 4, 8, 16])
def test_geometry_oracle_equals_product(nside):
    npix = 12 * nside * nside
    a = np.array(ohp.pix2vec(nside, range(npix)))
    b = np.array(php.pix2vec(nside, np.arange(npix)))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ohp.get_all_neighbours(nside, range(npix)),
                                  php.get_all_neighbours(nside, np.arange(npix)))


@pytest.mark.parametrize('nside', [2, 8, 32, 64])
def test_geometry_invariants(nside):
    npix = 12 * nside * nside
    pix = np.arange(npix)
    v = np.array(php.pix2vec(nside, pix))
    assert np.abs(np.linalg.norm(v, axis=0) - 1).max() < 1e-15
    assert len(np.unique(np.round(v[2], 12))) == 4 * nside - 1
    assert np.abs(v.sum(axis=1)).max() < 1e-9
    nb = php.get_all_neighbours(nside, pix)
    missing = (nb < 0).sum(axis=0)
    assert (missing == 1).sum() == 24 and (missing > 1).sum() == 0
    rows = np.repeat(pix, 8)
    cols = nb.T.reshape(-1)
    keep = cols >= 0
    A = sparse.csr_matrix((np.ones(keep.sum()), (rows[keep], cols[keep])), shape=(npix, npix))
    assert A.max() == 1                               # no duplicate neighbours
    assert (A != A.T).nnz == 0                        # symmetric relation
    assert A.nnz == 8 * npix - 24
    d2 = ((v[:, rows[keep]] - v[:, cols[keep]])**2).sum(axis=0) / (4 * np.pi / npix)
    assert 0.6 < d2.min() and d2.max() < 4.5
    if nside >= 2:                                    # nested parent = mean of the 4 children
        parent = np.array(php.pix2vec(nside // 2, np.arange(npix // 4)))
        child = v.reshape(3, npix // 4, 4).mean(axis=2)
        child /= np.linalg.norm(child, axis=0)
        ang = np.arccos(np.clip((parent * child).sum(axis=0), -1, 1))
        assert ang.max() < 0.2 * np.sqrt(4 * np.pi / (npix // 4))


def test_cosmo_patch_is_a_square_grid():
    """order-2 patches are open-boundary squares; degree histogram."""
    nside = 64
    idx = putils.nside2indexes([nside], 2)[0]
    W = putils.healpix_weightmatrix(nside=nside, indexes=idx)
    n = nside // 2
    deg = np.diff(W.indptr)
    assert (deg == 3).sum() == 4 and (deg == 5).sum() == 4 * (n - 2) and (deg == 8).sum() == (n - 2)**2
    assert abs(W - W.T).max() < 1e-7


@pytest.mark.parametrize('bw', [2, 4, 8])
def test_equiangular_graph_bit_exact(bw):
    """utils.equiangular_weightmatrix (vectorised) against the reference's own loop (upstream utils.py:136-229)."""
    W = putils.equiangular_weightmatrix(bw=bw, dtype=np.float32)
    assert_same(W, 'equi%d_W' % bw)
    assert_same(putils.build_laplacian(W, 'normalized'), 'equi%d_Lnorm' % bw)
    n = 2 * bw
    A = sparse.csr_matrix(W)
    assert A.shape == (n * n, n * n) and (abs(A - A.T) > 1e-6 * abs(A).max()).nnz == 0      # the neighbour relation is symmetric
    if bw >= 4:                                                     # (bw = 1 is degenerate: coincident neighbours, infinite weights)
        L, p = putils.build_laplacians([bw, bw // 2, bw // 2], sampling='equiangular')
        assert p == [4, 1] and [l.shape[0] for l in L] == [n * n, n * n // 4]


def test_equiangular_z_order_makes_2d_windows_consecutive():
    """models.cgcnn._equi_perm: in Z order every sqrt(p) x sqrt(p) window of the grid is p consecutive vertices, at every level."""
    from deepsphere_tf1_b200 import models
    m = models.cgcnn.__new__(models.cgcnn)
    for n in (2, 4, 16):
        perm, inv, _ = m._equi_perm(n * n)
        assert sorted(perm) == list(range(n * n)) and (inv[perm] == np.arange(n * n)).all()
        i, j = perm // n, perm % n
        for s in (2, 4):
            if s > n:
                continue
            blk = (i // s) * (n // s) + (j // s)
            g = blk.reshape(-1, s * s)
            assert (g == g[:, :1]).all()
            # ... and the windows themselves are in the Z order of the coarser grid
            pc, _, _ = m._equi_perm((n // s) ** 2) if n // s > 1 else (np.zeros(1, np.int64), None, None)
            if n // s > 1:
                assert (g[:, 0] == pc).all()


@pytest.mark.parametrize('nside,order', [(8, 2), (16, 1), (32, 2)])
def test_four_neighbour_patch_graph_bit_exact(nside, order):
    """utils.build_matrix_4_neighboors (use_4=True; array operations) against the reference's digit-walking loop
    (upstream utils.py:484-591): the 4-connected grid of a nested-order patch, Gaussian weights with width 3 x mean d^2."""
    idx = putils.nside2indexes([nside], order)[0]
    W = putils.build_matrix_4_neighboors(nside, idx)
    assert_same(W, 'four%d_%d_W' % (nside, order))
    A = sparse.csr_matrix(W)
    n = int(round(len(idx) ** 0.5))
    deg = np.diff(A.indptr)
    assert deg.min() == 2 and deg.max() == 4 and (deg == 2).sum() == 4 and A.nnz == 4 * n * (n - 1)     # an n x n grid
    L, p = putils.build_laplacians([nside, nside // 2, nside // 2], indexes=putils.nside2indexes([nside, nside // 2, nside // 2], order),
                                   use_4=True, new=False)
    assert p == [4, 1] and L[0].shape[0] == len(idx)
