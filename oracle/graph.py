"""Oracle restatement of the reference's HEALPix graph / Laplacian construction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows deepsphere/utils.py of the
reference statement by statement (same numpy / scipy calls in the same order, so that the
float32 results are bit-identical on the same numpy); `hp` is oracle.healpix, the scalar
restatement of the two healpy calls.

Pinned by tests/golden/graph_*.npz, produced by running the reference's own utils.py in
the build container (tests/golden/make_golden.py).
"""
import numpy as np
from scipy import sparse
import scipy.sparse.linalg  # noqa: F401  (sparse.linalg.eigsh, utils.py:423)

from . import healpix as hp


def healpix_weightmatrix(nside=16, nest=True, indexes=None, dtype=np.float32, std=None, full=False):
    """deepsphere/utils.py:25-133 (8-neighbour Gaussian weight matrix, `new=False` graph)."""
    if not nest:
        raise NotImplementedError()                                     # utils.py:40-41
    if indexes is None:
        indexes = range(nside**2 * 12)                                  # utils.py:43-44
    npix = len(indexes)
    if npix >= (max(indexes) + 1):                                      # utils.py:46-53
        usefast = True
        indexes = range(npix)
    else:
        usefast = False
        indexes = list(indexes)

    x, y, z = hp.pix2vec(nside, indexes, nest=nest)                     # utils.py:56
    coords = np.vstack([x, y, z]).transpose()
    coords = np.asarray(coords, dtype=dtype)                            # utils.py:58
    if full:                                                            # utils.py:60-61
        from scipy import spatial
        distances = spatial.distance.cdist(coords, coords)**2
    else:
        neighbors = hp.get_all_neighbours(nside, indexes, nest=nest)    # utils.py:63
        col_index = neighbors.T.reshape((npix * 8))                     # utils.py:83
        row_index = np.repeat(indexes, 8)                               # utils.py:84
        if usefast:                                                     # utils.py:87-92
            keep = (col_index < npix)
            keep &= (col_index >= 0)
            col_index = col_index[keep]
            row_index = row_index[keep]
        else:                                                           # utils.py:93-100
            col_index_set = set(indexes)
            keep = [c in col_index_set for c in col_index]
            inverse_map = [np.nan] * (nside**2 * 12)
            for i, index in enumerate(indexes):
                inverse_map[index] = i
            col_index = [inverse_map[el] for el in col_index[keep]]
            row_index = [inverse_map[el] for el in row_index[keep]]
        distances = np.sum((coords[row_index] - coords[col_index])**2, axis=1)  # utils.py:103

    if std is None:                                                     # utils.py:107-110
        kernel_width = np.mean(distances)
    else:
        kernel_width = std
    weights = np.exp(-distances / (2 * kernel_width))                   # utils.py:111

    if full:                                                            # utils.py:119-125
        W = weights
        for i in range(len(W)):
            W[i, i] = 0.
        k = 0.01
        W[W < k] = 0
        W = sparse.csr_matrix(W, dtype=dtype)
    else:                                                               # utils.py:127-128
        W = sparse.csr_matrix(
            (weights, (row_index, col_index)), shape=(npix, npix), dtype=dtype)
    return W


def build_laplacian(W, lap_type='normalized', dtype=np.float32):
    """deepsphere/utils.py:231-242."""
    d = np.ravel(W.sum(1))
    if lap_type == 'combinatorial':
        D = sparse.diags(d, 0, dtype=dtype)
        return (D - W).tocsc()
    elif lap_type == 'normalized':
        d12 = np.power(d, -0.5)
        D12 = sparse.diags(np.ravel(d12), 0, dtype=dtype).tocsc()
        return sparse.identity(d.shape[0], dtype=dtype) - D12 * W * D12
    else:
        raise ValueError('Unknown Laplacian type {}'.format(lap_type))


def healpix_laplacian(nside=16, nest=True, lap_type='normalized', indexes=None,
                      dtype=np.float32, std=None, full=False):
    """deepsphere/utils.py:320-342, `new=False` branch (the in-repo graph)."""
    W = healpix_weightmatrix(nside=nside, nest=nest, indexes=indexes, dtype=dtype, std=std, full=full)
    return build_laplacian(W, lap_type=lap_type)


def rescale_L(L, lmax=2, scale=1):
    """deepsphere/utils.py:379-385."""
    M, M = L.shape
    I = sparse.identity(M, format='csr', dtype=L.dtype)
    L /= (lmax / 2)
    L -= I
    return L * scale


def estimate_lmax(laplacian, v0=None):
    """`1.02 * eigsh(L, k=1, which='LM')` of deepsphere/utils.py:423.  ARPACK's start vector
    is random in the reference; v0 may be fixed here for reproducible tests."""
    return 1.02 * sparse.linalg.eigsh(laplacian, k=1, which='LM', return_eigenvectors=False, v0=v0)[0]


def build_laplacians(nsides, indexes=None, std=None, full=False, lmax=None, new=False):
    """deepsphere/utils.py:388-431 for sampling='healpix', new=False.

    `lmax` (optional list, one per built level) replaces the ARPACK call so that a test can
    hand the same value to the oracle and to the product."""
    if new:
        raise NotImplementedError('the oracle restates the in-repo graph (new=False); the kNN graph lives in the PyGSP fork')
    L = []
    p = []
    if indexes is None:
        indexes = [None] * len(nsides)
    if not isinstance(std, list):
        std = [std] * len(nsides)
    if not isinstance(full, list):
        full = [full] * len(nsides)
    nside_last = -1
    built = 0
    for i, (nside, index, sigma, mat) in enumerate(zip(nsides, indexes, std, full)):
        if i > 0:
            p.append((nside_last // nside)**2)                          # utils.py:405-406
        if nside == nside_last and i < len(nsides) - 1:                 # utils.py:407-409
            L.append(L[-1].copy().tocoo())
            continue
        nside_last = nside
        if i < len(nsides) - 1:                                         # utils.py:411
            laplacian = healpix_laplacian(nside=nside, indexes=index, std=sigma, full=mat)
            lm = estimate_lmax(laplacian) if lmax is None else lmax[built]
            built += 1
            laplacian = rescale_L(laplacian, lmax=lm)                   # utils.py:424
            laplacian = laplacian.tocoo()                               # utils.py:425
            L.append(laplacian)
    return L, p


def nside2indexes(nsides, order):
    """deepsphere/utils.py:434-449."""
    nsample = 12 * order**2
    if order == 0:
        nsample = 1
    return [np.arange(hp.nside2npix(nside) // nsample) for nside in nsides]


def ds_index(index, nsides, nest=True):
    """deepsphere/utils.py:452-473."""
    assert isinstance(nsides, list)
    assert len(nsides) > 1
    assert nest
    indexes = [index]
    for nside in nsides[1:]:
        p = (nsides[0] / nside)**2
        if p < 1:
            raise NotImplementedError("upsampling not implemented yet")
        temp_index = index // p
        indexes.append(np.unique(temp_index).astype(int))
    return indexes
