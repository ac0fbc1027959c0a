"""Graph / Laplacian construction for the ChebConv hot path (host side, numpy + scipy).

Same public functions, argument meaning and error behaviour as the reference's
`deepsphere/utils.py` for the HEALPix path of `models.deepsphere.__init__`
(reference deepsphere/models.py:1827-1839 -> utils.py:388-431), but built from whole-array
operations on the COO triplets instead of per-pixel healpy calls and sparse-sparse products:

  healpix_weightmatrix  <- utils.py:25-133      build_laplacian <- utils.py:231-242
  healpix_laplacian     <- utils.py:320-342     rescale_L       <- utils.py:379-385
  build_laplacians      <- utils.py:388-431     nside2indexes   <- utils.py:434-449
  ds_index              <- utils.py:452-473

Float32 results are bit-identical to the reference code run on the same numpy (checked in
tests/test_graph_golden.py against fixtures produced by the reference's own utils.py).
The icosahedral, kNN-HEALPix (`new=True`) and irregular station graphs live in `graphs.py`.
"""
import numpy as np
from scipy import sparse
import scipy.sparse.linalg  # noqa: F401

from . import healpix as hp


def healpix_weightmatrix(nside=16, nest=True, indexes=None, dtype=np.float32, std=None, full=False):
    """Unnormalised Gaussian weight matrix of the HEALPix 8-neighbour graph (CSR)."""
    if not nest:
        raise NotImplementedError()
    if indexes is None:
        indexes = np.arange(nside**2 * 12)
    indexes = np.asarray(indexes)
    npix = len(indexes)
    usefast = npix >= (int(indexes.max()) + 1)
    if usefast:
        indexes = np.arange(npix)

    x, y, z = hp.pix2vec(nside, indexes, nest=True)
    coords = np.asarray(np.vstack([x, y, z]).transpose(), dtype=dtype)
    if full:
        from scipy import spatial
        distances = spatial.distance.cdist(coords, coords)**2
        kernel_width = np.mean(distances) if std is None else std
        W = np.exp(-distances / (2 * kernel_width))
        np.fill_diagonal(W, 0.)
        W[W < 0.01] = 0
        return sparse.csr_matrix(W, dtype=dtype)

    neighbors = hp.get_all_neighbours(nside, indexes, nest=True)
    col_index = neighbors.T.reshape(npix * 8)
    if usefast:
        row_index = np.repeat(indexes, 8)
        keep = (col_index < npix) & (col_index >= 0)
        col_index = col_index[keep]
        row_index = row_index[keep]
    else:
        # arbitrary pixel subset: translate sphere pixel numbers to positions in `indexes`
        inverse_map = np.full(nside**2 * 12, -1, dtype=np.int64)
        inverse_map[indexes] = np.arange(npix)
        row_index = np.repeat(np.arange(npix), 8)
        keep = col_index >= 0
        keep[keep] = inverse_map[col_index[keep]] >= 0
        col_index = inverse_map[col_index[keep]]
        row_index = row_index[keep]

    distances = np.sum((coords[row_index] - coords[col_index])**2, axis=1)
    kernel_width = np.mean(distances) if std is None else std
    weights = np.exp(-distances / (2 * kernel_width))
    return sparse.csr_matrix((weights, (row_index, col_index)), shape=(npix, npix), dtype=dtype)


def build_matrix_4_neighboors(nside, indexes, nest=True, dtype=np.float32):
    """4-neighbour grid graph of a HEALPix patch (upstream utils.py:484-591, `use_4=True`).  The patch is one base pixel of the
    coarser resolution `order`, i.e. the nested pixels 0 .. (nside / order)^2 - 1: a square in Z (Morton) order.  Upstream walks
    the digits of every index; here the whole edge list is built with array operations, in upstream's ORDER of entries (per pixel:
    the two neighbours inside its 2 x 2 block, then per level the -x / +x and the -y / +y crossings it tests) because the kernel
    width is the mean over that list.  Gaussian weights exp(-d^2 / (3 mean d^2)) on float64 coordinates."""
    assert nest
    idx = np.asarray(indexes, dtype=np.int64)
    order = nside // int(round(((int(idx.max()) + 1)) ** 0.5))         # hp.npix2nside(12 * (max + 1))
    npix = hp.nside2npix(nside) // hp.nside2npix(order)
    assert set(idx.tolist()) == set(range(npix))
    x, y, z = hp.pix2vec(nside, idx, nest=True)
    coords = np.vstack([x, y, z]).transpose()
    d = idx % 4
    base = idx - d
    cols = [base + np.array([1, 3, 0, 2])[d], base + np.array([2, 0, 3, 1])[d]]
    valid = [np.ones(len(idx), bool), np.ones(len(idx), bool)]
    levels = int(np.log2(nside) - np.log2(order) - 1)
    digits = []
    for it in range(levels):
        digits.append((idx // 4 ** it) % 4)                         # digits of the levels 0 .. it (d3 upstream)
        d2 = (idx // 4 ** (it + 1)) % 4
        D = np.stack(digits)

        def all_in(a, b):
            return ((D == a) | (D == b)).all(axis=0)
        in13, in23, in02, in01 = all_in(1, 3), all_in(2, 3), all_in(0, 2), all_in(0, 1)
        shift = 4 ** (it + 1) - sum(4 ** j for j in range(it + 1))
        # first / second test of upstream's branch for d2 = 0, 1, 2, 3
        colA = idx + np.array([shift, -shift, -2 * shift, -2 * shift])[d2]
        okA = np.choose(d2, [in13, in02, in01, in01])
        colB = idx + np.array([2 * shift, 2 * shift, shift, -shift])[d2]
        okB = np.choose(d2, [in23, in23, in13, in02])
        cols += [colA, colB]
        valid += [okA, okB]
    C = np.stack(cols, axis=1)
    V = np.stack(valid, axis=1)
    row_index = np.repeat(idx, C.shape[1])[V.reshape(-1)]
    col_index = C.reshape(-1)[V.reshape(-1)]
    distances = np.sum((coords[row_index] - coords[col_index]) ** 2, axis=1)
    kernel_width = np.mean(distances)
    weights = np.exp(-distances / (3 * kernel_width))
    return sparse.csr_matrix((weights, (row_index, col_index)), shape=(npix, npix), dtype=dtype)


def build_laplacian(W, lap_type='normalized', dtype=np.float32):
    """Combinatorial `D - W` or normalised `I - D^-1/2 W D^-1/2` Laplacian of a weight matrix."""
    d = np.ravel(W.sum(1))
    if lap_type == 'combinatorial':
        D = sparse.diags(d, 0, dtype=dtype)
        return (D - W).tocsc()
    elif lap_type == 'normalized':
        # One product per stored entry, evaluated as (d_i^-1/2 * w_ij) * d_j^-1/2 like the
        # reference's `D12 * W * D12`; no sparse-sparse multiplication needed.
        Wc = W.tocoo()
        d12 = np.power(d, -0.5).astype(dtype)
        off = (d12[Wc.row] * Wc.data.astype(dtype)) * d12[Wc.col]
        n = d.shape[0]
        diag = np.arange(n)
        rows = np.concatenate([diag, Wc.row])
        cols = np.concatenate([diag, Wc.col])
        vals = np.concatenate([np.ones(n, dtype=dtype), -off])
        return sparse.csr_matrix((vals, (rows, cols)), shape=(n, n), dtype=dtype)
    else:
        raise ValueError('Unknown Laplacian type {}'.format(lap_type))


def healpix_laplacian(nside=16, nest=True, lap_type='normalized', indexes=None, dtype=np.float32,
                      use_4=False, std=None, full=False, new=True, n_neighbors=8):
    """HEALPix Laplacian (utils.py:320-342).  `new=True` (the upstream default): the kNN graph of the PyGSP fork's
    `SphereHealpix`, restated in graphs.py; `new=False`: the in-repo 8-neighbour graph."""
    if new:
        if not nest:
            raise NotImplementedError()
        from . import graphs
        return graphs.healpix_nn_laplacian(nside, indexes=indexes, n_neighbors=n_neighbors, lap_type=lap_type, dtype=dtype)
    if use_4:
        W = build_matrix_4_neighboors(nside, indexes, nest=nest, dtype=dtype)          # utils.py:336-337
        return build_laplacian(W, lap_type=lap_type)
    W = healpix_weightmatrix(nside=nside, nest=nest, indexes=indexes, dtype=dtype, std=std, full=full)
    return build_laplacian(W, lap_type=lap_type)


def equiangular_weightmatrix(bw=64, indexes=None, dtype=np.float32):
    """8-neighbour graph of the SOFT equiangular grid of bandwidth bw (2bw x 2bw vertices, row-major in (colatitude, longitude)),
    weights 1 / squared chord distance (upstream utils.py:136-229, the only equiangular graph that is fully in the reference's
    repository; its neighbour functions wrap the longitude and cross the poles by half a turn).  Vectorised; same float32
    arithmetic and the same entry order as upstream (neighbours SW, W, NW, N, NE, E, SE, S), so the CSR matrix is bit-identical.
    `bw` may be a pair (bw_colatitude, bw_longitude): the same construction on the 2 bw[0] x 2 bw[1] grid (upstream passes such
    pairs to the absent fork's SphereEquiangular, utils.py:401-416; this rectangular form is an extension with nothing upstream
    to be pinned to -- for equal bandwidths it is the square grid above, bit for bit)."""
    if indexes is not None:
        raise NotImplementedError('equiangular_weightmatrix: partial grids are not supported (upstream builds a full-size matrix)')
    bw_lat, bw_lon = (int(bw[0]), int(bw[1])) if isinstance(bw, (tuple, list)) else (int(bw), int(bw))
    nr, n = 2 * bw_lat, 2 * bw_lon                             # rows (colatitude), columns (longitude)
    npix = nr * n
    beta = np.pi * (2 * np.arange(nr) + 1) / (4. * bw_lat)     # SOFT
    alpha = np.arange(n) * np.pi / bw_lon
    theta, phi = np.meshgrid(beta, alpha, indexing='ij')
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    coords = np.vstack([(st * cp).flatten(), (st * sp).flatten(), ct.flatten()]).transpose()
    coords = np.asarray(coords, dtype=dtype)
    ind = np.arange(npix, dtype=np.int64)

    def south(x):
        return np.where(x >= npix - n, (x + bw_lon) % n + npix - n, x + n)

    def north(x):
        return np.where(x < n, (x + bw_lon) % n, x - n)

    def west(x):
        return np.where(x % n == 0, x + n, x) - 1

    def east(x):
        return np.where(x % n == n - 1, x - n, x) + 1

    w, e = west(ind), east(ind)
    nb = np.stack([south(w), w, north(w), north(ind), north(e), e, south(e), south(ind)], axis=1)
    col_index = nb.reshape(-1)
    row_index = np.repeat(ind, 8)
    distances = np.sum((coords[row_index] - coords[col_index])**2, axis=1)
    weights = 1 / distances
    return sparse.csr_matrix((weights, (row_index, col_index)), shape=(npix, npix), dtype=dtype)


def equiangular_block_order(rows, cols, depth=None):
    """Vertex order of a rows x cols equiangular grid in which every aligned 2^s x 2^s window (s <= depth) is 4^s CONSECUTIVE
    vertices, so that the 2-D pooling / unpooling of models.py:1131-1136, 1150-1155, 1348-1355 becomes the 1-D pooling over
    consecutive vertices of the HEALPix kernels: the grid is cut into 2^depth x 2^depth blocks, blocks in row-major order,
    Z (Morton) order inside a block.  Returns perm with perm[k] = row-major index of the k-th vertex.  depth defaults to the
    largest one both sides allow; the order induced on the grid pooled by 2^s is the same construction with depth - s."""
    rows, cols = int(rows), int(cols)
    dmax = 0
    while rows % (2 << dmax) == 0 and cols % (2 << dmax) == 0:
        dmax += 1
    depth = dmax if depth is None else int(depth)
    if depth > dmax:
        raise ValueError('a {} x {} grid cannot be pooled {} times by 2 x 2'.format(rows, cols, depth))
    k = np.arange(rows * cols, dtype=np.int64)
    blk, within = k >> (2 * depth), k & ((1 << (2 * depth)) - 1)
    bi, bj = np.divmod(blk, cols >> depth)
    wi = np.zeros_like(k)
    wj = np.zeros_like(k)
    for b in range(depth):
        wj |= ((within >> (2 * b)) & 1) << b
        wi |= ((within >> (2 * b + 1)) & 1) << b
    return ((bi << depth) + wi) * cols + (bj << depth) + wj


def equiangular_laplacian(bw=16, lap_type='normalized', indexes=None, dtype=np.float32, use_4=False):
    """Laplacian of the equiangular graph.  Upstream (utils.py:344-361) takes it from the absent PyGSP fork's
    SphereEquiangular(bandwidth, sampling='SOFT'), whose neighbourhood and weights are not pinned anywhere in the reference;
    this is the in-repository construction its commented-out lines name: equiangular_weightmatrix + build_laplacian."""
    if use_4:
        raise NotImplementedError()
    return build_laplacian(equiangular_weightmatrix(bw=bw, indexes=indexes, dtype=dtype), lap_type=lap_type)


def rescale_L(L, lmax=2, scale=1):
    """Rescale the Laplacian eigenvalues to [-scale, scale]:  L / (lmax/2) - I."""
    M, M = L.shape
    I = sparse.identity(M, format='csr', dtype=L.dtype)
    L /= (lmax / 2)
    L -= I
    return L * scale


def estimate_lmax(laplacian):
    """1.02 x largest eigenvalue (ARPACK, deterministic start vector instead of a random one)."""
    n = laplacian.shape[0]
    if n < 3:
        return 1.02 * np.linalg.eigvalsh(laplacian.toarray()).max().astype(laplacian.dtype)
    v0 = np.cos(np.arange(n, dtype=np.float64) * 0.7390851332151607).astype(laplacian.dtype)
    return 1.02 * sparse.linalg.eigsh(laplacian, k=1, which='LM', return_eigenvectors=False, v0=v0)[0]


def build_laplacians(nsides, indexes=None, use_4=False, sampling='healpix', std=None, full=False,
                     new=True, n_neighbors=8, lmax=None):
    """List of rescaled Laplacians (COO, float32) and pooling factors from a list of nsides.

    `lmax` (optional list, one value per level actually built) skips the eigenvalue solve.
    """
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
        bw = nside
        if isinstance(nside, tuple):                           # equiangular: (bandwidth in colatitude, in longitude), utils.py:401-403
            nside = nside[0]
        if i > 0 and sampling != 'icosahedron':
            p.append((nside_last // nside)**2)
        if nside == nside_last and i < len(nsides) - 1:
            L.append(L[-1].copy().tocoo())
            continue
        nside_last = nside
        if i < len(nsides) - 1:
            if sampling == 'healpix':
                laplacian = healpix_laplacian(nside=nside, indexes=index, use_4=use_4, std=sigma,
                                              full=mat, new=new, n_neighbors=n_neighbors)
            elif sampling == 'icosahedron':
                from . import graphs
                laplacian = graphs.icosahedron_laplacian(order=nside, indexes=index)    # utils.py:370-376, 412
            elif sampling == 'equiangular':
                laplacian = equiangular_laplacian(bw=bw, indexes=index, use_4=use_4)          # utils.py:415-416
            else:
                raise ValueError('Unknown sampling: ' + sampling)
            lm = estimate_lmax(laplacian) if lmax is None else lmax[built]
            built += 1
            laplacian = rescale_L(laplacian, lmax=lm)
            L.append(laplacian.tocoo())
    if sampling == 'icosahedron':
        for order in nsides[1:]:
            p.append(10 * 4 ** order + 2)
    return L, p


def nside2indexes(nsides, order):
    """Pixel index lists for a 1/(12 order^2) part of the sphere at every nside."""
    nsample = 12 * order**2
    if order == 0:
        nsample = 1
    return [np.arange(hp.nside2npix(nside) // nsample) for nside in nsides]


def ds_index(index, nsides, nest=True):
    """Down-sample a nested pixel list given at nsides[0] to every coarser nside."""
    assert isinstance(nsides, list)
    assert len(nsides) > 1
    assert nest
    indexes = [index]
    for nside in nsides[1:]:
        p = (nsides[0] / nside)**2
        if p < 1:
            raise NotImplementedError("upsampling not implemented yet")
        indexes.append(np.unique(index // p).astype(int))
    return indexes
