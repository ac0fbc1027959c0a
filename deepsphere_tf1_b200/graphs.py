"""Icosahedral, kNN-HEALPix and irregular (station) graphs: the graphs the reference gets from a PyGSP fork.

The reference builds these through `pygsp.graphs.{SphereIcosahedron, SphereHealpix, NNGraph}` of the fork
`Droxef/pygsp@new_sphere_graph` (reference deepsphere/utils.py:16, 331-334, 370-376; experiments/ghcn/
GHCN_preprocessing.py:398-416), which is not vendored and not installable here.  What IS in the reference repo is the
notebook the classes were developed in, notebooks/sphere_to_graph.ipynb: `SphereIcosahedron` (raw lines 962-1124: the 12 base
vertices, the `_upward` rotation, the per-level `divide()` that APPENDS the edge mid-points behind the existing vertices --
so the first 10 * 4^(l-1) + 2 vertices of level l are level l - 1, which is what the reference's icosahedral pooling
`x[:, :p, :]` (models.py:1137-1138) relies on -- and the re-projection onto the sphere), `SphereHealpixNN` (raw lines 71-97:
kNN graph on the pixel centres with the kernel-width table) and the `NNGraph` constructor calls.  The mesh part is restated
here statement by statement and pinned: tests/golden/make_golden.py executes the notebook's own class source and
tests/test_graph_golden.py compares vertices and faces bit for bit.  `NNGraph` itself (kNN search, Gaussian kernel,
symmetrisation) lives in PyGSP; `nn_graph` restates the published PyGSP 0.5.1 algorithm (pygsp/graphs/nngraphs/nngraph.py) --
parity unpinned, checked through invariants only.
"""
import collections

import numpy as np
from scipy import sparse, spatial

from . import healpix as hp


# ---------------------------------------------------------------------------------------------- PyGSP NNGraph
def nn_graph(coords, k=10, sigma=None, center=True, rescale=True, symmetrize='average', dtype=np.float32):
    """Weight matrix (CSR) of `pygsp.graphs.NNGraph(coords, NNtype='knn', k=k, sigma=sigma, center=..., rescale=...)`:
    optional centring / rescaling of the points, k nearest neighbours (Euclidean, kd-tree), `w = exp(-d^2 / sigma)` with
    sigma = the mean neighbour distance unless given, and symmetrisation `(W + W^T) / 2` ('average') or `max(W, W^T)`."""
    X = np.asarray(coords, dtype=np.float64)
    N, d = X.shape
    if center:
        X = X - np.mean(X, axis=0)
    if rescale:
        bounding_radius = 0.5 * np.linalg.norm(np.amax(X, axis=0) - np.amin(X, axis=0), 2)
        scale = np.power(N, 1. / float(min(d, 3))) / 10.
        X = X * (scale / bounding_radius)
    k = int(min(k, N - 1))
    D, NN = spatial.cKDTree(X).query(X, k=k + 1, p=2)
    if sigma is None:
        sigma = np.mean(D[:, 1:])
    rows = np.repeat(np.arange(N), k)
    cols = NN[:, 1:].reshape(-1)
    vals = np.exp(-np.power(D[:, 1:].reshape(-1), 2) / float(sigma))
    W = sparse.csc_matrix((vals, (rows, cols)), shape=(N, N))
    if symmetrize == 'average':
        W = (W + W.T) / 2.
    elif symmetrize == 'maximum':
        W = W.maximum(W.T)
    else:
        raise ValueError('unknown symmetrisation ' + str(symmetrize))
    W = sparse.csr_matrix(W, dtype=dtype)
    W.setdiag(0)
    W.eliminate_zeros()
    return W


def laplacian(W, lap_type='combinatorial', dtype=np.float32):
    """`Graph.compute_laplacian` of PyGSP for undirected graphs: D - W, or I - D^-1/2 W D^-1/2."""
    W = sparse.csr_matrix(W, dtype=np.float64)
    dw = np.asarray(W.sum(axis=1)).ravel()
    if lap_type == 'combinatorial':
        L = sparse.diags(dw) - W
    elif lap_type == 'normalized':
        d = np.zeros_like(dw)
        d[dw > 0] = np.power(dw[dw > 0], -0.5)
        Dh = sparse.diags(d)
        L = sparse.identity(W.shape[0]) - Dh @ W @ Dh
    else:
        raise ValueError('Unknown Laplacian type ' + str(lap_type))
    return sparse.csr_matrix(L, dtype=dtype)


# ---------------------------------------------------------------------------------------------- icosahedral mesh
def _decimal_to_digits(decimal, min_digits=None):
    digits = abs(int(np.log10(decimal)))
    if min_digits is not None:
        digits = np.clip(digits, min_digits, 20)
    return digits


def _unique_rows(data):
    """First occurrence of every distinct row after rounding to 8 decimals, and the inverse map, with the notebook's own
    hashing (unique_rows / hashable_rows / float_to_int, raw lines 827-936): the ORDER of the unique rows -- ascending hash --
    is the order in which `divide()` appends the new vertices."""
    digits = _decimal_to_digits(1e-8)
    data_max = np.abs(data).max() * 10 ** digits
    dt = [np.int32, np.int64][int(data_max > 2 ** 31)]
    as_int = np.round((data * 10 ** digits) - 1e-6).astype(dt)
    precision = int(np.floor(64 / as_int.shape[1]))
    if np.abs(as_int).max() < 2 ** (precision - 1):
        hashes = np.zeros(len(as_int), dtype=np.int64)
        for offset, column in enumerate(as_int.astype(np.int64).T):
            np.bitwise_xor(hashes, column << (offset * precision), out=hashes)
    else:
        hashes = np.ascontiguousarray(as_int).view(np.dtype((np.void, as_int.dtype.itemsize * as_int.shape[1]))).reshape(-1)
    _, unique, inverse = np.unique(hashes, return_index=True, return_inverse=True)
    return unique, inverse


def _rot_matrix(rot_axis, cos_t, sin_t):
    k = rot_axis / np.linalg.norm(rot_axis)
    eye = np.eye(3)
    R = []
    for i in range(3):
        v = eye[i]
        R.append(v * cos_t + np.cross(k, v) * sin_t + k * (k.dot(v)) * (1 - cos_t))
    return np.stack(R, axis=-1)


def _upward(V, F, ind=11):
    """Rotate the icosahedron so that vertex `ind` points to +z and one of its neighbours to +y (raw lines 1076-1096)."""
    V0 = V[ind]
    Z0 = np.array([0, 0, 1])
    k = np.cross(V0, Z0)
    ct = np.dot(V0, Z0)
    st = -np.linalg.norm(k)
    V = V.dot(_rot_matrix(k, ct, st))
    FF = np.unique(np.concatenate([F[i] for i in range(F.shape[0]) if ind in F[i]]))
    ni = [f for f in FF if f != ind][0]
    vec = V[ni].copy()
    vec[2] = 0
    vec = vec / np.linalg.norm(vec)
    y_ = np.eye(3)[1]
    k = np.eye(3)[2]
    crs = np.cross(vec, y_)
    ct = -np.dot(vec, y_)
    st = -np.sign(crs[-1]) * np.linalg.norm(crs)
    return V.dot(_rot_matrix(k, ct, st))


_ICO_FACES = [1, 2, 7, 1, 7, 10, 1, 10, 9, 1, 9, 5, 1, 5, 2, 2, 7, 12, 12, 7, 8, 7, 8, 10, 8, 10, 3, 10, 3, 9, 3, 9, 6, 9, 6, 5, 6,
              5, 11, 5, 11, 2, 11, 2, 12, 4, 11, 12, 4, 12, 8, 4, 8, 3, 4, 3, 6, 4, 6, 11]


def icosahedron_mesh(level):
    """Vertices [10 * 4^level + 2, 3] (float64, unit norm) and faces [20 * 4^level, 3] of `SphereIcosahedron(level)`
    (notebook raw lines 962-1075).  Vertices of level l - 1 are a prefix of level l."""
    PHI = (1 + np.sqrt(5)) / 2
    radius = np.sqrt(PHI ** 2 + 1)
    coords = np.zeros((12, 3))
    pts = [collections.deque(p) for p in ([0, 1, PHI], [0, -1, PHI], [0, 1, -PHI], [0, -1, -PHI])]
    for i in range(3):
        for j in range(4):
            coords[4 * i + j] = pts[j]
            pts[j].rotate()
    coords = coords / radius
    faces = np.reshape(_ICO_FACES, (20, 3)) - 1
    coords = _upward(coords, faces)
    for _ in range(level):
        # divide(): the mid-point of every edge becomes a vertex, every triangle becomes four
        triangles = coords[faces]
        mid = np.vstack([triangles[:, g, :].mean(axis=1) for g in [[0, 1], [1, 2], [2, 0]]])
        mid_idx = (np.arange(len(faces) * 3)).reshape((3, -1)).T
        unique, inverse = _unique_rows(mid)
        mid = mid[unique]
        mid_idx = inverse[mid_idx] + len(coords)
        f = np.column_stack([faces[:, 0], mid_idx[:, 0], mid_idx[:, 2],
                             mid_idx[:, 0], faces[:, 1], mid_idx[:, 1],
                             mid_idx[:, 2], mid_idx[:, 1], faces[:, 2],
                             mid_idx[:, 0], mid_idx[:, 1], mid_idx[:, 2]]).reshape((-1, 3))
        new_faces = np.vstack((faces, f[len(faces):]))
        new_faces[np.arange(len(faces))] = f[:len(faces)]
        coords = np.vstack((coords, mid))
        faces = new_faces
        # normalize(): back onto the unit sphere
        scalar = (coords ** 2).sum(axis=1) ** .5
        unit = coords / scalar.reshape((-1, 1))
        coords = coords + unit * (1 - scalar).reshape((-1, 1))
    return coords, faces


def icosahedron_weights(level, dtype=np.float32):
    """`SphereIcosahedron(level).W`: NNGraph on the mesh vertices with k = 6 (5 at level 0) and PyGSP's defaults."""
    coords, _ = icosahedron_mesh(level)
    return nn_graph(coords, k=5 if level == 0 else 6, dtype=dtype)


def icosahedron_laplacian(order=0, lap_type='combinatorial', indexes=None, dtype=np.float32):
    """reference deepsphere/utils.py:370-376."""
    return laplacian(icosahedron_weights(order, dtype=np.float64), lap_type, dtype)


# ---------------------------------------------------------------------------------------------- kNN HEALPix graph (new=True)
HEALPIX_NN_SIGMA = {1: 0.5, 2: 0.15, 4: 0.05, 8: 0.0125, 16: 0.005, 32: 0.001}      # notebook raw line 82


def healpix_nn_weights(nside, indexes=None, n_neighbors=8, dtype=np.float32):
    """`SphereHealpix(nside, indexes, nest=True, n_neighbors)` of the fork (utils.py:331-334, the `new=True` default):
    kNN graph on the float32 pixel centres, kernel width from the notebook's table (0.001 beyond nside 32), no centring /
    rescaling (notebook raw lines 71-97).  6 neighbours at nside 1."""
    npix = hp.nside2npix(nside)
    indexes = np.arange(npix) if indexes is None else np.asarray(indexes)
    x, y, z = hp.pix2vec(nside, indexes, nest=True)
    coords = np.asarray(np.vstack([x, y, z]).transpose(), dtype=np.float32)
    k = 6 if nside == 1 else n_neighbors
    sigma = HEALPIX_NN_SIGMA.get(int(nside), 0.001)
    return nn_graph(coords, k=k, sigma=sigma, center=False, rescale=False, dtype=dtype)


def healpix_nn_laplacian(nside, indexes=None, n_neighbors=8, lap_type='normalized', dtype=np.float32):
    return laplacian(healpix_nn_weights(nside, indexes, n_neighbors, np.float64), lap_type, dtype)


# ---------------------------------------------------------------------------------------------- irregular station graph (GHCN)
def sphere_knn_weights(lon, lat, neighbors, rad=True, dtype=np.float32):
    """`sphereGraph(phi=lon, theta=lat, neighbors)` of experiments/ghcn/GHCN_preprocessing.py:398-416: unit vectors from
    longitude / latitude (theta shifted by -pi/2 as upstream), kNN graph without centring / rescaling."""
    phi, theta = np.array(lon, dtype=np.float64), np.array(lat, dtype=np.float64)
    if not rad:
        theta, phi = np.deg2rad(theta), np.deg2rad(phi)
    theta = theta - np.pi / 2
    ct, st = np.cos(theta).flatten(), np.sin(theta).flatten()
    cp, sp = np.cos(phi).flatten(), np.sin(phi).flatten()
    coords = np.vstack([st * cp, st * sp, ct]).T
    return nn_graph(coords, k=neighbors, center=False, rescale=False, dtype=dtype), coords


def sphere_knn_laplacian(lon, lat, neighbors, rad=True, lap_type='combinatorial', dtype=np.float32):
    W, _ = sphere_knn_weights(lon, lat, neighbors, rad, np.float64)
    return laplacian(W, lap_type, dtype)
