"""Shared helpers for the parity tests: small graphs and oracle <-> product model pairs."""
import numpy as np
import torch
from scipy import sparse

from deepsphere_tf1_b200 import utils as putils


def healpix_L(nside, order=0):
    """Rescaled fp32 Laplacian of the in-repo 8-neighbour HEALPix graph (full or partial sphere)."""
    idx = putils.nside2indexes([nside], order)[0]
    L = putils.healpix_laplacian(nside=nside, indexes=idx, new=False)
    return putils.rescale_L(L, lmax=putils.estimate_lmax(L)).tocoo()


def random_knn_L(n, k, seed=0, rescale=True):
    """Irregular-degree symmetric kNN graph on random sphere points (GHCN-like), combinatorial L."""
    rs = np.random.RandomState(seed)
    pts = rs.randn(n, 3)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    nn = np.argsort(d2, axis=1)[:, 1:k + 1]
    rows = np.repeat(np.arange(n), k)
    w = np.exp(-d2[rows, nn.ravel()] / (2 * d2[rows, nn.ravel()].mean()))
    W = sparse.csr_matrix((w, (rows, nn.ravel())), shape=(n, n))
    W = W.maximum(W.T).astype(np.float32)                       # symmetrise -> irregular degrees
    L = putils.build_laplacian(W, 'combinatorial').astype(np.float32)
    if rescale:
        L = putils.rescale_L(sparse.csr_matrix(L), lmax=putils.estimate_lmax(L))
    return sparse.coo_matrix(L, dtype=np.float32)


def random_L(n, deg, seed=0):
    """Random NON-symmetric sparse matrix (tests that the backward really uses L^T)."""
    rs = np.random.RandomState(seed)
    rows = np.repeat(np.arange(n), deg)
    cols = rs.randint(0, n, size=n * deg)
    vals = (rs.randn(n * deg) / deg).astype(np.float32)
    A = sparse.csr_matrix((vals, (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    return A.tocoo()


def copy_weights(oracle, model):
    """Give the product model the oracle's parameters (by reference variable name)."""
    assert list(oracle.params.keys()) == [s[0] for s in model._P.specs], \
        (list(oracle.params.keys()), [s[0] for s in model._P.specs])
    for name, v in oracle.params.items():
        model.set_var(name, v.detach().numpy())
    for name, v in oracle.state.items():
        model.set_var(name, v.detach().numpy())


def assert_close(actual, expected, rtol, name='', floor=0.0):
    """Max error relative to the largest magnitude of the expected tensor (robust to near-zero
    entries); `floor` is a lower bound of that scale for tensors that are numerically zero."""
    a = actual.detach().cpu().double().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual, np.float64)
    e = expected.detach().cpu().double().numpy() if isinstance(expected, torch.Tensor) else np.asarray(expected, np.float64)
    assert a.shape == e.shape, (name, a.shape, e.shape)
    scale = max(np.abs(e).max(), floor, 1e-30)
    err = np.abs(a - e).max() / scale
    assert err <= rtol, '{}: max error {:.3e} relative to max |expected| {:.3e} (rtol {})'.format(name, err, scale, rtol)


def free_port():
    """A TCP port nobody listens on (for torchrun's rendezvous)."""
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port
