"""HEALPix graph + rescaled Laplacian built ON THE DEVICE (csrc/healpix_build.cu).

`healpix_level(nside, indexes, device, ...)` does what `utils.healpix_laplacian(new=False)` + `utils.rescale_L` do on the host
(reference deepsphere/utils.py:25-133, 231-242, 320-342, 379-385, 423-425) and returns a ready `graph.DeviceGraph`: coordinates,
neighbours, distances, degrees, Laplacian entries and the ELL packing come from CUDA kernels that follow the numpy arithmetic
operation by operation; between them the host applies numpy's own `mean`, `exp` and `power` to the device results (their rounding
belongs to numpy's libm, and the weights are to stay bit-identical to the numpy path: tests/test_graph_build.py).  `lmax` comes from
a Lanczos iteration on the device (the reference calls ARPACK, utils.py:423) unless it is given.  The scipy matrix the reference's
API exposes (`model.L`) is assembled from the ELL arrays.  Full-sphere Nside = 1024 (12.58 M vertices): seconds instead of the
~220 s of the numpy path (26 s graph + 186 s ARPACK).
"""
import numpy as np
import torch
from scipy import sparse

from . import healpix as hp
from ._lib import call
from .graph import DeviceGraph


def lanczos_lmax(graph, steps=320, tol=2e-6, seed=0):
    """Largest eigenvalue of the symmetric ELL matrix of `graph` (fwd) by Lanczos with full re-orthogonalisation (two passes of
    classical Gram-Schmidt against all previous vectors); stops when the residual bound |beta_k s_k| of the largest Ritz pair
    falls under tol * lambda, or after `steps` steps (the Ritz value approaches lambda_max from below: the spectrum of a grid
    Laplacian is dense at its upper end, and the factor 1.02 of utils.py:423 absorbs what is left).  The sparse products run
    through libdsb200, the vector algebra through torch (set-up plumbing, not the hot path)."""
    from . import ops
    M = graph.M
    dev = graph.fwd[1].device
    steps = int(min(steps, M, max(16, (6 << 30) // (4 * M))))         # at most 6 GB of Lanczos vectors
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    Q = torch.empty((steps + 1, M), dtype=torch.float32, device=dev)
    q = torch.randn(M, generator=gen, device=dev, dtype=torch.float32)
    Q[0] = q / q.norm()
    alphas, betas = [], []
    lam = None
    for k in range(steps):
        w = ops.spmm(graph, Q[k].view(1, M, 1)).view(M)
        alphas.append(float(torch.dot(w.double(), Q[k].double())))
        for _ in range(2):
            w = w - Q[:k + 1].t() @ (Q[:k + 1] @ w)
        b = float(w.double().norm())
        T = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1)
        ev, evec = np.linalg.eigh(T)
        lam = float(ev[-1])
        if b < 1e-10 or (k >= 4 and abs(b * evec[-1, -1]) <= tol * abs(lam)):
            break
        betas.append(b)
        Q[k + 1] = w / b
    return lam


def healpix_level(nside, indexes=None, device=None, lap_type='normalized', std=None, lmax=None, rescale=True, with_scipy=True):
    """One level: returns (DeviceGraph, scipy COO float32 | None, lmax).  `indexes`: NESTED pixel list of a partial sphere."""
    device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    nside = int(nside)
    npix = hp.nside2npix(nside)
    ids = inv = None
    if indexes is None:
        M = npix
    else:
        indexes = np.asarray(indexes)
        M = len(indexes)
        if not M >= int(indexes.max()) + 1:                  # `usefast` of utils.py:46-50: otherwise the vertices are pixels 0..M-1
            ids = torch.from_numpy(indexes.astype(np.int64)).to(device)
            inv_h = np.full(npix, -1, dtype=np.int32)
            inv_h[indexes] = np.arange(M, dtype=np.int32)
            inv = torch.from_numpy(inv_h).to(device)
    coords = torch.empty((M, 3), dtype=torch.float32, device=device)
    call('hpx_coords', nside, ids, M, coords)
    nb = torch.empty((M, 8), dtype=torch.int32, device=device)
    d2 = torch.empty((M, 8), dtype=torch.float32, device=device)
    call('hpx_neighbours', nside, ids, inv, M, coords, nb, d2)
    # host: the two libm-dependent array operations of utils.py:107-111 on the device's distances
    d2h = d2.cpu().numpy()
    valid = nb.cpu().numpy() >= 0
    kernel_width = np.mean(d2h[valid]) if std is None else std
    wh = np.exp(-d2h / (2 * kernel_width))
    w = torch.from_numpy(np.ascontiguousarray(wh, dtype=np.float32)).to(device)
    scol = torch.empty((M, 8), dtype=torch.int32, device=device)
    sw = torch.empty((M, 8), dtype=torch.float32, device=device)
    cnt = torch.empty(M, dtype=torch.int32, device=device)
    deg = torch.empty(M, dtype=torch.float32, device=device)
    call('hpx_sort_degree', nb, w, M, scol, sw, cnt, deg)
    normalized = lap_type == 'normalized'
    if lap_type not in ('normalized', 'combinatorial'):
        raise ValueError('Unknown Laplacian type {}'.format(lap_type))
    d12 = None
    if normalized:                                           # utils.py:238: numpy's float32 power
        d12 = torch.from_numpy(np.power(deg.cpu().numpy(), -0.5).astype(np.float32)).to(device)
    width = int(cnt.max().item()) + 1
    col = torch.empty((M, width), dtype=torch.int32, device=device)
    val = torch.empty((M, width), dtype=torch.float32, device=device)
    call('hpx_laplacian', scol, sw, cnt, d12, deg, M, width, 1 if normalized else 0, col, val)
    nnz = int(cnt.sum().item()) + M
    g = DeviceGraph.from_ell(col, val, nnz, device, nested=(nside, indexes if ids is not None else None))
    if rescale:
        if lmax is None:
            lmax = np.float32(1.02 * lanczos_lmax(g))
        recip = 1.0 / (lmax / 2)                             # scipy: `L /= s` is `L.data *= 1.0 / s`, with lmax's own scalar type
        call('ell_rescale', col, val, M, width, float(np.float32(recip)))
        g = DeviceGraph.from_ell(col, val, nnz, device, nested=(nside, indexes if ids is not None else None))
    L = None
    if with_scipy:
        colh, valh, cnth = col.cpu().numpy(), val.cpu().numpy(), cnt.cpu().numpy()
        keep = np.arange(width)[None, :] <= cnth[:, None]
        rows = np.repeat(np.arange(M, dtype=np.int32), width).reshape(M, width)
        L = sparse.coo_matrix((valh[keep], (rows[keep], colh[keep])), shape=(M, M), dtype=np.float32)
    return g, L, lmax


def build_laplacians(nsides, indexes=None, device=None, std=None, lmax=None):
    """`utils.build_laplacians(nsides, indexes, sampling='healpix', new=False)` on the device: (L list of scipy COO, p, graphs)."""
    if indexes is None:
        indexes = [None] * len(nsides)
    if not isinstance(std, list):
        std = [std] * len(nsides)
    L, p, graphs = [], [], []
    nside_last = -1
    built = 0
    for i, (nside, index, sigma) in enumerate(zip(nsides, indexes, std)):
        if i > 0:
            p.append((nside_last // nside) ** 2)
        if nside == nside_last and i < len(nsides) - 1:
            L.append(L[-1].copy().tocoo())
            graphs.append(graphs[-1])
            continue
        nside_last = nside
        if i < len(nsides) - 1:
            g, Li, _ = healpix_level(nside, index, device, std=sigma, lmax=None if lmax is None else lmax[built])
            built += 1
            L.append(Li)
            graphs.append(g)
    return L, p, graphs
