"""Vectorised HEALPix NESTED geometry for the graph builder (host side, numpy).

Replaces the two healpy calls on the reference's graph-construction path:
`hp.pix2vec(nside, indexes, nest=True)` (reference deepsphere/utils.py:56) and
`hp.pixelfunc.get_all_neighbours(nside, indexes, nest=True)` (utils.py:63), plus
`hp.nside2npix` (utils.py:448).  Whole-array bit manipulation instead of per-pixel
calls, so the 12.58 M-pixel Nside=1024 sphere is handled in a few seconds.
"""
import numpy as np

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)

# direction d = 0..7 : SW, W, NW, N, NE, E, SE, S (healpy's output order)
_DX = np.array([-1, -1, 0, 1, 1, 1, 0, -1], dtype=np.int64)
_DY = np.array([0, 1, 1, 1, 0, -1, -1, -1], dtype=np.int64)

# face adjacency, indexed [border code][face]; border code = 4 + (x over/underflow) + 3*(y ...)
_FACE_NB = np.array([
    [8, 9, 10, 11, -1, -1, -1, -1, 10, 11, 8, 9],
    [5, 6, 7, 4, 8, 9, 10, 11, 9, 10, 11, 8],
    [-1, -1, -1, -1, 5, 6, 7, 4, -1, -1, -1, -1],
    [4, 5, 6, 7, 11, 8, 9, 10, 11, 8, 9, 10],
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11],
    [1, 2, 3, 0, 0, 1, 2, 3, 5, 6, 7, 4],
    [-1, -1, -1, -1, 7, 4, 5, 6, -1, -1, -1, -1],
    [3, 0, 1, 2, 3, 0, 1, 2, 4, 5, 6, 7],
    [2, 3, 0, 1, -1, -1, -1, -1, 0, 1, 2, 3]], dtype=np.int64)
# coordinate fix-up bits when crossing to another face, indexed [border code][face // 4]
_SWAP = np.array([[0, 0, 3], [0, 0, 6], [0, 0, 0], [0, 0, 5], [0, 0, 0],
                  [5, 0, 0], [0, 0, 0], [6, 0, 0], [3, 0, 0]], dtype=np.int64)


def nside2npix(nside):
    return 12 * int(nside) * int(nside)


def _log2(nside):
    nside = int(nside)
    order = nside.bit_length() - 1
    if nside < 1 or (1 << order) != nside:
        raise ValueError('nside must be a power of 2')
    return order


def _compact_even_bits(v):
    """Keep bits 0,2,4,... of v and pack them (inverse of Morton spreading), int64 arrays."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def _spread_bits(v):
    """Insert a zero bit between consecutive bits of v (Morton spreading)."""
    v = v & 0x00000000FFFFFFFF
    v = (v | (v << 16)) & 0x0000FFFF0000FFFF
    v = (v | (v << 8)) & 0x00FF00FF00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v << 2)) & 0x3333333333333333
    v = (v | (v << 1)) & 0x5555555555555555
    return v


def nest2xyf(nside, pix):
    order = _log2(nside)
    pix = np.asarray(pix, dtype=np.int64)
    face = pix >> (2 * order)
    r = pix & (nside * nside - 1)
    return _compact_even_bits(r), _compact_even_bits(r >> 1), face


def xyf2nest(nside, ix, iy, face):
    order = _log2(nside)
    return (face << (2 * order)) + _spread_bits(ix) + (_spread_bits(iy) << 1)


def pix2vec(nside, ipix, nest=True):
    """Unit vectors (float64) of NESTED pixel centres; returns (x, y, z) like healpy."""
    if not nest:
        raise NotImplementedError('only nest=True is supported (as in the reference)')
    nside = int(nside)
    ix, iy, face = nest2xyf(nside, np.asarray(ipix, dtype=np.int64))
    npix = nside2npix(nside)
    fact2 = 4.0 / npix
    fact1 = (nside << 1) * fact2
    jr = _JRLL[face] * nside - ix - iy - 1
    north = jr < nside
    south = jr > 3 * nside
    nr = np.where(north, jr, np.where(south, 4 * nside - jr, nside))
    cap = (nr * nr) * fact2                      # 1 - |z| in the polar caps
    z = np.where(north, 1.0 - cap, np.where(south, cap - 1.0, (2 * nside - jr) * fact1))
    t = _JPLL[face] * nr + ix - iy
    t = np.where(t < 0, t + 8 * nr, t)
    halfpi = 0.5 * np.pi
    phi = np.where(nr == nside, 0.75 * halfpi * t * fact1, (0.5 * halfpi * t) / nr)
    # healpix_cxx switches to the cancellation-free form of sin(theta) close to the poles
    near_pole = (north & (z > 0.99)) | (south & (z < -0.99))
    sth = np.where(near_pole, np.sqrt(cap * (2.0 - cap)), np.sqrt((1.0 - z) * (1.0 + z)))
    return sth * np.cos(phi), sth * np.sin(phi), z


def get_all_neighbours(nside, ipix, nest=True):
    """8 neighbours (SW, W, NW, N, NE, E, SE, S) of NESTED pixels, shape (8, n); -1 = none."""
    if not nest:
        raise NotImplementedError('only nest=True is supported (as in the reference)')
    nside = int(nside)
    scalar = np.isscalar(ipix)
    ix, iy, face = nest2xyf(nside, np.atleast_1d(np.asarray(ipix, dtype=np.int64)))
    out = np.empty((8, ix.shape[0]), dtype=np.int64)
    for d in range(8):
        x = ix + _DX[d]
        y = iy + _DY[d]
        code = np.full(x.shape, 4, dtype=np.int64)
        lo, hi = x < 0, x >= nside
        x = np.where(lo, x + nside, np.where(hi, x - nside, x))
        code += hi.astype(np.int64) - lo.astype(np.int64)
        lo, hi = y < 0, y >= nside
        y = np.where(lo, y + nside, np.where(hi, y - nside, y))
        code += 3 * (hi.astype(np.int64) - lo.astype(np.int64))
        f = _FACE_NB[code, face]
        bits = _SWAP[code, face >> 2]
        x = np.where(bits & 1, nside - x - 1, x)
        y = np.where(bits & 2, nside - y - 1, y)
        swap = (bits & 4) != 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        out[d] = np.where(f >= 0, xyf2nest(nside, x, y, np.maximum(f, 0)), -1)
    return out[:, 0] if scalar else out
