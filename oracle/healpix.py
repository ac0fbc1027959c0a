"""Oracle restatement of the two healpy calls on the reference's graph-construction path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls `hp.pix2vec(nside, indexes, nest=True)` (deepsphere/utils.py:56) and
`hp.pixelfunc.get_all_neighbours(nside, indexes, nest=True)` (deepsphere/utils.py:63).
healpy 1.11.0 (environment.yml:29) wraps healpix_cxx; it is not vendored under
/root/reference and cannot be installed here, so its published algorithm (Gorski et al.
2005; healpix_cxx `T_Healpix_Base::pix2loc`, `::neighbors`, `nest2xyf`, `xyf2nest`) is
restated below, deliberately in scalar per-pixel form (the product uses an independent
vectorised implementation in deepsphere_tf1_b200/healpix.py; tests compare the two).

Parity: pinned by the known answers that exist outside this repository -- both `get_all_neighbours` docstring examples of
healpy, the five (theta, phi) pairs of its `pix2ang` docstring at Nside 16, the base-pixel centres and the Nside-2 children of
base pixel 0 from the ring formulas of Gorski et al. 2005 (tests/test_graph_golden.py) -- and by the geometric invariants of
SURVEY.md Appendix C; every other absolute value is "parity unpinned" (healpy cannot run here).
"""
import math

import numpy as np

# ring-number / phase tables of the 12 base faces (healpix_cxx jrll / jpll)
_JRLL = (2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4)
_JPLL = (1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7)

# neighbour order returned by healpy: SW, W, NW, N, NE, E, SE, S
_XOFF = (-1, -1, 0, 1, 1, 1, 0, -1)
_YOFF = (0, 1, 1, 1, 0, -1, -1, -1)

# rows nb = 0..8 : S, SE, E, SW, centre, NE, W, NW, N
_FACEARRAY = (
    (8, 9, 10, 11, -1, -1, -1, -1, 10, 11, 8, 9),
    (5, 6, 7, 4, 8, 9, 10, 11, 9, 10, 11, 8),
    (-1, -1, -1, -1, 5, 6, 7, 4, -1, -1, -1, -1),
    (4, 5, 6, 7, 11, 8, 9, 10, 11, 8, 9, 10),
    (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11),
    (1, 2, 3, 0, 0, 1, 2, 3, 5, 6, 7, 4),
    (-1, -1, -1, -1, 7, 4, 5, 6, -1, -1, -1, -1),
    (3, 0, 1, 2, 3, 0, 1, 2, 4, 5, 6, 7),
    (2, 3, 0, 1, -1, -1, -1, -1, 0, 1, 2, 3),
)
_SWAPARRAY = (
    (0, 0, 3), (0, 0, 6), (0, 0, 0), (0, 0, 5), (0, 0, 0),
    (5, 0, 0), (0, 0, 0), (6, 0, 0), (3, 0, 0),
)


def nside2npix(nside):
    """healpy.nside2npix (used at deepsphere/utils.py:263,448)."""
    return 12 * nside * nside


def _order(nside):
    order = int(nside).bit_length() - 1
    if nside < 1 or (1 << order) != nside:
        raise ValueError("nside must be a power of two")
    return order


def nest2xyf(nside, pix):
    """Nested pixel -> (ix, iy, face): de-interleave the Morton bits."""
    order = _order(nside)
    npface = nside * nside
    face = pix >> (2 * order)
    r = pix & (npface - 1)
    ix = iy = 0
    for b in range(order):
        ix |= ((r >> (2 * b)) & 1) << b
        iy |= ((r >> (2 * b + 1)) & 1) << b
    return ix, iy, face


def xyf2nest(nside, ix, iy, face):
    order = _order(nside)
    r = 0
    for b in range(order):
        r |= ((ix >> b) & 1) << (2 * b)
        r |= ((iy >> b) & 1) << (2 * b + 1)
    return (face << (2 * order)) + r


def pix2vec_one(nside, pix):
    """healpix_cxx pix2loc + pix2vec for one NESTED pixel, float64."""
    npix = nside2npix(nside)
    fact2 = 4.0 / npix
    fact1 = (nside << 1) * fact2
    ix, iy, face = nest2xyf(nside, pix)
    jr = _JRLL[face] * nside - ix - iy - 1
    have_sth = False
    sth = 0.0
    if jr < nside:
        nr = jr
        tmp = (nr * nr) * fact2
        z = 1.0 - tmp
        if z > 0.99:
            sth = math.sqrt(tmp * (2.0 - tmp))
            have_sth = True
    elif jr > 3 * nside:
        nr = 4 * nside - jr
        tmp = (nr * nr) * fact2
        z = tmp - 1.0
        if z < -0.99:
            sth = math.sqrt(tmp * (2.0 - tmp))
            have_sth = True
    else:
        nr = nside
        z = (2 * nside - jr) * fact1
    t = _JPLL[face] * nr + ix - iy
    if t < 0:
        t += 8 * nr
    halfpi = 0.5 * math.pi
    if nr == nside:
        phi = 0.75 * halfpi * t * fact1
    else:
        phi = (0.5 * halfpi * t) / nr
    if not have_sth:
        sth = math.sqrt((1.0 - z) * (1.0 + z))
    return sth * math.cos(phi), sth * math.sin(phi), z


def pix2vec(nside, ipix, nest=True):
    """healpy.pix2vec(nside, ipix, nest=True) -> (x, y, z) float64 arrays (utils.py:56)."""
    if not nest:
        raise NotImplementedError("the reference only uses nest=True (utils.py:40-41)")
    ipix = [int(i) for i in ipix]
    out = np.empty((3, len(ipix)), dtype=np.float64)
    for n, p in enumerate(ipix):
        out[:, n] = pix2vec_one(nside, p)
    return out[0], out[1], out[2]


def neighbours_one(nside, pix):
    """healpix_cxx T_Healpix_Base::neighbors for one NESTED pixel: 8 ints, -1 = missing."""
    ix, iy, face = nest2xyf(nside, pix)
    res = []
    for i in range(8):
        x = ix + _XOFF[i]
        y = iy + _YOFF[i]
        nb = 4
        if x < 0:
            x += nside
            nb -= 1
        elif x >= nside:
            x -= nside
            nb += 1
        if y < 0:
            y += nside
            nb -= 3
        elif y >= nside:
            y -= nside
            nb += 3
        f = _FACEARRAY[nb][face]
        if f < 0:
            res.append(-1)
            continue
        bits = _SWAPARRAY[nb][face >> 2]
        if bits & 1:
            x = nside - x - 1
        if bits & 2:
            y = nside - y - 1
        if bits & 4:
            x, y = y, x
        res.append(xyf2nest(nside, x, y, f))
    return res


def get_all_neighbours(nside, ipix, nest=True):
    """healpy.pixelfunc.get_all_neighbours(nside, ipix, nest=True) -> int array (8, n)
    in the order SW, W, NW, N, NE, E, SE, S (utils.py:63, flattened at utils.py:83)."""
    if not nest:
        raise NotImplementedError("the reference only uses nest=True (utils.py:40-41)")
    if np.isscalar(ipix):
        return np.array(neighbours_one(nside, int(ipix)), dtype=np.int64)
    ipix = [int(i) for i in ipix]
    out = np.empty((8, len(ipix)), dtype=np.int64)
    for n, p in enumerate(ipix):
        out[:, n] = neighbours_one(nside, p)
    return out
