"""Device-side packing of the rescaled Laplacians for the sparse products of the hot path.

The reference turns each scipy Laplacian into a `tf.SparseTensor` constant, reordered row-major
and cast to float32 (deepsphere/models.py:1044-1047).  Here each level becomes a `DeviceGraph`:
fixed-width ELL (HEALPix 8-neighbour / icosahedral graphs: <= 9 / 7 entries per row) or CSR
(irregular kNN graphs), together with the same packing of L^T for the backward pass (L is
symmetric only up to 1 ulp because of `D12 * W * D12`, reference utils.py:240).
"""
import numpy as np
import torch
from scipy import sparse

ELL_MAX_WIDTH = 16
TILE_SLOTS = 10            # neighbour slots of a row record of the tiled kernels (csrc/sparse_tiled.cu: RowRec)
TILE_ROWS = 256            # rows per tile: 16 x 16 pixels in HEALPix nested order
TILE_SMEM_ROWS = 1600      # tile + ring 1 + tile + ring 1 + ring 2 rows that fit shared memory with >= 2 CTAs per SM


class TilePlan(object):
    """Row tiles with two halo rings for the shared-memory tiled recurrence kernels.

    Tile t owns the consecutive rows [t*TR, (t+1)*TR).  Ring 1 = rows referenced by the tile's rows that lie
    outside the tile, ring 2 = rows referenced by ring 1 outside tile + ring 1 (both sorted by global id).
    LOCAL indices: tile rows 0..nt-1, ring 1 nt..n1-1, ring 2 n1..n2-1.  Every row of tile + ring 1 gets one
    64-byte record: 10 fp32 values + 10 uint16 local columns (unused slots: value 0, own index)."""

    def __init__(self, col, val, device, TR=TILE_ROWS):
        M, W = col.shape
        self.valid = False
        if W > TILE_SLOTS or M == 0:
            return
        T = (M + TR - 1) // TR
        info = np.zeros((T, 4), np.int32)
        vals, lcols, exts = [], [], []
        soff = eoff = 0
        n1cap = n2cap = 0
        for t in range(T):
            r0, r1 = t * TR, min(M, (t + 1) * TR)
            nt = r1 - r0
            c_t = col[r0:r1]
            u = np.unique(c_t)
            ring1 = u[(u < r0) | (u >= r1)]
            c_1 = col[ring1]
            u2 = np.unique(c_1)
            u2 = u2[(u2 < r0) | (u2 >= r1)]
            ring2 = np.setdiff1d(u2, ring1, assume_unique=True)
            h1, h2 = len(ring1), len(ring2)
            n1, n2 = nt + h1, nt + h1 + h2
            if 2 * n1 + h2 > TILE_SMEM_ROWS or n2 > 65535:
                return                                   # vertex order without locality: keep the gather kernels
            allc = np.concatenate([c_t, c_1], 0).astype(np.int64)
            loc = allc - r0
            out = (allc < r0) | (allc >= r1)
            g = allc[out]
            if len(g):
                i1 = np.minimum(np.searchsorted(ring1, g), max(h1 - 1, 0))
                in1 = ring1[i1] == g
                i2 = np.minimum(np.searchsorted(ring2, g), max(h2 - 1, 0))
                if h2:
                    assert np.all(in1 | (ring2[i2] == g))
                else:
                    assert np.all(in1)
                loc[out] = np.where(in1, nt + i1, nt + h1 + i2)
            v = np.zeros((n1, TILE_SLOTS), np.float32)
            v[:, :W] = np.concatenate([val[r0:r1], val[ring1]], 0)
            lc = np.repeat(np.arange(n1, dtype=np.uint16)[:, None], TILE_SLOTS, axis=1)
            lc[:, :W] = loc.astype(np.uint16)
            vals.append(v)
            lcols.append(lc)
            exts.append(ring1.astype(np.int32))
            exts.append(ring2.astype(np.int32))
            info[t] = (soff, n1, n2, eoff)
            soff += n1
            eoff += h1 + h2
            n1cap, n2cap = max(n1cap, n1), max(n2cap, n2)
        recs = np.zeros((soff, 64), np.uint8)
        recs[:, :40] = np.concatenate(vals, 0).view(np.uint8).reshape(soff, 40)
        recs[:, 40:60] = np.concatenate(lcols, 0).view(np.uint8).reshape(soff, 20)
        ext = np.concatenate(exts + [np.zeros(1, np.int32)])
        self.info = torch.from_numpy(info).to(device)
        self.recs = torch.from_numpy(recs).to(device)
        self.ext = torch.from_numpy(ext).to(device)
        self.ntiles, self.TR, self.n1cap, self.n2cap, self.width = T, TR, n1cap, n2cap, W
        self.valid = True

    def args(self):
        return (self.info, self.recs, self.ext, self.ntiles, self.TR, self.n1cap, self.n2cap, self.width)


class DeviceGraph(object):
    """One Laplacian level on the GPU.  `.fwd` / `.bwd` = (rowptr | None, col, val, width)."""

    def __init__(self, L, device, force_csr=False, tiled=True):
        L = sparse.csr_matrix(L, dtype=np.float32)
        L.sum_duplicates()
        L.sort_indices()                                   # row-major, ascending columns (tf.sparse_reorder)
        if L.shape[0] != L.shape[1]:
            raise ValueError('Laplacian must be square')
        self.M = L.shape[0]
        self.nnz = int(L.nnz)
        self.device = device
        counts = np.diff(L.indptr)
        self.max_degree = int(counts.max()) if self.M else 0
        self.is_ell = (not force_csr) and self.max_degree <= ELL_MAX_WIDTH
        self._want_tiles = tiled
        self.fwd_tiles = self.bwd_tiles = None
        self.fwd = self._pack(L, 'fwd')
        LT = sparse.csr_matrix(L.T)
        LT.sort_indices()
        self.symmetric_pattern = (LT.nnz == L.nnz and np.array_equal(LT.indptr, L.indptr)
                                  and np.array_equal(LT.indices, L.indices))
        if self.symmetric_pattern and np.array_equal(LT.data, L.data):
            self.bwd = self.fwd
            self.bwd_tiles = self.fwd_tiles
        else:
            self.bwd = self._pack(LT, 'bwd')

    def _pack(self, A, which='fwd'):
        if self.is_ell:
            width = max(self.max_degree, int(np.diff(A.indptr).max()))
            # padding slots: (col = row, val = 0) -- a harmless gather, so kernels need no early exit
            col = np.repeat(np.arange(self.M, dtype=np.int32)[:, None], width, axis=1)
            val = np.zeros((self.M, width), dtype=np.float32)
            counts = np.diff(A.indptr)
            rows = np.repeat(np.arange(self.M), counts)
            slot = np.arange(A.nnz) - np.repeat(A.indptr[:-1], counts)
            col[rows, slot] = A.indices
            val[rows, slot] = A.data
            if self._want_tiles:
                plan = TilePlan(col, val, self.device)
                setattr(self, which + '_tiles', plan if plan.valid else None)
            return (None, torch.from_numpy(col).to(self.device), torch.from_numpy(val).to(self.device), width)
        return (torch.from_numpy(A.indptr.astype(np.int32)).to(self.device),
                torch.from_numpy(A.indices.astype(np.int32)).to(self.device),
                torch.from_numpy(A.data.astype(np.float32)).to(self.device), 0)

    def bytes_stored(self):
        """Bytes one sparse product reads for the matrix (8 per stored entry)."""
        return 8 * self.nnz + (0 if self.is_ell else 4 * (self.M + 1))
