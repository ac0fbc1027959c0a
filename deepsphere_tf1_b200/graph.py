"""Device-side packing of the rescaled Laplacians for the sparse products of the hot path.

The reference turns each scipy Laplacian into a `tf.SparseTensor` constant, reordered row-major
and cast to float32 (deepsphere/models.py:1044-1047).  Here each level becomes a `DeviceGraph`:
fixed-width ELL (HEALPix 8-neighbour / icosahedral graphs: <= 9 / 7 entries per row) or CSR
(irregular kNN graphs), together with the same packing of L^T for the backward pass (L is
symmetric only up to 1 ulp because of `D12 * W * D12`, reference utils.py:240).
"""
import os

import numpy as np
import torch
from scipy import sparse

ELL_MAX_WIDTH = 16


PIPE_TILE_BYTES = 8192     # x bytes of one row tile of the pipelined kernel (csrc/sparse_pipe.cu): TR = 8192 / (4 F) rows
PIPE_SMEM = 113 * 1024


class PipePlan(object):
    """Row tiles + ring-1 halo lists for the bulk-asynchronous pipelined sparse product (csrc/sparse_pipe.cu).

    Tile t owns the consecutive rows [t*TR, (t+1)*TR).  LOCAL column of an entry: col - t*TR if the column lies in
    the tile, else TR + (rank of the column among the tile's distinct outside columns, ascending).
      rec int32[T, W, TR, 2] = (local column, value bits), slot-major per tile (padding rows / slots: (row, 0))
      ext int32[T, HC]        = global row of every halo slot (unused slots: the tile's first row)
      nh  int32[T]            = halo rows of the tile
    Built from the ELL packing itself, so it works for any graph whose vertex order has locality; `valid` is False
    when the halos are too large for shared memory (random vertex order): the gather kernels stay in charge."""

    def __init__(self, col, val, device, TR):
        M, W = col.shape
        self.valid = False
        if M == 0 or W not in (7, 9) or TR < 8:
            return
        T = (M + TR - 1) // TR
        tile = (np.arange(M, dtype=np.int64) // TR)
        c64 = col.astype(np.int64)
        outside = (c64 // TR) != tile[:, None]
        ri, si = np.nonzero(outside)
        keys = tile[ri] * M + c64[ri, si]
        uk, inv = np.unique(keys, return_inverse=True)
        ut, ug = uk // M, uk % M
        counts = np.bincount(ut, minlength=T).astype(np.int64)
        HC = int(counts.max()) if len(uk) else 0
        HC = max(8, (HC + 7) // 8 * 8)
        if HC > 2 * TR:
            return
        start = np.cumsum(counts) - counts
        rank = np.arange(len(uk), dtype=np.int64) - start[ut]
        lcol = c64 - tile[:, None] * TR
        lcol[ri, si] = TR + rank[inv]
        ext = np.repeat((np.arange(T, dtype=np.int32) * TR)[:, None], HC, axis=1)
        ext[ut, rank] = ug.astype(np.int32)
        Mp = T * TR
        rec = np.zeros((Mp, W, 2), dtype=np.int32)
        rec[:, :, 0] = (np.arange(Mp) % TR)[:, None]                     # padding rows point at themselves
        rec[:M, :, 0] = lcol
        rec[:M, :, 1] = val.view(np.int32)
        rec = np.ascontiguousarray(rec.reshape(T, TR, W, 2).transpose(0, 2, 1, 3))
        self.rec_np, self.ext_np, self.nh_np = rec, ext, counts.astype(np.int32)
        self.dense_np = self._dense_groups(lcol, val, M, Mp, T, TR, W)
        self.T, self.TR, self.HC, self.width, self.M = T, TR, HC, W, M
        if device is not None:
            self.dense = torch.from_numpy(self.dense_np).to(device) if self.dense_np is not None else None
            self.rec = torch.from_numpy(rec).to(device)
            self.ext = torch.from_numpy(np.ascontiguousarray(ext)).to(device)
            self.nh = torch.from_numpy(self.nh_np).to(device)
        self.valid = True

    @staticmethod
    def _dense_groups(lcol, val, M, Mp, T, TR, W):
        """Dense-group records (include/dsb200.h: dsb200_spmm_pipe, dense != 0), int32[T, 18, TR/4, 4], or None when
        some group of 4 consecutive rows touches more than 16 distinct columns or the layout does not apply."""
        if W != 9 or TR % 4 or TR + 2 * TR > 65535:
            return None
        G = TR // 4
        # Padding rows (>= M, weights 0) point at the FIRST row of their group, never at themselves: the kernel stages only
        # the tile's M - t*TR real rows, and a group that mixes real and padding rows would otherwise multiply whatever the
        # shared-memory stage held before (uninitialised in a CTA's first item: possibly NaN bits) by 0.0 into its real rows.
        lc = np.repeat(((np.arange(Mp, dtype=np.int64) // 4 * 4) % TR)[:, None], W, axis=1)
        lc[:M] = lcol
        v = np.zeros((Mp, W), np.float32)
        v[:M] = val
        ng = Mp // 4
        lcg = lc.reshape(ng, 4 * W)
        order = np.argsort(lcg, axis=1, kind='stable')
        srt = np.take_along_axis(lcg, order, axis=1)
        first = np.ones_like(srt, dtype=bool)
        first[:, 1:] = srt[:, 1:] != srt[:, :-1]
        rank_sorted = np.cumsum(first, axis=1) - 1
        if int(rank_sorted[:, -1].max()) + 1 > 16:
            return None
        rank = np.empty_like(rank_sorted)
        np.put_along_axis(rank, order, rank_sorted, axis=1)
        # Slot of every distinct column.  Shared-memory rows of 64 bytes (F = 16) sit in one HALF of a 128-byte bank line,
        # chosen by the parity of the local column; the 8 groups a warp works on read slot u together, so slot u of group g
        # gets a column of parity (u + g) & 1 where the group has at most 8 columns of each parity (always, inside the
        # sphere: a 4 x 4 block): 4 + 4 segments per instruction, no bank conflict.  Other groups: ascending order.
        gpar = (np.arange(ng) & 1)[:, None]
        par = srt & 1
        k_in_class = np.where(par == 1, np.cumsum(first & (par == 1), axis=1), np.cumsum(first & (par == 0), axis=1)) - 1
        n_even = (first & (par == 0)).sum(axis=1)
        n_odd = (first & (par == 1)).sum(axis=1)
        balanced = ((n_even <= 8) & (n_odd <= 8))[:, None]
        slot_sorted = np.where(balanced, 2 * k_in_class + (par ^ gpar), rank_sorted)
        slot = np.empty_like(slot_sorted)
        np.put_along_axis(slot, order, slot_sorted, axis=1)
        cols = np.repeat(lcg[:, :1], 16, axis=1)                                      # unused slots: the group's first row
        gi = np.repeat(np.arange(ng), 4 * W)
        cols[gi, slot.ravel()] = lcg.ravel()
        # row r of the stored block = row r ^ (g & 1) of the group: the y / z reads and the stores of an instruction then
        # touch both halves of the bank lines / both 64-byte halves of the global lines (include/dsb200.h)
        wd = np.zeros((ng, 4, 16), np.float32)
        ri = np.tile(np.repeat(np.arange(4), W), ng) ^ np.repeat(np.arange(ng) & 1, 4 * W)
        np.add.at(wd, (gi, ri, slot.ravel()), v.reshape(ng, 4 * W).ravel())
        out = np.empty((T, 18, G, 4), np.int32)
        out[:, :16] = wd.reshape(T, G, 16, 4).transpose(0, 2, 1, 3).view(np.int32)   # piece 4r + q = row r, columns 4q..4q+3
        c16 = cols.astype(np.uint16).reshape(T, G, 2, 8)
        out[:, 16:] = np.ascontiguousarray(c16.transpose(0, 2, 1, 3)).view(np.int32).reshape(T, 2, G, 4)
        return np.ascontiguousarray(out)

    def fits(self, F, extra=2, esize=4):
        """Two stages of (x tile + halo, `extra` more tiles, records) fit in the CTA's shared memory, and the tile
        shape is one the kernel's warp layout covers.  `esize`: bytes per element (2 in the BF16 mode)."""
        rb = esize * F
        stage = (self.TR + self.HC) * rb + self.TR * rb * extra + self.width * self.TR * 8
        return (2 * stage + 128 <= PIPE_SMEM and self.TR % (8 * (128 // F)) == 0 and self.HC * (rb // 16) <= 512)

    def args(self, dense=False, reverse=False):
        """Leading arguments of dsb200_spmm_pipe; the last one carries the flags: bit 0 = dense-group records, bit 1 = walk the
        items from the end of memory to its start."""
        rev = 2 if reverse else 0
        if dense and self.dense is not None:
            return (self.dense, self.ext, self.nh, self.T, self.TR, self.HC, self.width, 1 | rev)
        return (self.rec, self.ext, self.nh, self.T, self.TR, self.HC, self.width, rev)


STENCIL_T = 16             # a fused-recurrence tile = 16 x 16 pixels = 256 consecutive NESTED indices (csrc/cheb_tile.cu)
STENCIL_W = 12             # floats per stencil record: 9 weights (dy, dx row-major) + 3 of padding (three 16-byte pieces)


class StencilPlan(object):
    """Tile plan of the fused Chebyshev recurrence (csrc/cheb_tile.cu): all K - 1 sparse products of a layer in ONE launch.

    HEALPix NESTED order is Morton order inside a base face, so T*T consecutive indices are a T x T square of pixels and
    the 8-neighbour Laplacian is a 3 x 3 stencil with per-pixel weights on it.  For a recurrence of S = K - 1 steps a tile
    is staged in shared memory with an S-ring halo as a flat (T + 2S)^2 image; step s is then computed on the image
    shrunk by s + 1 rings, every neighbour read from shared memory, nothing but the results touching HBM.
      p2v int32[nt, P, P]      vertex shown at every image position (P = T + 2S; row = iy, column = ix), -1 = none
      wts float32[nt, P-2, P-2, 12]  the 3 x 3 weights of L (or L^T) at every position that is ever computed:
                                entry 3 (dy + 1) + (dx + 1) multiplies the value at (row + dy, column + dx)
    The flat image crosses face boundaries with healpix's own face tables; around the 8 vertices of the base tessellation
    where only three faces meet no flat image exists.  The plan is therefore VERIFIED against the matrix itself: every
    stored entry of every row that is computed must sit at one of the 9 image positions around its row.  `valid` is False
    otherwise (or when the vertex order is not whole aligned tiles), and the per-step kernels stay in charge."""

    def __init__(self, col, val, nside, ids, S, device, T=STENCIL_T):
        from . import healpix as hp
        self.valid = False
        M, W = col.shape
        T2 = T * T
        if M == 0 or M % T2 or S < 1 or S > 7 or nside < T:
            return
        ids = np.arange(M, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        if len(ids) != M:
            return
        nt = M // T2
        blocks = ids.reshape(nt, T2)
        base = blocks[:, 0]
        if np.any(base % T2) or np.any(blocks != base[:, None] + np.arange(T2)):
            return
        npface = int(nside) * int(nside)
        if int(ids.max()) >= 12 * npface:
            return
        ix0, iy0, face = hp.nest2xyf(nside, base)
        P = T + 2 * S
        off = np.arange(P, dtype=np.int64) - S
        x = (ix0[:, None, None] + off[None, None, :]) + np.zeros((1, P, 1), np.int64)          # [nt, row, column]
        y = (iy0[:, None, None] + off[None, :, None]) + np.zeros((1, 1, P), np.int64)
        f = np.broadcast_to(face[:, None, None], x.shape)
        code = np.full(x.shape, 4, dtype=np.int64)
        lo, hi = x < 0, x >= nside
        x = np.where(lo, x + nside, np.where(hi, x - nside, x))
        code += hi.astype(np.int64) - lo.astype(np.int64)
        lo, hi = y < 0, y >= nside
        y = np.where(lo, y + nside, np.where(hi, y - nside, y))
        code += 3 * (hi.astype(np.int64) - lo.astype(np.int64))
        nf = hp._FACE_NB[code, f]
        bits = hp._SWAP[code, f >> 2]
        x = np.where(bits & 1, nside - x - 1, x)
        y = np.where(bits & 2, nside - y - 1, y)
        swap = (bits & 4) != 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        nested = np.where(nf >= 0, hp.xyf2nest(nside, x, y, np.maximum(nf, 0)), -1)
        # nested pixel number -> vertex (position in `ids`)
        order = np.argsort(ids, kind='stable')
        sid = ids[order]
        pos = np.clip(np.searchsorted(sid, nested), 0, M - 1)
        p2v = np.where((nested >= 0) & (sid[pos] == nested), order[pos], -1).astype(np.int64)
        # 3 x 3 weights on the computed region [1, P-1)^2, verified against the matrix
        Q = P - 2
        ctr = p2v[:, 1:P - 1, 1:P - 1]
        have = ctr >= 0
        vc = np.where(have, ctr, 0)
        cols = col[vc].astype(np.int64)                                     # [nt, Q, Q, W]
        vals = np.where(have[..., None], val[vc], np.float32(0))
        wts = np.zeros((nt, Q, Q, STENCIL_W), np.float32)
        matched = np.zeros(cols.shape, bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nb = p2v[:, 1 + dy:P - 1 + dy, 1 + dx:P - 1 + dx]
                hit = (cols == nb[..., None]) & (nb[..., None] >= 0) & ~matched
                wts[..., 3 * (dy + 1) + (dx + 1)] = np.where(hit, vals, np.float32(0)).sum(axis=-1, dtype=np.float32)
                matched |= hit
        if np.any((vals != 0) & ~matched):
            return                                                          # some neighbour is not on the flat image
        self.p2v_np = np.ascontiguousarray(p2v.astype(np.int32))
        self.wts_np = np.ascontiguousarray(wts)
        self.nt, self.T, self.S, self.P, self.M = nt, T, S, P, M
        if device is not None:
            self.p2v = torch.from_numpy(self.p2v_np).to(device)
            self.wts = torch.from_numpy(self.wts_np).to(device)
        self.valid = True


class DeviceGraph(object):
    """One Laplacian level on the GPU.  `.fwd` / `.bwd` = (rowptr | None, col, val, width)."""

    def __init__(self, L, device, force_csr=False, LT=None, nested=None):
        """`LT`: the matrix to use for the transposed products instead of L.T (vertex-sharded mode: the rows of
        L^T this rank owns are not the transpose of the rows of L it owns)."""
        L = sparse.csr_matrix(L, dtype=np.float32)
        L.sum_duplicates()
        L.sort_indices()                                   # row-major, ascending columns (tf.sparse_reorder)
        if L.shape[0] != L.shape[1]:
            raise ValueError('Laplacian must be square')
        self.M = L.shape[0]
        # `nested` = (nside, nested pixel number of every vertex | None): HEALPix geometry for the fused tile recurrence
        self.nested = nested
        self.nnz = int(L.nnz)
        self.device = device
        counts = np.diff(L.indptr)
        self.max_degree = int(counts.max()) if self.M else 0
        if LT is not None:
            LT = sparse.csr_matrix(LT, dtype=np.float32)
            LT.sum_duplicates()
            LT.sort_indices()
            self.max_degree = max(self.max_degree, int(np.diff(LT.indptr).max()))
        self.is_ell = (not force_csr) and self.max_degree <= ELL_MAX_WIDTH
        self.fwd_rec = self.bwd_rec = None
        self.fwd = self._pack(L, 'fwd')
        if LT is None:
            LT = sparse.csr_matrix(L.T)
            LT.sort_indices()
        self.symmetric_pattern = (LT.nnz == L.nnz and np.array_equal(LT.indptr, L.indptr)
                                  and np.array_equal(LT.indices, L.indices))
        if self.symmetric_pattern and np.array_equal(LT.data, L.data):
            self.bwd = self.fwd
            self.bwd_rec = self.fwd_rec
        else:
            self.bwd = self._pack(LT, 'bwd')

    @classmethod
    def from_ell(cls, col, val, nnz, device, nested=None):
        """A graph whose ELL arrays (int32 [M, w] / fp32 [M, w], ascending columns, padding = (row, 0)) already live on the
        device (graph_build.py).  The transposed values come from `dsb200_ell_transpose` (symmetric pattern required)."""
        from ._lib import call
        self = cls.__new__(cls)
        M, width = col.shape
        self.M, self.nnz, self.device, self.nested = int(M), int(nnz), device, nested
        self.max_degree, self.is_ell = int(width), True
        self.fwd = (None, col, val, int(width))
        valT = torch.empty_like(val)
        missing = torch.zeros(1, dtype=torch.int32, device=col.device)
        call('ell_transpose', col, val, M, width, valT, missing)
        if int(missing.item()):
            raise ValueError('from_ell: the sparsity pattern is not symmetric')
        self.symmetric_pattern = True
        rec = torch.empty((width, M, 2), dtype=torch.int32, device=col.device)
        call('ell_pack', col, val, width, M, rec)
        self.fwd_rec = rec
        if torch.equal(valT, val):
            self.bwd, self.bwd_rec = self.fwd, self.fwd_rec
        else:
            self.bwd = (None, col, valT, int(width))
            recT = torch.empty_like(rec)
            call('ell_pack', col, valT, width, M, recT)
            self.bwd_rec = recT
        return self

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
            # packed slot-major records rec[j, m] = (col, val bits): one contiguous piece per slot for the consecutive
            # rows of a warp (csrc/sparse.cu: spmm_rec_kernel); same bits as dsb200_ell_pack produces on the device
            rec = np.empty((width, self.M, 2), dtype=np.int32)
            rec[:, :, 0] = col.T
            rec[:, :, 1] = val.view(np.int32).T
            setattr(self, which + '_rec', torch.from_numpy(rec).to(self.device))
            return (None, torch.from_numpy(col).to(self.device), torch.from_numpy(val).to(self.device), width)
        return (torch.from_numpy(A.indptr.astype(np.int32)).to(self.device),
                torch.from_numpy(A.indices.astype(np.int32)).to(self.device),
                torch.from_numpy(A.data.astype(np.float32)).to(self.device), 0)

    def pipe_plan(self, which, F, esize=4):
        """PipePlan of L ('fwd') or L^T ('bwd') for F features of `esize` bytes per row (cached per tile size), or None."""
        if not self.is_ell or F not in (16, 32, 64):
            return None
        TR = int(os.environ.get('DSB200_PIPE_TILE_BYTES', PIPE_TILE_BYTES)) // (esize * F)
        if which == 'bwd' and self.bwd is self.fwd:
            which = 'fwd'
        key = (which, TR)
        cache = self.__dict__.setdefault('_pipe_plans', {})
        if key not in cache:
            _, col, val, width = getattr(self, which)
            plan = PipePlan(col.cpu().numpy(), val.cpu().numpy(), self.device, TR)
            cache[key] = plan if plan.valid else None
        plan = cache[key]
        return plan if plan is not None and plan.fits(F, esize=esize) else None

    def stencil_plan(self, which, S):
        """StencilPlan of L ('fwd') or L^T ('bwd') for a recurrence of S steps (cached), or None."""
        if not self.is_ell or self.nested is None:
            return None
        if which == 'bwd' and self.bwd is self.fwd:
            which = 'fwd'
        cache = self.__dict__.setdefault('_stencil_plans', {})
        key = (which, S)
        if key not in cache:
            _, col, val, width = getattr(self, which)
            nside, ids = self.nested
            plan = StencilPlan(col.cpu().numpy(), val.cpu().numpy(), nside, ids, S, self.device)
            cache[key] = plan if plan.valid else None
        return cache[key]

    def bytes_stored(self):
        """Bytes one sparse product reads for the matrix (8 per stored entry)."""
        return 8 * self.nnz + (0 if self.is_ell else 4 * (self.M + 1))


class ShardedGraph(object):
    """One Laplacian level of the vertex-sharded mode (new, the reference is single-device).

    Rank r of W owns the contiguous vertex range [r*M/W, (r+1)*M/W) -- in HEALPix nested order a set of whole
    base-pixel quadrants, so pooling by 4 stays rank-local at every level.  The local matrix has Ml = M/W owned rows
    and Me = Ml + H columns: the owned vertices first, then the H ghost vertices (1-ring halo) sorted by global id,
    i.e. grouped by owner.  It is stored square (Me x Me, empty ghost rows) so that the kernels of the single-GPU
    path run unchanged on [1, Me, F] tensors whose last H rows are ghosts.  `send[q]` lists the owned rows rank q
    needs (ascending), `recv[q]` = (first ghost slot, count) of the ghosts owned by q.

    `rings` > 1: the halo holds every vertex within `rings` edges of the owned range, and the local matrix also has the rows of
    the ghosts of rings 1 .. rings-1 (all their columns are local).  ONE exchange then feeds `rings` consecutive sparse products
    computed on all local rows: after product j the values are valid on the owned vertices and on rings 1 .. rings-j (the outer
    rings hold finite garbage that no valid row reads) -- a whole Chebyshev recurrence of order K = rings + 1 per exchange,
    instead of one exchange per product."""

    def __init__(self, L, rank, world, device, rings=1):
        L = sparse.csr_matrix(L, dtype=np.float32)
        L.sum_duplicates()
        M = L.shape[0]
        if M % world:
            raise ValueError('{} vertices cannot be split over {} ranks'.format(M, world))
        Ml = M // world
        LT = sparse.csr_matrix(L.T)
        self.M_global, self.Ml, self.rank, self.world = M, Ml, rank, world
        self.rings = int(rings)

        def ghosts_of(q, want_ring=False):
            """Vertices within `rings` edges (of L or L^T) of rank q's range, ascending global id (+ their ring number)."""
            lo, hi = q * Ml, (q + 1) * Ml
            c = np.union1d(L[lo:hi].indices, LT[lo:hi].indices)
            found = c[(c < lo) | (c >= hi)]
            ring = np.ones(len(found), np.int32)
            frontier = found
            for j in range(2, self.rings + 1):
                if not len(frontier):
                    break
                c = np.union1d(L[frontier].indices, LT[frontier].indices)
                c = c[(c < lo) | (c >= hi)]
                new = np.setdiff1d(c, found, assume_unique=True)
                found = np.concatenate([found, new])
                ring = np.concatenate([ring, np.full(len(new), j, np.int32)])
                frontier = new
            order = np.argsort(found, kind='stable')
            return (found[order], ring[order]) if want_ring else found[order]

        lo, hi = rank * Ml, (rank + 1) * Ml
        ghosts, ghost_ring = ghosts_of(rank, True)
        H = len(ghosts)
        self.H, self.Me = H, Ml + H
        self.ghost_ids, self.ghost_ring = ghosts, ghost_ring
        inner = np.nonzero(ghost_ring < self.rings)[0]           # ghosts that get a matrix row (none for rings == 1)

        def localise(A):
            rowsel = np.concatenate([np.arange(lo, hi, dtype=np.int64), ghosts[inner].astype(np.int64)])
            lrow = np.concatenate([np.arange(Ml, dtype=np.int64), Ml + inner.astype(np.int64)])
            rows = A[rowsel].tocoo()
            c = rows.col.astype(np.int64)
            own = (c >= lo) & (c < hi)
            pos = np.minimum(np.searchsorted(ghosts, c), max(H - 1, 0))
            local = own | ((ghosts[pos] == c) if H else own)
            assert local.all(), 'a row inside the halo references a vertex outside of it'
            lc = np.where(own, c - lo, Ml + pos)
            return sparse.csr_matrix((rows.data, (lrow[rows.row], lc)), shape=(self.Me, self.Me), dtype=np.float32)

        self.L_local, self.LT_local = localise(L), localise(LT)
        owner = ghosts // Ml
        self.recv = {}
        for q in range(world):
            n = int((owner == q).sum())
            if n:
                self.recv[q] = (int(np.searchsorted(owner, q)), n)
        self.send, self.send_off = {}, {}
        for q in range(world):
            if q == rank:
                continue
            g = ghosts_of(q)
            mine = g[(g >= lo) & (g < hi)] - lo
            if len(mine):
                self.send[q] = mine.astype(np.int32)
                self.send_off[q] = int(np.searchsorted(g, lo))        # where my rows start in rank q's ghost block
        self.device = device
        if device is not None:
            self.graph = DeviceGraph(self.L_local, device, LT=self.LT_local)
            self.M = self.Me
            self.send_dev = {q: torch.from_numpy(v).to(device) for q, v in self.send.items()}
            # one-kernel exchange over peer memory (csrc/peer_halo.cu): neighbour table and the concatenated row lists
            meta, first = [len(self.send), len(self.recv)], 0
            for q in sorted(self.send):
                meta += [q, len(self.send[q]), self.send_off[q], first]
                first += len(self.send[q])
            meta += sorted(self.recv)
            self.halo_meta = torch.tensor(meta, dtype=torch.int32, device=device)
            idx = np.concatenate([self.send[q] for q in sorted(self.send)]) if self.send else np.zeros(1, np.int32)
            self.halo_idx = torch.from_numpy(idx.astype(np.int32)).to(device)
            self.halo_max_rows = max([len(v) for v in self.send.values()] + [0])
