"""Host-side tile plan of the pipelined sparse product (graph.PipePlan; csrc/sparse_pipe.cu): what the kernel will read from its
shared-memory stage must have been staged -- the tile's real rows and its halo slots, nothing else."""
import numpy as np
import pytest
from scipy import sparse

from deepsphere_tf1_b200.graph import PipePlan

from helpers import healpix_L


def _ell(L, width=9):
    L = sparse.csr_matrix(L, dtype=np.float32)
    L.sort_indices()
    M = L.shape[0]
    col = np.repeat(np.arange(M, dtype=np.int32)[:, None], width, axis=1)          # padding slots: (row, 0)
    val = np.zeros((M, width), np.float32)
    for r in range(M):
        a, b = L.indptr[r], L.indptr[r + 1]
        col[r, :b - a], val[r, :b - a] = L.indices[a:b], L.data[a:b]
    return col, val


@pytest.mark.parametrize('M', [3001, 3002, 3003, 3072, 2900])
@pytest.mark.parametrize('TR', [64, 128, 256])
def test_dense_groups_only_read_rows_the_kernel_stages(M, TR):
    """A row count that is not a multiple of 4 leaves a group of the last tile with padding rows next to real ones; the kernel
    stages only the nt = M - t*TR real rows of a tile (bulk copy of nt rows) and processes every group with 4 g < nt.  Every
    column slot of such a group must be a real row of the tile or one of its halo slots -- a slot that pointed at a padding row
    would multiply stale (in a CTA's first item: uninitialised) shared memory by 0.0 into the real rows."""
    L = sparse.csr_matrix(healpix_L(16))[:M, :M]
    col, val = _ell(L)
    plan = PipePlan(col, val, None, TR)
    assert plan.valid and plan.dense_np is not None
    T, G = plan.T, TR // 4
    cols = np.ascontiguousarray(plan.dense_np[:, 16:]).view(np.uint16).reshape(T, 2, G, 8)
    cols = cols.transpose(0, 2, 1, 3).reshape(T, G, 16).astype(np.int64)             # [tile, group, slot] local columns
    wts = np.ascontiguousarray(plan.dense_np[:, :16]).view(np.float32).reshape(T, 4, 4, G, 4)   # [tile, r, q, group, 4]
    for t in range(T):
        nt = min(TR, M - t * TR)
        live = np.arange(G) * 4 < nt
        c = cols[t, live]
        inside = c < TR
        assert (c[inside] < nt).all(), 'tile {}: a slot points at a padding row'.format(t)
        assert (c[~inside] - TR < max(int(plan.nh_np[t]), 1)).all(), 'tile {}: a slot points past the halo list'.format(t)
    # the plan still is the matrix: rebuild L x from the dense groups of every tile
    rs = np.random.RandomState(0)
    x = rs.randn(M).astype(np.float64)
    want = sparse.csr_matrix(L, dtype=np.float64) @ x
    got = np.zeros(M)
    for t in range(T):
        nt = min(TR, M - t * TR)
        stage = np.full(TR + plan.HC, np.nan)                                        # what was never staged is poison
        stage[:nt] = x[t * TR:t * TR + nt]
        stage[TR:TR + plan.HC] = x[plan.ext_np[t]]
        for g in range(G):
            if 4 * g >= nt:
                break
            v = stage[cols[t, g]]
            for r in range(4):
                row = 4 * g + (r ^ (g & 1))
                if row < nt:
                    w = wts[t, r, :, g, :].reshape(16).astype(np.float64)
                    got[t * TR + row] = float(np.dot(w, v))
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)


@pytest.mark.parametrize('M, width', [(3001, 9), (3072, 9), (2562, 7), (2900, 7)])
def test_record_form_only_reads_rows_the_kernel_stages(M, width):
    """The record form of the same plan (one lane group per row, `width` gathers): emulate the kernel's reads on a stage whose
    unstaged part is NaN and compare with L x."""
    if width == 9:
        L = sparse.csr_matrix(healpix_L(16))[:M, :M]
    else:                                   # a banded matrix of degree <= 7 (the icosahedral graphs have ELL width 7)
        rs = np.random.RandomState(1)
        rows = np.repeat(np.arange(M), 6)
        cols = np.clip(rows + rs.randint(-40, 41, size=rows.size), 0, M - 1)
        L = sparse.coo_matrix((rs.randn(rows.size).astype(np.float32), (rows, cols)), shape=(M, M)).tocsr()
        L.sum_duplicates()
    col, val = _ell(L, width)
    TR = 128
    plan = PipePlan(col, val, None, TR)
    assert plan.valid and (plan.dense_np is None) == (width == 7)
    rs = np.random.RandomState(0)
    x = rs.randn(M)
    want = sparse.csr_matrix(L, dtype=np.float64) @ x
    got = np.zeros(M)
    for t in range(plan.T):
        nt = min(TR, M - t * TR)
        stage = np.full(TR + plan.HC, np.nan)
        stage[:nt] = x[t * TR:t * TR + nt]
        stage[TR:TR + plan.HC] = x[plan.ext_np[t]]
        rec = plan.rec_np[t]                                                         # [W, TR, 2]
        lc, w = rec[:, :nt, 0], np.ascontiguousarray(rec[:, :nt, 1]).view(np.float32)
        got[t * TR:t * TR + nt] = (w.astype(np.float64) * stage[lc]).sum(axis=0)
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
