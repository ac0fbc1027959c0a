"""Host logic of the shared-memory tiled recurrence: the tile plan (row tiles + two halo rings + 64-byte row
records, deepsphere_tf1_b200/graph.py: TilePlan) is executed here by a numpy emulation of the kernel's three
phases (csrc/sparse_tiled.cu) and compared with scipy `L @ x` -- no GPU needed."""
import numpy as np
import pytest
import torch
from scipy import sparse

from helpers import healpix_L, random_L
from deepsphere_tf1_b200.graph import DeviceGraph, TilePlan, TILE_SLOTS


def _emulate(plan, M, x, y=None, z=None, u=None, coef1=(1.0, 0.0, 0.0), coef2=None):
    """What cheb_tiled_kernel computes, tile by tile, from the plan's own arrays."""
    info, recs, ext = plan.info.numpy(), plan.recs.numpy().view(np.uint32), plan.ext.numpy()
    out1 = np.zeros_like(x)
    out2 = np.zeros_like(x) if coef2 is not None else None
    for t in range(plan.ntiles):
        woff, n1, n2, eoff = info[t]
        r0 = t * plan.TR
        nt = min(plan.TR, M - r0)
        words = recs[woff:woff + 16 * n1].reshape(16, n1)
        vals = words[:TILE_SLOTS].T.copy().view(np.float32)
        pk = words[TILE_SLOTS:TILE_SLOTS + 5].T
        lcol = np.empty((n1, TILE_SLOTS), np.int64)
        lcol[:, 0::2] = pk & 0xffff
        lcol[:, 1::2] = pk >> 16
        rows = np.concatenate([np.arange(r0, r0 + nt), ext[eoff:eoff + n2 - nt]])     # global row of every local index
        A = x[rows]                                                    # phase 0
        nrows1 = n1 if coef2 is not None else nt
        v, c = vals[:nrows1], lcol[:nrows1]
        assert c.max() < (n2 if coef2 is not None else n1)
        T1 = coef1[0] * np.einsum('rj,rjf->rf', v, A[c])               # phase 1
        if y is not None:
            T1 = T1 + coef1[1] * y[rows[:nrows1]]
        if z is not None:
            T1 = T1 + coef1[2] * z[rows[:nrows1]]
        out1[r0:r0 + nt] = T1[:nt]
        if coef2 is not None:                                          # phase 2
            v, c = vals[:nt], lcol[:nt]
            assert c.max() < n1
            T2 = coef2[0] * np.einsum('rj,rjf->rf', v, T1[c]) + coef2[1] * A[:nt]
            if u is not None:
                T2 = T2 + coef2[2] * u[r0:r0 + nt]
            out2[r0:r0 + nt] = T2
    return out1, out2


@pytest.mark.parametrize('case', ['full_sphere', 'partial_sphere', 'ragged_last_tile'])
def test_tile_plan_reproduces_two_recurrence_steps(case):
    L = {'full_sphere': lambda: healpix_L(16), 'partial_sphere': lambda: healpix_L(32, order=2),
         'ragged_last_tile': lambda: healpix_L(8).tocsr()[:700, :700]}[case]()
    L = sparse.csr_matrix(L, dtype=np.float32)
    M = L.shape[0]
    g = DeviceGraph(L, torch.device('cpu'), tiled=True)
    plan = g.fwd_tiles
    assert plan is not None and plan.valid and plan.ntiles == (M + plan.TR - 1) // plan.TR
    rs = np.random.RandomState(0)
    x, y, u = [rs.randn(M, 4).astype(np.float64) for _ in range(3)]
    Ld = L.astype(np.float64)
    # single step
    o1, _ = _emulate(plan, M, x, y, None, None, (2.0, -1.0, 0.0))
    np.testing.assert_allclose(o1, 2 * (Ld @ x) - y, rtol=0, atol=1e-5)
    # two fused steps: X_k = 2 L x - y ; X_k+1 = 2 L X_k - x (+ u, as in the adjoint pairs)
    o1, o2 = _emulate(plan, M, x, y, None, u, (2.0, -1.0, 0.0), (2.0, -1.0, 0.5))
    t1 = 2 * (Ld @ x) - y
    np.testing.assert_allclose(o1, t1, rtol=0, atol=1e-5)
    np.testing.assert_allclose(o2, 2 * (Ld @ t1) - x + 0.5 * u, rtol=0, atol=1e-5)


def test_tile_plan_halo_rings_are_minimal_and_keep_bank_bits():
    L = sparse.csr_matrix(healpix_L(16), dtype=np.float32)
    g = DeviceGraph(L, torch.device('cpu'), tiled=True)
    plan = g.fwd_tiles
    info, ext = plan.info.numpy(), plan.ext.numpy()
    pattern = sparse.csr_matrix((np.ones(L.nnz), L.indices, L.indptr), shape=L.shape)
    recs = plan.recs.numpy().view(np.uint32)
    for t in range(plan.ntiles):
        woff, n1, n2, eoff = info[t]
        r0 = t * plan.TR
        nt = min(plan.TR, L.shape[0] - r0)
        tile = np.arange(r0, r0 + nt)
        e = ext[eoff:eoff + n2 - nt]
        loc = np.arange(nt, n2)
        vals = recs[woff:woff + 16 * n1].reshape(16, n1)[:TILE_SLOTS].T.copy().view(np.float32)
        hole = np.ones(n2 - nt, bool)
        hole[:n1 - nt] = ~np.any(vals[nt:n1] != 0, axis=1)              # ring-1 holes have all-zero records
        real1 = e[:n1 - nt][~hole[:n1 - nt]]
        reach1 = np.unique(pattern[tile].indices)
        assert set(real1) == set(reach1) - set(tile)
        reach2 = set(np.unique(pattern[np.concatenate([tile, real1])].indices)) - set(tile) - set(real1)
        ring2 = e[n1 - nt:]
        assert reach2 <= set(ring2) and set(ring2) <= reach2 | {r0}
        # a real halo row keeps the low 3 bits of its global id in its local index
        keep = np.concatenate([~hole[:n1 - nt], ring2 != r0])
        assert np.all((e[keep] & 7) == (loc[keep] & 7))
        assert n1 % 8 == 0 and n2 % 8 == 0


def test_graphs_without_locality_or_wide_rows_keep_the_gather_kernels():
    # random sparse matrix: the halo of a 256-row tile is (almost) the whole graph
    g = DeviceGraph(random_L(6000, 6), torch.device('cpu'), tiled=True)
    assert g.fwd_tiles is None and g.bwd_tiles is None
    # more than 10 entries per row do not fit a row record
    n = 600
    A = sparse.random(n, n, density=14.0 / n, random_state=1, format='csr', dtype=np.float32)
    plan = TilePlan(np.zeros((n, 12), np.int32), np.zeros((n, 12), np.float32), torch.device('cpu'))
    assert not plan.valid and A.shape == (n, n)


def test_nonsymmetric_values_get_their_own_transposed_plan():
    L = sparse.csr_matrix(healpix_L(8), dtype=np.float32)
    L = sparse.csr_matrix(L.multiply(sparse.csr_matrix(1.0 + 0.1 * np.random.RandomState(0).rand(*L.shape))), dtype=np.float32)
    g = DeviceGraph(L, torch.device('cpu'), tiled=True)
    assert g.bwd_tiles is not None and g.bwd_tiles is not g.fwd_tiles
    rs = np.random.RandomState(1)
    x = rs.randn(L.shape[0], 4)
    o1, _ = _emulate(g.bwd_tiles, L.shape[0], x)
    np.testing.assert_allclose(o1, L.T.astype(np.float64) @ x, rtol=0, atol=1e-5)


@pytest.mark.parametrize('case', ['full_sphere', 'partial_sphere', 'ragged_last_block', 'nonsym_values'])
def test_block_plan_is_the_matrix(case):
    """Tensor-core block formulation (csrc/sparse_blk.cu): the column lists and the A fragments of the 16-row blocks,
    unpacked the way mma.m16n8k8 reads them, reproduce L x (and L^T x from the transposed plan)."""
    from deepsphere_tf1_b200.graph import BLOCK_ROWS
    L = {'full_sphere': lambda: healpix_L(16), 'partial_sphere': lambda: healpix_L(32, order=2),
         'ragged_last_block': lambda: healpix_L(8).tocsr()[:700, :700], 'nonsym_values': lambda: healpix_L(8)}[case]()
    L = sparse.csr_matrix(L, dtype=np.float32)
    if case == 'nonsym_values':
        L = sparse.csr_matrix(L.multiply(sparse.csr_matrix(1.0 + 0.1 * np.random.RandomState(0).rand(*L.shape))), dtype=np.float32)
    M = L.shape[0]
    g = DeviceGraph(L, torch.device('cpu'), blocks=True)
    x = np.random.RandomState(1).randn(M, 3)
    for bp, A_ref in ((g.fwd_blocks, L), (g.bwd_blocks, L.T)):
        assert bp is not None and bp.nblocks == (M + BLOCK_ROWS - 1) // BLOCK_ROWS and 5 <= bp.KS <= 7
        cols, fr = bp.cols.numpy(), bp.frags.numpy()
        assert cols.shape == (bp.nblocks, 8 * bp.KS) and fr.shape == (bp.nblocks, bp.KS, 32, 4)
        assert cols.min() >= 0 and cols.max() < M
        A = np.zeros((bp.nblocks, BLOCK_ROWS, 8 * bp.KS), np.float32)
        lane = np.arange(32)
        gq, t = lane // 4, lane % 4
        for s_ in range(bp.KS):
            A[:, gq, 8 * s_ + t] = fr[:, s_, :, 0]
            A[:, gq + 8, 8 * s_ + t] = fr[:, s_, :, 1]
            A[:, gq, 8 * s_ + t + 4] = fr[:, s_, :, 2]
            A[:, gq + 8, 8 * s_ + t + 4] = fr[:, s_, :, 3]
        out = np.einsum('bik,bkf->bif', A.astype(np.float64), x[cols]).reshape(-1, 3)[:M]
        np.testing.assert_allclose(out, A_ref.astype(np.float64) @ x, rtol=0, atol=1e-12)


def test_staged_block_plan_localises_the_block_columns():
    """csrc/sparse_stile.cu: per 128-row tile the ring-1 halo list and the block columns as LOCAL indices."""
    from deepsphere_tf1_b200.graph import STILE_ROWS
    L = sparse.csr_matrix(healpix_L(16), dtype=np.float32)
    M = L.shape[0]
    g = DeviceGraph(L, torch.device('cpu'), blocks=True, staged=True)
    sp, bp = g.fwd_staged, g.fwd_blocks
    assert sp is not None and sp.ntiles == (M + STILE_ROWS - 1) // STILE_ROWS and sp.nblocks == bp.nblocks and sp.KS == bp.KS
    tinfo, ext, lcols = sp.tinfo.numpy(), sp.ext.numpy(), sp.lcols.numpy().view(np.uint16).astype(np.int64)
    cols = bp.cols.numpy()
    for t in range(sp.ntiles):
        eoff, n1 = tinfo[t, 0], tinfo[t, 1]
        r0 = t * STILE_ROWS
        nt = min(STILE_ROWS, M - r0)
        rows = np.concatenate([np.arange(r0, r0 + nt), ext[eoff:eoff + n1 - nt]])       # global row of every local index
        assert n1 <= sp.n1cap and np.all(np.diff(ext[eoff:eoff + n1 - nt]) > 0)
        blk = slice(t * 8, min((t + 1) * 8, sp.nblocks))
        np.testing.assert_array_equal(rows[lcols[blk]], cols[blk])


@pytest.mark.parametrize('TR', [32, 64, 128])
@pytest.mark.parametrize('graph', ['healpix_full', 'healpix_patch', 'ragged'])
def test_pipe_plan_local_gather_equals_the_global_product(graph, TR):
    """PipePlan (csrc/sparse_pipe.cu): emulate the kernel's staging on the CPU -- tile rows followed by the halo rows of
    `ext` -- and check that the local-column records reproduce L x exactly (same products, same order)."""
    from scipy import sparse
    from deepsphere_tf1_b200.graph import DeviceGraph, PipePlan
    from helpers import healpix_L
    L = {'healpix_full': lambda: healpix_L(8), 'healpix_patch': lambda: healpix_L(32, order=2),
         'ragged': lambda: sparse.coo_matrix(sparse.csr_matrix(healpix_L(8))[:700, :700])}[graph]()
    g = DeviceGraph(L, 'cpu')
    _, col, val, width = g.fwd
    col, val = col.numpy(), val.numpy()
    plan = PipePlan(col, val, None, TR)
    assert plan.valid and plan.width == width and plan.T == (g.M + TR - 1) // TR
    rs = np.random.RandomState(TR)
    x = rs.randn(g.M, 4).astype(np.float32)
    ref = np.zeros_like(x)
    for j in range(width):
        ref = ref + val[:, j:j + 1] * x[col[:, j]]
    out = np.zeros_like(x)
    for t in range(plan.T):
        r0, nt = t * TR, min(TR, g.M - t * TR)
        nh = int(plan.nh_np[t])
        assert nh <= plan.HC
        stage = np.zeros((TR + plan.HC, 4), np.float32)
        stage[:nt] = x[r0:r0 + nt]
        stage[TR:TR + nh] = x[plan.ext_np[t, :nh]]
        assert np.all((plan.ext_np[t, :nh] < r0) | (plan.ext_np[t, :nh] >= r0 + nt))
        acc = np.zeros((nt, 4), np.float32)
        for j in range(width):
            lc = plan.rec_np[t, j, :nt, 0]
            assert lc.max() < TR + max(nh, 1) and lc.min() >= 0
            acc = acc + plan.rec_np[t, j, :nt, 1].view(np.float32)[:, None] * stage[lc]
        out[r0:r0 + nt] = acc
        # dense-group records: 4 rows x <= 16 distinct columns, [18][G] 16-byte pieces per tile
        d = plan.dense_np[t]
        G = TR // 4
        wd = d[:16].view(np.float32).reshape(4, 4, G, 4).transpose(2, 0, 1, 3).reshape(G, 4, 16)
        cols = d[16:].view(np.uint16).reshape(2, G, 8).transpose(1, 0, 2).reshape(G, 16)
        for gidx in range((nt + 3) // 4):
            rows = slice(4 * gidx, min(4 * gidx + 4, nt))
            dn = wd[gidx].astype(np.float64) @ stage[cols[gidx]].astype(np.float64)
            dn = dn[np.arange(4) ^ (gidx & 1)]                       # stored row r = row r ^ (g & 1)
            assert np.allclose(dn[:rows.stop - rows.start], acc[rows], rtol=0, atol=1e-5)
    assert np.array_equal(out, ref)
