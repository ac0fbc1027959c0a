"""Host logic of the fused tile recurrence (csrc/cheb_tile.cu): the tile plan (deepsphere_tf1_b200/graph.py: StencilPlan --
flat (T + 2S)^2 images of the NESTED pixel squares and per-pixel 3 x 3 weights) is executed here by a numpy emulation of the
kernel's steps and compared with scipy `L @ x` -- no GPU needed."""
import numpy as np
import pytest
import torch
from scipy import sparse

from helpers import healpix_L
from deepsphere_tf1_b200 import utils
from deepsphere_tf1_b200.graph import DeviceGraph, StencilPlan


def _emulate_forward(plan, x, K):
    """What cheb_tile_kernel computes in its forward direction: planes 1..K-1 of the Chebyshev basis of x [M, F]."""
    M, F = x.shape
    S, P, T = plan.S, plan.P, plan.T
    assert K - 1 == S
    out = np.zeros((K - 1, M, F))
    for t in range(plan.nt):
        p2v = plan.p2v_np[t]
        w = plan.wts_np[t].astype(np.float64)
        img = np.where((p2v >= 0)[..., None], x[np.maximum(p2v, 0)], 0.0)           # staged tile + halo, absent = 0
        prev, cur = None, img
        for s in range(S):
            lo, hi = s + 1, P - 1 - s
            acc = np.zeros((hi - lo, hi - lo, F))
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    acc += w[lo - 1:hi - 1, lo - 1:hi - 1, 3 * (dy + 1) + dx + 1, None] * cur[lo + dy:hi + dy, lo + dx:hi + dx]
            new = np.zeros_like(img)
            new[lo:hi, lo:hi] = acc if s == 0 else 2 * acc - prev[lo:hi, lo:hi]
            prev, cur = cur, new
            inner = p2v[S:S + T, S:S + T]
            out[s][inner.ravel()] = cur[S:S + T, S:S + T].reshape(-1, F)
    return out


@pytest.mark.parametrize('nside,order,K', [(64, 2, 5), (64, 1, 4), (32, 1, 2), (128, 4, 5)])
def test_stencil_plan_reproduces_the_recurrence(nside, order, K):
    L = healpix_L(nside, order)
    ids = utils.nside2indexes([nside], order)[0]
    g = DeviceGraph(L, torch.device('cpu'), nested=(nside, ids))
    plan = g.stencil_plan('fwd', K - 1)
    assert plan is not None and plan.nt == L.shape[0] // 256
    # inner positions of every tile are its own 256 vertices in Morton order
    t = plan.nt - 1
    inner = plan.p2v_np[t, plan.S:plan.S + 16, plan.S:plan.S + 16]
    assert sorted(inner.ravel().tolist()) == list(range(256 * t, 256 * (t + 1)))
    assert inner[0, 1] == 256 * t + 1 and inner[1, 0] == 256 * t + 2                  # x = even bits, y = odd bits
    rs = np.random.RandomState(0)
    x = rs.randn(L.shape[0], 3)
    L64 = sparse.csr_matrix(L, dtype=np.float64)
    ref = [x, L64 @ x]
    for k in range(2, K):
        ref.append(2 * (L64 @ ref[-1]) - ref[-2])
    got = _emulate_forward(plan, x, K)
    np.testing.assert_allclose(got, np.stack(ref[1:]), rtol=0, atol=1e-12)


def test_transposed_plan_uses_the_transposed_values():
    L = healpix_L(64, 2)
    ids = utils.nside2indexes([64], 2)[0]
    g = DeviceGraph(L, torch.device('cpu'), nested=(64, ids))
    assert g.bwd is not g.fwd                                      # D^-1/2 W D^-1/2 in fp32 is symmetric to 1 ulp only
    pf, pb = g.stencil_plan('fwd', 2), g.stencil_plan('bwd', 2)
    assert pb is not None and pb is not pf
    x = np.random.RandomState(1).randn(L.shape[0], 2)
    LT = sparse.csr_matrix(L.T, dtype=np.float64)
    got = _emulate_forward(pb, x, 3)
    np.testing.assert_allclose(got[0], LT @ x, rtol=0, atol=1e-12)


def test_graphs_without_a_flat_image_have_no_plan():
    # full sphere: around the 8 three-face vertices of the base tessellation no flat image exists
    L = healpix_L(16, 0)
    g = DeviceGraph(L, torch.device('cpu'), nested=(16, None))
    assert g.stencil_plan('fwd', 4) is None
    # wrong geometry (a 32 x 32 patch declared as four faces of nside 16): the verification against the matrix refuses it
    L = healpix_L(64, 2)
    g = DeviceGraph(L, torch.device("cpu"), nested=(16, None))
    assert g.stencil_plan('fwd', 4) is None
    # no geometry, ragged sizes, CSR graphs
    assert DeviceGraph(L, torch.device('cpu')).stencil_plan('fwd', 4) is None
    col = np.zeros((300, 9), np.int32)
    assert not StencilPlan(col, np.zeros((300, 9), np.float32), 64, None, 4, None).valid
