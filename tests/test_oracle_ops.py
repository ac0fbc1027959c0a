"""CPU checks of the oracle itself, next to its pinning against the reference's own graph code (tests/test_ops_golden.py):
mathematical identities and the known answers the reference prints in experiments/cosmo/demo.ipynb."""
import numpy as np
import pytest
import torch

from helpers import healpix_L
from oracle import ops as oops
from oracle.model import OracleCgcnn


def test_parameter_counts_of_the_cosmo_demo():
    """demo.ipynb:861-889: weights 80 + 2,560 + 10,240 + 20,480 + 20,480 + 640, biases 16+32+64+64+64."""
    L = [healpix_L(1)] * 6
    m = OracleCgcnn(L, F=[16, 32, 64, 64, 64, 2], K=[5] * 6, p=[1] * 6, batch_norm=[True] * 6, M=[], statistics='mean')
    sizes = [int(np.prod(v.shape)) for k, v in m.params.items() if k.endswith('weights')]
    assert sizes == [80, 2560, 10240, 20480, 20480, 640]
    assert m.n_params() == 54720


def test_chebyshev_recurrence_is_cos_k_arccos_on_eigenvectors():
    """T_k(L) v = cos(k arccos(lambda)) v for an eigenpair of the rescaled Laplacian."""
    L = healpix_L(2)
    lam, V = np.linalg.eigh(L.toarray().astype(np.float64))
    Lt = oops.sparse_to_torch(L, torch.float64)
    K = 6
    for idx in (0, 7, len(lam) - 1):
        v = torch.from_numpy(V[:, idx]).reshape(1, -1, 1)
        # W picks order k for output feature k: y[..., k] = T_k(L) v
        W = torch.eye(K, dtype=torch.float64)
        y = oops.chebyshev5(v, Lt, W, K)[0]
        for k in range(K):
            expected = np.cos(k * np.arccos(np.clip(lam[idx], -1, 1))) * V[:, idx]
            np.testing.assert_allclose(y[:, k].numpy(), expected, atol=1e-5)   # L is stored in float32


def test_chebyshev5_k1_is_a_per_vertex_linear_map_and_is_linear():
    L = oops.sparse_to_torch(healpix_L(2))
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.randn(2, 48, 3).astype(np.float32))
    W = torch.from_numpy(rs.randn(3, 4).astype(np.float32))
    torch.testing.assert_close(oops.chebyshev5(x, L, W, 1), x @ W)
    W5 = torch.from_numpy(rs.randn(15, 4).astype(np.float32))
    a = oops.chebyshev5(2 * x, L, W5, 5)
    torch.testing.assert_close(a, 2 * oops.chebyshev5(x, L, W5, 5), rtol=1e-5, atol=1e-5)


def test_autograd_gradient_matches_finite_differences_in_float64():
    L = [healpix_L(2), healpix_L(1)]
    m = OracleCgcnn(L, F=[4, 3], K=[3, 3], p=[4, 1], batch_norm=[True, False], M=[], statistics='mean', num_feat_in=2,
                    dtype=torch.float64, regularization=0.1)
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.randn(3, 48, 2))
    labels = torch.from_numpy(rs.randint(0, 3, size=3))
    total, _, grads, _ = m.loss_and_grads(x, labels)
    name = 'conv1/weights'
    base = m.params[name].clone()
    for (i, j) in [(0, 0), (3, 2), (5, 1)]:
        eps = 1e-6
        vals = []
        for sgn in (+1, -1):
            m.params[name] = base.clone()
            m.params[name][i, j] += sgn * eps
            t, _, _, _ = m.loss_and_grads(x, labels)
            vals.append(float(t))
        m.params[name] = base
        assert (vals[0] - vals[1]) / (2 * eps) == pytest.approx(float(grads[name][i, j]), rel=1e-5, abs=1e-9)


def test_initial_loss_is_ln2_for_two_classes():
    """demo.ipynb:987: the first logged cross-entropy of the 2-class cosmo net is 6.92e-1 = ln 2 (zero logits)."""
    logits = torch.zeros(16, 2)
    labels = torch.arange(16) % 2
    _, ce = oops.loss(logits, labels, False, [(torch.ones(2, 2), 1.0)], 0.0)
    assert float(ce) == pytest.approx(np.log(2), rel=1e-6)


def test_monomials_are_powers_of_L_on_eigenvectors():
    """cgcnn.monomials (models.py:1085-1120): the k-th basis plane of an eigenvector v is lambda^k v."""
    L = healpix_L(2)
    lam, V = np.linalg.eigh(L.toarray().astype(np.float64))
    Lt = oops.sparse_to_torch(L, torch.float64)
    K = 5
    for idx in (1, len(lam) - 1):
        v = torch.from_numpy(V[:, idx]).reshape(1, -1, 1)
        y = oops.monomials(v, Lt, torch.eye(K, dtype=torch.float64), K)          # y[..., k] = L^k v
        for k in range(K):
            assert torch.allclose(y[0, :, k], lam[idx] ** k * v[0, :, 0], atol=1e-6)   # (L is fp32, symmetric to 1 ulp)


def test_weighted_cross_entropy_matches_its_definition():
    """models.py:795-802: ce_r * (one_hot(label_r, 3) . weights), averaged over all rows."""
    rs = np.random.RandomState(0)
    logits = torch.from_numpy(rs.randn(7, 3))
    labels = torch.from_numpy(rs.randint(0, 3, size=7))
    w = np.array(oops.CLASS_WEIGHTS)
    lse = torch.logsumexp(logits, dim=1)
    ce = (lse - logits[torch.arange(7), labels]).numpy()
    expect = float(np.mean(ce * w[labels.numpy()]))
    got, _ = oops.loss(logits, labels, False, [(torch.zeros(1), 1.0)], 0.0, weighted=True)
    assert float(got) == pytest.approx(expect, rel=1e-12)
    with pytest.raises(ValueError):
        oops.loss(torch.zeros(4, 5), torch.zeros(4, dtype=torch.long), False, [(torch.zeros(1), 1.0)], 0.0, weighted=True)


def test_learned_histogram_initial_bins_partition_their_range():
    """models.py:1173-1190: triangular bins with 50 % overlap -- inside the initial range every value falls into exactly
    two neighbouring bins whose memberships sum to a constant, so a histogram row sums to that constant times the scale."""
    F, bins, rng = 3, 20, 2
    c, w = oops.histogram_init(F, bins, rng, torch.float64)
    assert c.shape == (1, 1, F, bins) and float(w[0, 0, 0, 0]) == pytest.approx(4 * rng / bins)
    x = (torch.rand(2, 500, F, dtype=torch.float64) * 2 - 1) * 1.7          # well inside [-2, 2]
    h = oops.learned_histogram(x, c, w, bins, rng).reshape(2, F, bins)
    # spacing of the centers d = 2 rng / (bins - 1); membership 1 - |x - c| * 0.4 summed over the (up to 5) active bins
    assert torch.all(h >= 0)
    single = oops.learned_histogram(c[:, :, :, 7:8].reshape(1, 1, F).clone(), c, w, bins, rng).reshape(F, bins)
    assert float(single[0, 7]) == pytest.approx(bins / rng / 4)            # a value on a center: membership 1 in that bin
