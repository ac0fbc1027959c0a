"""GPU parity of whole models (`cgcnn` / `deepsphere` with the reference's constructor) against
the CPU oracle `OracleCgcnn` on the same weights and inputs: logits, loss, every gradient, and a
few optimizer steps.  Small versions of the five BASELINE.json configurations.
Tolerance: fp32, max error <= 1e-4 of the largest expected magnitude (north_star); after several
Adam steps 5e-4 (the optimizer divides by sqrt(v), which amplifies the 1e-7 noise of tiny gradients)."""
import numpy as np
import pytest
import torch

from helpers import assert_close, copy_weights, healpix_L, random_knn_L

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _pair(L, sampling='healpix', seed=2, batch_size=4, finest=None, **kw):
    from deepsphere_tf1_b200 import models, optimizers
    from oracle.model import OracleCgcnn
    okw = {k: v for k, v in kw.items() if k in ('F', 'K', 'p', 'batch_norm', 'M', 'num_feat_in', 'pool', 'activation',
                                                'statistics', 'Fseg', 'regularization', 'dropout', 'regression', 'dense', 'conv', 'weighted')}
    attrs = {'sampling': sampling}
    if finest is not None:
        okw['finest_vertices'] = finest
        attrs['icosahedron_finest'] = finest
    oracle = OracleCgcnn(L, sampling=sampling, seed=seed, **okw)
    cls = type('cgcnn_' + sampling, (models.cgcnn,), attrs)
    kw.setdefault('optimizer', lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8))
    kw.setdefault('scheduler', lambda step: optimizers.exponential_decay(2e-3, step, 10, 0.9))
    model = cls(L=L, num_epochs=1, batch_size=batch_size, eval_frequency=10, dir_name='_test_parity', verbose=False, **kw)
    copy_weights(oracle, model)
    return oracle, model


def _check_step(oracle, model, x, labels, steps=3, rtol_after=5e-4, opt=None, wfloor=1e-5):
    from oracle import optim_tf1
    from deepsphere_tf1_b200 import ops
    xt = torch.from_numpy(x)
    lt = torch.from_numpy(labels)
    # ---- forward / loss / gradients on identical weights
    total, logits_ref, grads, _ = oracle.loss_and_grads(xt, lt, training=True)
    xd, ld = model._to_device(x), model._labels_to_device(labels)
    model._P.g.zero_()
    model._loss.zero_()
    logits, _ = model._forward(xd, training=True, save=True)
    assert_close(logits, logits_ref.reshape(logits.shape), RTOL, 'logits')
    d = model._loss_fwd_bwd(logits, ld, True)
    model._backward(d)
    if model._P.has_reg:
        ops.l2_reg(model._P.w, model._P.coef, model._P.g, model._loss)
    assert_close(model._loss, total.reshape(1), RTOL, 'loss')
    # A bias in front of a batch-normalised filter has a mathematically zero gradient (T_k(L) maps
    # constants to constants on a combinatorial Laplacian); such gradients are pure rounding noise in
    # any implementation, so every gradient is compared on a scale of at least 1e-3 of the largest one.
    gmax = max(float(g.abs().max()) for g in grads.values())
    for name, g in grads.items():
        assert_close(model._P.gviews[name].reshape(-1), g.reshape(-1), 2 * RTOL, 'grad ' + name, floor=1e-3 * gmax)
    # undo the batch-norm moving-average update of the probe forward
    for k, v in oracle.state.items():
        model.set_var(k, v.numpy())
    # ---- a few optimizer steps (sess.run([op_train, op_loss]))
    opt = opt or optim_tf1.Adam(epsilon=1e-8)
    for t in range(steps):
        lr = model.scheduler(t)
        loss_ref, _ = oracle.train_step(xt, lt, opt, lr)
        lr_p, loss_p = model.train_step(x, labels)
        assert lr_p == pytest.approx(lr)
        assert float(loss_p) == pytest.approx(loss_ref, rel=rtol_after)
    for name, v in oracle.params.items():
        # floor: a bias in front of a batch-normalised filter only ever receives rounding noise (see above); after a few
        # momentum steps it holds ~1e-9 in any implementation
        assert_close(model._P.views[name].reshape(-1), v.reshape(-1), rtol_after, 'weights ' + name, floor=wfloor)
    for name, v in oracle.state.items():
        # a moving mean is compared on the scale of the feature's standard deviation
        floor = float(oracle.state[name.replace('moving_mean', 'moving_variance')].max()) ** 0.5
        assert_close(model._state[name], v, rtol_after, name, floor=floor if name.endswith('moving_mean') else 0.0)
    # ---- inference mode (moving statistics)
    ref, desc_ref = oracle.predict_logits(xt)
    out, desc = model._forward(xd, training=False, save=False)
    assert_close(out, ref.reshape(out.shape), rtol_after, 'eval logits')
    assert_close(desc, desc_ref, rtol_after, 'descriptor')


def _momentum():
    """Multi-step parity for the decoder nets uses Momentum: Adam's first updates are lr * sign(g),
    so gradient entries that are rounding noise (see above) move weights by +-lr in ANY two
    implementations; the Adam rule itself is pinned in test_parity_ops and by the two classifier nets."""
    from deepsphere_tf1_b200 import optimizers
    from oracle import optim_tf1
    return dict(optimizer=lambda lr: optimizers.MomentumOptimizer(lr, 0.9)), optim_tf1.Momentum(0.9)


@pytest.mark.parametrize('pipelined', [False, True])
@pytest.mark.parametrize('fused_small', [True, False])
def test_cosmo_like_fcn_statistics_head(fused_small, pipelined, monkeypatch):
    """config 2 in miniature: partial sphere (order 2), K=5, max-pool 4, BN everywhere, last layer
    linear, statistics='mean', M=[]  (experiments/cosmo/experiment_deepsphere.py:39-47); with the
    input layer on the fused never-materialise-y path (K*Fin = 5) and on the generic path; with the sparse
    products of every level forced through the bulk-asynchronous pipelined kernel (csrc/sparse_pipe.cu, which
    the size threshold would reserve for the large levels) and on the gather kernels."""
    from deepsphere_tf1_b200 import ops, utils
    monkeypatch.setattr(ops, 'PIPE_MIN_ITEMS', 1 if pipelined else 1 << 30)
    nsides = [32, 16, 8, 4, 4]
    idx = utils.nside2indexes(nsides, 2)
    L, p = utils.build_laplacians(nsides, indexes=idx, new=False)
    assert p == [4, 4, 4, 1] and [l.shape[0] for l in L] == [256, 64, 16, 4]
    oracle, model = _pair(L, F=[16, 32, 64, 2], K=[5] * 4, p=p, batch_norm=[True] * 4, M=[], statistics='mean',
                          fused_small=fused_small)
    assert model._encoder[0].small == fused_small
    rs = np.random.RandomState(0)
    x = (rs.randn(4, 256, 1) * 3.6).astype(np.float32)
    labels = np.random.RandomState(1).randint(0, 2, size=4).astype(np.int32)
    _check_step(oracle, model, x, labels)


def test_shrec_like_cnn_head_full_sphere():
    """config 1 in miniature: full sphere, 6 input features, K=4, fully connected softmax layer,
    L2 regularisation, Adam with epsilon=0.1 (experiments/shrec17/hyperparameters.py:114-150)."""
    from deepsphere_tf1_b200 import optimizers, utils
    from oracle import optim_tf1
    nsides = [8, 4, 2, 1, 1]
    L, p = utils.build_laplacians(nsides, new=False)
    oracle, model = _pair(L, F=[16, 32, 64, 128], K=[4] * 4, p=p, batch_norm=[True] * 4, M=[55], num_feat_in=6,
                          statistics='mean', regularization=0.1, batch_size=5,
                          optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=0.1))
    rs = np.random.RandomState(0)
    x = rs.randn(5, 768, 6).astype(np.float32)
    labels = np.random.RandomState(1).randint(0, 55, size=5).astype(np.int32)
    _check_step(oracle, model, x, labels, opt=optim_tf1.Adam(epsilon=0.1))


@pytest.mark.parametrize('stat,M,pool', [(None, [20, 3], 'average'), ('meanvar', [3], 'max'), ('max', [3], 'max'),
                                         ('var', [3], 'average')])
def test_heads_and_pools(stat, M, pool):
    L = [healpix_L(4), healpix_L(2)]
    oracle, model = _pair(L, F=[8, 12], K=[3, 3], p=[4, 1], batch_norm=[True, False], M=M, num_feat_in=2,
                          statistics=stat, pool=pool, activation='elu' if stat == 'var' else 'relu')
    rs = np.random.RandomState(4)
    x = rs.randn(4, 192, 2).astype(np.float32)
    labels = rs.randint(0, 3, size=4).astype(np.int32)
    # fully connected layers only support relu in the product; the elu case has no hidden fc layer
    _check_step(oracle, model, x, labels)


@pytest.mark.parametrize('pool', ['max', 'average'])
def test_healpix_encoder_decoder_segmentation(pool):
    """dense=True with HEALPix pooling: unpool -> upconv -> concat skip -> deconv, outconv K=1
    (models.py:1289-1344)."""
    L = [healpix_L(4), healpix_L(2), healpix_L(1)]
    kw, opt = _momentum()
    oracle, model = _pair(L, F=[8, 16, 16], K=[3, 3, 3], p=[4, 4, 1], batch_norm=[True, True, True], M=[],
                          num_feat_in=3, pool=pool, dense=True, Fseg=3, regularization=0.05, **kw)
    rs = np.random.RandomState(6)
    x = rs.randn(4, 192, 3).astype(np.float32)
    labels = rs.randint(0, 3, size=(4, 192)).astype(np.int32)
    _check_step(oracle, model, x, labels, opt=opt)


def test_climate_like_icosahedral_segmentation():
    """config 3 in miniature: icosahedral pooling = keep the first p vertices, unpooling = pad with
    ones (models.py:1137-1138, 1356-1358); level sizes 162 -> 42 -> 12 with kNN stand-in graphs
    (the pooling logic only depends on the vertex counts)."""
    sizes = [162, 162, 42, 12]
    L = [random_knn_L(n, 6, seed=n) for n in sizes]
    p = [162, 42, 12, 12]
    # the reference pads the finest level to a hard-coded 10*4**5+2 vertices (models.py:1300), which
    # only fits level-5 meshes; the miniature overrides that constant with its own finest size
    kw, opt = _momentum()
    oracle, model = _pair(L, sampling='icosahedron', finest=162, F=[8, 16, 16, 16], K=[4] * 4, p=p,
                          batch_norm=[True] * 4, M=[], num_feat_in=4, pool='average', dense=True, Fseg=3, **kw)
    rs = np.random.RandomState(8)
    x = rs.randn(4, 162, 4).astype(np.float32)
    labels = rs.choice(3, size=(4, 162), p=[0.9, 0.02, 0.08]).astype(np.int32)
    _check_step(oracle, model, x, labels, opt=opt)


def test_ghcn_like_irregular_graph_dense_regression():
    """config 5 in miniature: irregular kNN graph (CSR path), un-rescaled combinatorial Laplacian,
    p = 1 everywhere, dense regression (experiments/ghcn/GHCN_train.py:7-32)."""
    Lg = random_knn_L(300, 10, rescale=False)
    L = [Lg] * 3
    oracle, model = _pair(L, F=[10, 20, 1], K=[5] * 3, p=[1] * 3, batch_norm=[True] * 3, M=[], num_feat_in=5,
                          regression=True, dense=True)
    assert not model._graphs[0].is_ell or model._graphs[0].max_degree <= 16
    rs = np.random.RandomState(9)
    x = rs.randn(4, 300, 5).astype(np.float32)
    labels = rs.randn(4, 300).astype(np.float32)
    _check_step(oracle, model, x, labels, rtol_after=2e-3)


def test_cuda_graph_replay_matches_eager():
    L = [healpix_L(4), healpix_L(2)]
    kw = dict(F=[8, 4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='mean')
    _, eager = _pair(L, **kw)
    _, graphed = _pair(L, cuda_graph=True, **kw)
    rs = np.random.RandomState(3)
    for t in range(4):
        x = rs.randn(4, 192, 1).astype(np.float32)
        labels = rs.randint(0, 4, size=4).astype(np.int32)
        _, l1 = eager.train_step(x, labels)
        _, l2 = graphed.train_step(x, labels)
        assert float(l1) == pytest.approx(float(l2), rel=1e-5)
    assert_close(graphed._P.w, eager._P.w, 1e-5, 'weights after graph replay')


@pytest.mark.parametrize('pipelined', [False, True])
def test_monomial_filters_classifier(pipelined, monkeypatch):
    """conv='monomials' (models.py:1085-1120): basis x, Lx, L^2 x, ... instead of the Chebyshev recurrence; same weights
    layout, same contraction.  Forward, every gradient (the adjoint A_k = G_k + L^T A_(k+1)) and three optimizer steps."""
    from deepsphere_tf1_b200 import ops, utils
    monkeypatch.setattr(ops, 'PIPE_MIN_ITEMS', 1 if pipelined else 1 << 30)
    nsides = [16, 8, 4, 4]
    L, p = utils.build_laplacians(nsides, new=False)
    oracle, model = _pair(L, F=[16, 32, 8], K=[4, 3, 2], p=p, batch_norm=[True, True, False], M=[5], statistics='meanvar',
                          conv='monomials', num_feat_in=2)
    rs = np.random.RandomState(0)
    x = rs.randn(4, L[0].shape[0], 2).astype(np.float32)
    labels = np.random.RandomState(1).randint(0, 5, size=4).astype(np.int32)
    _check_step(oracle, model, x, labels)


def test_learned_histogram_head():
    """statistics='histogram' (models.py:1263-1266, 1166-1192): 20 learned bins per feature map in front of the fully
    connected layers; the centers and widths train with the rest."""
    from deepsphere_tf1_b200 import utils
    nsides = [8, 4, 2, 2]
    L, p = utils.build_laplacians(nsides, new=False)
    oracle, model = _pair(L, F=[8, 16, 8], K=[3, 3, 3], p=p, batch_norm=[True, True, True], M=[12, 3], statistics='histogram',
                          regularization=0.02)
    assert 'stat/centers' in oracle.params and model._P.shapes['stat/widths'] == (1, 1, 8, 20)
    rs = np.random.RandomState(0)
    x = rs.randn(4, 768, 1).astype(np.float32)
    labels = rs.randint(0, 3, size=4).astype(np.int32)
    _check_step(oracle, model, x, labels)


@pytest.mark.parametrize('conv', ['chebyshev5', 'monomials'])
def test_filter_dropout(conv):
    """dropFilt < 1 (models.py:1068-1076, 1110-1118): while training, W is multiplied by a {0, 1} mask [Fin, Fout] shared by the
    K orders.  The product draws the mask on the device; the oracle gets the same masks: logits, loss and every gradient agree,
    and inference uses the full filters."""
    from deepsphere_tf1_b200 import utils
    nsides = [8, 4, 2, 2]
    L, p = utils.build_laplacians(nsides, new=False)
    keep = 0.6
    oracle, model = _pair(L, F=[8, 16, 8], K=[3, 3, 3], p=p, batch_norm=[True, False, True], M=[4], statistics='mean', conv=conv,
                          dropFilt=keep, num_feat_in=2)
    rs = np.random.RandomState(0)
    x = rs.randn(4, 768, 2).astype(np.float32)
    labels = rs.randint(0, 4, size=4).astype(np.int32)
    xd, ld = model._to_device(x), model._labels_to_device(labels)
    model._P.g.zero_()
    model._loss.zero_()
    logits, _ = model._forward(xd, training=True, save=True)
    masks = {}
    for i in range(3):
        u = model._dropout_u['conv{}/filt'.format(i + 1)]
        m = (u[:, 0, :] < keep).float().cpu()
        assert 0 < m.mean() < 1
        masks['conv{}'.format(i + 1)] = m
    total, logits_ref, grads, _ = oracle.loss_and_grads(torch.from_numpy(x), torch.from_numpy(labels), training=True, filt_masks=masks)
    assert_close(logits, logits_ref.reshape(logits.shape), RTOL, 'logits')
    model._backward(model._loss_fwd_bwd(logits, ld, True))
    assert_close(model._loss, total.reshape(1), RTOL, 'loss')
    gmax = max(float(g.abs().max()) for g in grads.values())
    for name, g in grads.items():
        assert_close(model._P.gviews[name].reshape(-1), g.reshape(-1), 2 * RTOL, 'grad ' + name, floor=1e-3 * gmax)
        if name.endswith('/weights') and name.startswith('conv'):
            mk = masks[name[:5]]
            Fin, Fout = mk.shape
            gw = model._P.gviews[name].cpu().reshape(Fin, -1, Fout)
            dropped = (mk == 0)[:, None, :].expand_as(gw)
            assert float(gw[dropped].abs().max()) == 0.0                # dropped filters get no gradient
    # inference: full filters (the moving statistics were updated by the training forward on both sides)
    for k, v in oracle.state.items():
        model.set_var(k, v.numpy())
    ref, _ = oracle.predict_logits(torch.from_numpy(x))
    out, _ = model._forward(xd, training=False, save=False)
    assert_close(out, ref.reshape(out.shape), 5e-4, 'eval logits')


def test_weighted_cross_entropy_dense_segmentation():
    """weighted=True (models.py:795-802): per-pixel cross-entropy weighted by the hard-coded class weights of the climate
    data set (3 classes), on a small encoder-decoder."""
    L = [healpix_L(4), healpix_L(2), healpix_L(1)]
    kw, opt = _momentum()
    oracle, model = _pair(L, F=[8, 16, 16], K=[3, 3, 3], p=[4, 4, 1], batch_norm=[True, True, True], M=[], Fseg=3, dense=True,
                          num_feat_in=4, pool='average', weighted=True, **kw)
    rs = np.random.RandomState(0)
    x = rs.randn(4, 192, 4).astype(np.float32)
    labels = rs.choice(3, size=(4, 192), p=[0.9, 0.02, 0.08]).astype(np.int32)
    _check_step(oracle, model, x, labels, opt=opt, rtol_after=2e-3)
    # the weights are defined for 3 classes only
    _, bad = _pair(L[:2], F=[8, 4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='mean', weighted=True)
    with pytest.raises(ValueError):
        bad.train_step(rs.randn(4, 192, 1).astype(np.float32), rs.randint(0, 4, size=4).astype(np.int32))


@pytest.mark.parametrize('graphed', [False, True])
def test_prefetched_batches_train_exactly_like_plain_steps(graphed):
    """train_step(next_batch=...) stages the following batch on a copy stream while the step runs (what fit does);
    the weights after a few steps must equal those of plain train_step calls on the same batches."""
    L = [healpix_L(4), healpix_L(2)]
    kw = dict(F=[8, 4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='mean', cuda_graph=graphed)
    _, plain = _pair(L, **kw)
    _, pre = _pair(L, **kw)
    rs = np.random.RandomState(5)
    batches = [(torch.from_numpy(rs.randn(4, 192, 1).astype(np.float32)).pin_memory(),
                torch.from_numpy(rs.randint(0, 4, size=4).astype(np.int32)).pin_memory()) for _ in range(5)]
    for t, (x, labels) in enumerate(batches):
        _, l1 = plain.train_step(x, labels)
        _, l2 = pre.train_step(x, labels, next_batch=batches[t + 1] if t + 1 < len(batches) else None)
        assert pre._prefetched is None or pre._prefetched[0] is batches[t + 1][0]
        assert float(l1) == pytest.approx(float(l2), rel=1e-5)       # (atomic accumulation order differs run to run)
    assert_close(pre._P.w, plain._P.w, 1e-5, 'weights after prefetched steps')
    # pageable numpy batches (the reference's dataset protocol) go through the pinned staging buffers
    _, plain2 = _pair(L, **kw)
    _, pre2 = _pair(L, **kw)
    nb = [(b[0].numpy().copy(), b[1].numpy().copy()) for b in batches]
    for t, (x, labels) in enumerate(nb):
        plain2.train_step(x, labels)
        pre2.train_step(x, labels, next_batch=nb[t + 1] if t + 1 < len(nb) else None)
    assert_close(pre2._P.w, plain2._P.w, 1e-5, 'weights after prefetched numpy batches')
    # a caller that passes a different batch than the one it announced just gets a plain copy
    x, labels = batches[0]
    pre.train_step(x, labels, next_batch=batches[1])
    plain.train_step(x, labels)
    pre.train_step(batches[2][0], batches[2][1])
    plain.train_step(batches[2][0], batches[2][1])
    assert_close(pre._P.w, plain._P.w, 1e-5, 'weights after an unannounced batch')


@pytest.mark.parametrize('graphed', [False, True])
def test_loss_handle_delivers_the_loss_of_every_step(graphed):
    """`loss_handle()` (what fit() and bench.py's e2e loop read): the loss of step i, copied into a pinned slot before step
    i + 1 clears the accumulator, read one step behind -- the same numbers as synchronous reads, in the same order."""
    L = [healpix_L(4), healpix_L(2)]
    kw = dict(F=[8, 4], K=[3, 3], p=[4, 1], batch_norm=[True, True], M=[], statistics='mean', cuda_graph=graphed)
    _, m = _pair(L, **kw)
    rs = np.random.RandomState(8)
    batches = [(torch.from_numpy(rs.randn(4, 192, 1).astype(np.float32)).pin_memory(),
                torch.from_numpy(rs.randint(0, 4, size=4).astype(np.int32)).pin_memory()) for _ in range(6)]
    sync, lagged, pending = [], [], None
    for t, (x, labels) in enumerate(batches):
        _, loss = m.train_step(x, labels, next_batch=batches[t + 1] if t + 1 < len(batches) else None)
        handle = m.loss_handle()
        if t % 2 == 0:
            sync.append(float(loss))              # a synchronous read in between must not disturb the slots
        else:
            sync.append(None)
        if pending is not None:
            lagged.append(pending.value())
        pending = handle
    lagged.append(pending.value())
    assert len(lagged) == len(batches) and len(set(lagged)) == len(batches)
    for a, b in zip(sync, lagged):
        assert a is None or a == b
    assert lagged[-1] == float(m._loss)


@pytest.mark.parametrize('shard', [False, True])
def test_fit_with_one_sample_per_step(shard):
    """batch_size = 1 (the only batch size of the vertex-sharded mode): `LabeledDataset.iter(1)` yields un-batched samples like
    the reference (data.py:58-60); fit / predict(cache=True) restore the batch axis.  `shard`: the vertex-sharded layer program
    on a single rank (no ghosts, no exchange)."""
    from deepsphere_tf1_b200 import models, optimizers
    from deepsphere_tf1_b200.data import LabeledDataset
    rs = np.random.RandomState(3)
    n = 12
    labels = rs.randint(0, 2, size=n)
    x = rs.randn(n, 192).astype(np.float32) + labels[:, None] * 1.5
    train, val = LabeledDataset(x[:8], labels[:8]), LabeledDataset(x[8:], labels[8:], shuffle=False)
    model = models.deepsphere(nsides=[4, 2, 2], new=False, F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean',
                              num_epochs=2, batch_size=1, eval_frequency=8, seed=1, verbose=False, vertex_shard=shard,
                              scheduler=lambda step: 1e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                              dir_name='_test_fit_b1')
    acc, loss_val, loss_train, t_step, t_batch = model.fit(train, val, restore=False, verbose=False)
    assert len(acc) == len(loss_val) == len(loss_train) == 2 and np.isfinite(loss_train).all() and np.isfinite(loss_val).all()
    pred, loss = model.predict(val, sess=model, cache=True)              # labels come from the dataset
    assert pred.shape == (4,) and np.isfinite(loss)
    pred2 = model.predict(x[8:].reshape(4, 192, 1), sess=model)
    np.testing.assert_array_equal(pred, pred2)


def test_fit_predict_evaluate_api(tmp_path):
    """Reference-style end-to-end use: datasets in, tuples out (models.py:432-618, 343-430)."""
    from deepsphere_tf1_b200 import models, optimizers
    from deepsphere_tf1_b200.data import LabeledDataset
    rs = np.random.RandomState(0)
    n = 40
    labels = rs.randint(0, 2, size=n)
    x = rs.randn(n, 192).astype(np.float32) + labels[:, None] * 1.5          # separable by the mean
    train, val = LabeledDataset(x[:32], labels[:32]), LabeledDataset(x[32:], labels[32:], shuffle=False)
    model = models.deepsphere(nsides=[4, 2, 2], F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean',
                              num_epochs=12, batch_size=8, eval_frequency=16, seed=1, verbose=False,
                              scheduler=lambda step: 2e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                              dir_name='_test_fit')
    acc, loss_val, loss_train, t_step, t_batch = model.fit(train, val, restore=False, verbose=False)
    assert len(acc) == len(loss_val) == len(loss_train) == 3 and t_step > 0 and t_batch > 0
    assert loss_train[-1] < loss_train[0]
    pred = model.predict(x[32:].reshape(8, 192, 1), sess=model)
    assert pred.shape == (8,)
    pred2, loss = model.predict(x[32:].reshape(8, 192, 1), labels[32:], sess=model)
    np.testing.assert_array_equal(pred, pred2)
    string, accuracy, f1, loss2, metrics = model.evaluate(x[32:].reshape(8, 192, 1), labels[32:], sess=model)
    assert 'accuracy' in string and loss2 == pytest.approx(loss) and accuracy >= 75
    # TensorBoard event file with the reference's tags (models.py:575-603, 815-819, 868), one record per evaluation
    import glob
    from deepsphere_tf1_b200 import summary
    files = glob.glob(model._get_path('summaries') + '/events.out.tfevents.*')
    assert len(files) == 1
    ev = summary.read_events(files[0])
    assert [e[0] for e in ev] == [16, 32, 48]
    for tag in ('loss/total', 'learning_rate', 'validation/accuracy', 'validation/f1', 'validation/loss', 'conv1/weights',
                'conv1/weights/gradients', 'conv1/bias'):
        assert tag in ev[-1][1], tag
    assert ev[-1][1]['validation/loss'] == pytest.approx(loss_val[-1], rel=1e-5)
    # the evaluation string against the same metrics computed from the ORACLE's predictions (models.py:421-427)
    import sklearn.metrics
    from oracle.model import OracleCgcnn
    oracle = OracleCgcnn(model.L, F=[8, 2], K=[3, 3], p=model.p, batch_norm=[True, True], M=[], statistics='mean')
    for name in oracle.params:
        oracle.params[name] = torch.from_numpy(model.get_var(name)).reshape(oracle.params[name].shape)
    for name in oracle.state:
        oracle.state[name] = torch.from_numpy(model.get_var(name))
    ologits, _ = oracle.predict_logits(torch.from_numpy(x[32:].reshape(8, 192, 1).astype(np.float32)))
    opred = ologits.argmax(-1).numpy()
    oloss, _ = oracle.loss(ologits, torch.from_numpy(labels[32:]), oracle.params)
    expect = 'accuracy: {:.2f} ({:d} / {:d}), f1 (weighted): {:.2f}, loss: {:.2e}'.format(
        100 * sklearn.metrics.accuracy_score(labels[32:], opred), int((opred == labels[32:]).sum()), 8,
        100 * sklearn.metrics.f1_score(labels[32:], opred, average='weighted'), float(oloss))
    assert string.split('\n')[0] == expect
    # a fresh model restores the latest checkpoint when no session is given (models.py:851-860)
    again = models.deepsphere(nsides=[4, 2, 2], F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean',
                              num_epochs=12, batch_size=8, eval_frequency=16, seed=5, verbose=False,
                              scheduler=lambda step: 2e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                              dir_name='_test_fit')
    np.testing.assert_array_equal(again.predict(x[32:].reshape(8, 192, 1)), pred)
    assert again.get_var('conv1/weights').shape == (3, 8)
    import shutil
    shutil.rmtree(model._get_path('checkpoints'), ignore_errors=True)
    shutil.rmtree(model._get_path('summaries'), ignore_errors=True)


def test_climate_encoder_decoder_on_the_real_icosahedral_meshes():
    """config 3 through the reference's own entry point: `deepsphere(sampling='icosahedron', nsides=[5, 5, 4, 3, 3])` builds the
    level-5 / 4 / 3 meshes (graphs.py, pinned to the reference's SphereIcosahedron) and their combinatorial Laplacians
    (utils.py:370-376, 412); prefix-slice pooling, pad-with-ones unpooling up to the hard-coded 10,242 vertices, weighted CE."""
    from deepsphere_tf1_b200 import models, optimizers
    from oracle.model import OracleCgcnn
    from oracle import optim_tf1
    kw = dict(F=[8, 16, 16, 16], K=[4] * 4, batch_norm=[True] * 4, M=[], num_feat_in=4, pool='average', dense=True, Fseg=3,
              weighted=True)
    model = models.deepsphere(nsides=[5, 5, 4, 3, 3], sampling='icosahedron', num_epochs=1, batch_size=2, eval_frequency=10,
                              dir_name='_test_parity', verbose=False, scheduler=lambda step: 1e-2,
                              optimizer=lambda lr: optimizers.MomentumOptimizer(lr, 0.9), **kw)
    assert [l.shape[0] for l in model.L] == [10242, 10242, 2562, 642] and model.p == [10242, 2562, 642, 642]
    oracle = OracleCgcnn(model.L, p=model.p, sampling='icosahedron', **kw)
    copy_weights(oracle, model)
    rs = np.random.RandomState(9)
    x = rs.randn(2, 10242, 4).astype(np.float32)
    labels = rs.choice(3, size=(2, 10242), p=[0.9, 0.02, 0.08]).astype(np.int32)
    # 14 stages deep: the biases in front of batch-normalised filters collect ~1e-8 of rounding noise over the momentum steps
    _check_step(oracle, model, x, labels, opt=optim_tf1.Momentum(0.9), wfloor=1e-4)


def test_default_healpix_graph_is_the_knn_graph():
    """`deepsphere(nsides)` without `new=`: the upstream default new=True (models.py:1827) -> kNN graph, 11-wide ELL rows."""
    from deepsphere_tf1_b200 import models, optimizers
    from oracle.model import OracleCgcnn
    kw = dict(F=[8, 4], K=[4, 3], batch_norm=[True, True], M=[5], statistics='meanvar')
    model = models.deepsphere(nsides=[8, 4, 2], num_epochs=1, batch_size=3, eval_frequency=10, dir_name='_test_parity',
                              verbose=False, scheduler=lambda step: 1e-2,
                              optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), **kw)
    assert np.diff(model.L[0].tocsr().indptr).max() > 9
    oracle = OracleCgcnn(model.L, p=model.p, **kw)
    copy_weights(oracle, model)
    rs = np.random.RandomState(10)
    _check_step(oracle, model, rs.randn(3, 768, 1).astype(np.float32), rs.randint(0, 5, size=3).astype(np.int32))


def test_masked_dense_regression():
    """`mask=[train, val]` (models.py:1030-1032): the masked MSE of models.py:785-787 -- predictions * data[..., -2],
    labels * data[..., -1] -- on a GHCN-like dense regression net."""
    L = [random_knn_L(300, 8, seed=3)] * 3
    kw, opt = _momentum()
    oracle, model = _pair(L, F=[12, 12, 1], K=[3, 3, 3], p=[1, 1, 1], batch_norm=[True, True, True], M=[], num_feat_in=4,
                          regression=True, dense=True, mask=[None, None], **kw)
    oracle.masked = True
    rs = np.random.RandomState(12)
    x = rs.randn(4, 300, 4).astype(np.float32)
    x[..., -2:] = (rs.rand(4, 300, 2) > 0.3).astype(np.float32)
    labels = rs.randn(4, 300).astype(np.float32)
    _check_step(oracle, model, x, labels, opt=opt)


@pytest.mark.parametrize('head', ['mean', 'flatten_fc', 'dense'])
def test_equiangular_grid_2d_pooling(head):
    """sampling='equiangular' (models.py:1131-1136, 1150-1155, 1348-1355): 2-D pooling / unpooling of the (colatitude, longitude)
    grid.  The product keeps the vertices in Z order and pools consecutive vertices; inputs, descriptors, the flattened head
    and dense outputs are in the reference's row-major order, which this test pins against the oracle."""
    from deepsphere_tf1_b200 import utils
    L, p = utils.build_laplacians([8, 4, 2, 2], sampling='equiangular')
    assert p == [4, 4, 1] and [l.shape[0] for l in L] == [256, 64, 16]
    rs = np.random.RandomState(5)
    x = rs.randn(3, 256, 2).astype(np.float32)
    if head == 'dense':
        kw = dict(F=[6, 8, 4], K=[3, 3, 2], p=p, batch_norm=[True, True, False], M=[], pool='average', dense=True, Fseg=3,
                  num_feat_in=2)
        labels = rs.randint(0, 3, size=(3, 256)).astype(np.int32)
        mom, opt = _momentum()
        kw.update(mom)
    elif head == 'flatten_fc':
        kw = dict(F=[6, 8], K=[3, 3], p=p[:2], batch_norm=[True, True], M=[12, 3], pool='max', statistics=None, num_feat_in=2)
        labels = rs.randint(0, 3, size=3).astype(np.int32)
        opt = None
        L = L[:2]
    else:
        kw = dict(F=[6, 8, 3], K=[3, 3, 2], p=p, batch_norm=[True, True, False], M=[], pool='max', statistics='mean', num_feat_in=2)
        labels = rs.randint(0, 3, size=3).astype(np.int32)
        opt = None
    oracle, model = _pair(L, sampling='equiangular', batch_size=3, **kw)
    _check_step(oracle, model, x, labels.astype(np.int64) if head != 'dense' else labels.astype(np.int64), opt=opt)


def test_rectangular_equiangular_grid_through_the_front_end():
    """deepsphere(nsides=[(bw_colatitude, bw_longitude), ...], sampling='equiangular'): `ratio` = columns / rows
    (models.py:1832-1836) drives the reshape of the 2-D pooling (models.py:1134).  16 x 24 grid, ratio 1.5 (the aspect of the
    climate data's 768 x 1152 grid), pooled twice; the vertices go through the blocked Z order of
    utils.equiangular_block_order.  (The operator code is pinned to the reference by the golden cases equi_rect_*.)"""
    from deepsphere_tf1_b200 import models, optimizers
    from oracle.model import OracleCgcnn
    nsides = [(8, 12), (4, 6), (2, 3), (2, 3)]
    kw = dict(F=[6, 8, 3], K=[3, 3, 2], batch_norm=[True, True, False], M=[], pool='max', statistics='mean', num_feat_in=2)
    model = models.deepsphere(nsides=nsides, sampling='equiangular', num_epochs=1, batch_size=3, eval_frequency=10,
                              dir_name='_test_parity', verbose=False,
                              optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8),
                              scheduler=lambda step: optimizers.exponential_decay(2e-3, step, 10, 0.9), **kw)
    assert model.ratio == 1.5 and list(model.p) == [4, 4, 1] and [l.shape[0] for l in model.L] == [384, 96, 24]
    oracle = OracleCgcnn(model.L, p=model.p, sampling='equiangular', ratio=model.ratio, seed=2, **kw)
    copy_weights(oracle, model)
    rs = np.random.RandomState(6)
    x = rs.randn(3, 384, 2).astype(np.float32)
    _check_step(oracle, model, x, rs.randint(0, 3, size=3).astype(np.int64))


def test_use_4_neighbour_graph_model():
    """`use_4=True` (utils.py:336-337, 484-591): the 4-neighbour patch graph (ELL width 5) through the whole model."""
    from deepsphere_tf1_b200 import utils
    nsides = [16, 8, 4, 4]
    idx = utils.nside2indexes(nsides, 1)
    L, p = utils.build_laplacians(nsides, indexes=idx, use_4=True, new=False)
    assert p == [4, 4, 1] and [l.shape[0] for l in L] == [256, 64, 16]
    oracle, model = _pair(L, F=[8, 16, 3], K=[4, 4, 3], p=p, batch_norm=[True, True, False], M=[], statistics='mean', num_feat_in=2)
    rs = np.random.RandomState(9)
    x = rs.randn(4, 256, 2).astype(np.float32)
    labels = rs.randint(0, 3, size=4).astype(np.int32)
    _check_step(oracle, model, x, labels)
