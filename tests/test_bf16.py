"""BF16 mode (cgcnn(dtype='bfloat16'), north star: "within rtol 1e-4 in FP32 (1e-2 in BF16)").

Kernels: every bf16 kernel against a float64 evaluation of the SAME bf16-representable inputs; what remains is the one rounding of
the stored result (2^-9 relative per element) -- tolerance 5e-3 of the largest expected magnitude; the weight gradient (fp32
output, fp32 accumulation in TMEM) to 1e-5.
Models: the forward pass (logits, loss) within 1e-2 of the float64 oracle.  Gradients: BF16 storage flips arg-max / ReLU decisions
of a max-pooling network (the loss is not differentiable there), so their deviation is held to the noise floor of BF16 storage
itself: the CPU oracle evaluated with the same tensors rounded to bfloat16 (oracle.ops.STORE)."""
import numpy as np
import pytest
import torch

from helpers import assert_close, healpix_L

BF = torch.bfloat16
RT_STORE = 5e-3          # one bf16 rounding of a stored activation, relative to the largest expected magnitude
RT_FWD = 1e-2            # north star


def _bfr(t):
    return t.to(BF).to(torch.float32)


def test_oracle_bf16_storage_emulation_is_off_by_default_and_differentiable():
    from oracle.model import OracleCgcnn
    from deepsphere_tf1_b200 import utils
    nsides = [8, 4, 4]
    L, p = utils.build_laplacians(nsides, indexes=utils.nside2indexes(nsides, 1), new=False)
    kw = dict(F=[8, 2], K=[3, 3], p=p, batch_norm=[True, True], M=[], statistics='mean')
    a, b = OracleCgcnn(L, **kw), OracleCgcnn(L, **kw)
    b.bf16_layers = {2}
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.randn(3, L[0].shape[0], 1).astype(np.float32))
    lab = torch.from_numpy(rs.randint(0, 2, 3))
    ta, la, ga, _ = a.loss_and_grads(x, lab)
    tb, lb, gb, _ = b.loss_and_grads(x, lab)
    ta2, la2, _, _ = a.loss_and_grads(x, lab)
    assert torch.equal(la, la2)                                     # the hook leaves no state behind
    assert not torch.equal(la, lb) and float((la - lb).abs().max()) < 2e-2 * float(la.abs().max())
    for k in ga:
        assert torch.isfinite(gb[k]).all()
    assert float((ga['conv1/weights'] - gb['conv1/weights']).norm()) < 0.2 * float(ga['conv1/weights'].norm())


@pytest.mark.gpu
def test_casts(dev):
    from deepsphere_tf1_b200 import ops
    for n in (1, 3, 4, 1000003):
        t = torch.randn(n, device=dev)
        b = ops.to_bf16(t)
        assert b.dtype == BF and torch.equal(b, t.to(BF))
        assert torch.equal(ops.to_f32(b), t.to(BF).float())


@pytest.mark.gpu
@pytest.mark.parametrize('dense', [True, False])
@pytest.mark.parametrize('F', [16, 32, 64])
@pytest.mark.parametrize('nside,order', [(16, 0), (32, 2)])
def test_bf16_pipelined_sparse_product(dev, nside, order, F, dense, monkeypatch):
    import scipy.sparse as sp
    from deepsphere_tf1_b200 import ops
    from deepsphere_tf1_b200.graph import DeviceGraph
    monkeypatch.setattr(ops, 'USE_DENSE_GROUPS', dense)
    L = healpix_L(nside, order)
    g = DeviceGraph(L, dev)
    M, B = L.shape[0], 3
    assert ops._pipe(g, 'fwd', B, M, F, esize=2, min_items=1) is not None
    Ls = sp.csr_matrix(L).astype(np.float64)
    gen = torch.Generator().manual_seed(F)
    x, y, z = [_bfr(torch.randn(B, M, F, generator=gen)) for _ in range(3)]
    for tr in (False, True):
        Lm = Ls.T.tocsr() if tr else Ls
        exp = 2.0 * torch.from_numpy(np.stack([Lm @ x[b].double().numpy() for b in range(B)])) - y.double() + 0.5 * z.double()
        out = ops.spmm(g, x.to(dev).to(BF), y.to(dev).to(BF), z.to(dev).to(BF), alpha=2.0, beta=-1.0, gamma=0.5, transpose=tr)
        assert out.dtype == BF
        assert_close(out.float(), exp, RT_STORE, 'spmm bf16 transpose={}'.format(tr))
    # the recurrence and its adjoint (in place on the planes of G, out aliasing y)
    K = 5
    basis = ops.cheb_basis(g, x.to(dev).to(BF), K)
    planes = [x.double()]
    planes.append(torch.from_numpy(np.stack([Ls @ planes[0][b].numpy() for b in range(B)])))
    for k in range(2, K):
        planes.append(2 * torch.from_numpy(np.stack([Ls @ planes[-1][b].numpy() for b in range(B)])) - planes[-2])
    assert_close(basis.float(), torch.stack(planes[1:]), 2 * RT_STORE, 'basis bf16')      # roundings accumulate over the orders
    G = _bfr(torch.randn(K, B, M, F, generator=gen))
    LT = Ls.T.tocsr()
    A = [None] * K
    A[K - 1] = G[K - 1].double()
    for k in range(K - 2, -1, -1):
        LA = torch.from_numpy(np.stack([LT @ A[k + 1][b].numpy() for b in range(B)]))
        A[k] = G[k].double() + (2.0 if k >= 1 else 1.0) * LA - (A[k + 2] if k + 2 <= K - 1 else 0.0)
    dx = ops.cheb_basis_bwd(g, G.to(dev).to(BF).contiguous())
    assert_close(dx.float(), A[0], 3 * RT_STORE, 'adjoint recurrence bf16')


@pytest.mark.gpu
@pytest.mark.parametrize('R,Fin,K,Fout', [(1024, 16, 5, 32), (1000, 16, 5, 32), (4096, 32, 5, 64), (2048, 64, 5, 64), (512, 64, 5, 2),
                                          (2048, 128, 3, 64), (768, 16, 1, 16), (3072, 32, 4, 32), (4096, 16, 5, 64), (300, 6, 3, 5),
                                          (65536, 16, 5, 32)])
def test_bf16_contraction_three_directions(dev, R, Fin, K, Fout):
    """tcgen05 kind::f16 kernels where dsb200_contract_bf16_supported says so, fp32 kernels between casts elsewhere: same
    results either way.  Reference: float64 products of the bf16-rounded operands (the kernels round the weights to bf16)."""
    from deepsphere_tf1_b200 import ops
    gen = torch.Generator().manual_seed(R + Fin)
    x = _bfr(torch.randn(K, R, Fin, generator=gen))
    W = torch.randn(Fin * K, Fout, generator=gen) * 0.2
    dy = _bfr(torch.randn(R, Fout, generator=gen))
    native = [ops._lib_int('contract_bf16_supported', d, R, Fin, K, Fout) for d in range(3)]
    Wd = W.double().reshape(Fin, K, Fout)
    Wb = _bfr(W).double().reshape(Fin, K, Fout)
    xb = x.to(dev).to(BF).contiguous()
    x0, xk = xb[0].view(1, R, Fin), (xb[1:].view(K - 1, 1, R, Fin) if K > 1 else None)
    dyb = dy.to(dev).to(BF).view(1, R, Fout)
    sums = torch.zeros(2 * Fout, dtype=torch.float64, device=dev)
    y = ops.contract_fwd(x0, xk, W.to(dev), K, Fout, sums)
    assert y.dtype == BF
    assert_close(y.view(R, Fout).float(), torch.einsum('krf,fko->ro', x.double(), Wb if native[0] else Wd), RT_STORE, 'forward')
    ys = y.view(R, Fout).double()
    if native[0]:                                                   # statistics of the stored tensor
        assert_close(sums, torch.cat([ys.sum(0), (ys * ys).sum(0)]), 1e-5, 'bn sums')
    else:
        assert_close(sums, torch.cat([ys.sum(0), (ys * ys).sum(0)]), RT_STORE, 'bn sums')
    G = ops.contract_bwd_data(dyb, W.to(dev), K, Fin)
    assert G.dtype == BF
    assert_close(G.view(K, R, Fin).float(), torch.einsum('ro,fko->krf', dy.double(), Wb if native[1] else Wd), RT_STORE, 'data gradient')
    dW = torch.zeros(Fin * K, Fout, device=dev)
    ops.contract_bwd_weight(x0, xk, dyb, dW, K)
    assert_close(dW, torch.einsum('krf,ro->fko', x.double(), dy.double()).reshape(Fin * K, Fout), 1e-5, 'weight gradient')


@pytest.mark.gpu
@pytest.mark.parametrize('B,M,F,p,pool,act', [(3, 64, 16, 4, 'max', 'relu'), (2, 256, 32, 4, 'max', 'relu'), (2, 128, 64, 1, 'max', 'relu'),
                                              (2, 96, 6, 2, 'average', 'elu'), (2, 64, 32, 4, 'average', 'relu'), (1, 60, 5, 1, 'max', 'tanh')])
def test_bf16_batch_norm_chain_equals_the_fp32_kernels_on_the_same_values(dev, B, M, F, p, pool, act):
    from deepsphere_tf1_b200 import ops
    gen = torch.Generator().manual_seed(F + p)
    y = _bfr(torch.randn(B, M, F, generator=gen))
    mean, rstd, bias = torch.randn(F, generator=gen) * 0.1, torch.rand(F, generator=gen) + 0.5, torch.randn(F, generator=gen) * 0.1
    dout = _bfr(torch.randn(B, M // p, F, generator=gen))
    a32 = (y.to(dev), mean.to(dev), rstd.to(dev), bias.to(dev))
    abf = (y.to(dev).to(BF), mean.to(dev), rstd.to(dev), bias.to(dev))
    o32, obf = ops.bn_act_pool_fwd(*a32, p, pool, act), ops.bn_act_pool_fwd(*abf, p, pool, act)
    assert obf.dtype == BF and torch.equal(obf, o32.to(BF))          # same arithmetic, one rounding
    d32, dbf = dout.to(dev), dout.to(dev).to(BF)
    g32, s32 = ops.bn_act_pool_bwd(*a32, d32, p, pool, act)
    gbf, sbf = ops.bn_act_pool_bwd(*abf, dbf, p, pool, act)
    assert torch.equal(gbf, g32.to(BF))
    assert_close(sbf, s32, 1e-6, 'sums')
    _, s1 = ops.bn_act_pool_bwd(*abf, dbf, p, pool, act, want_gz=False)
    assert_close(s1, s32, 1e-6, 'sums, pass 1 only')
    e32 = ops.bn_act_pool_bwd_dy(*a32, d32, s32, float(B * M), p, pool, act)
    ebf = ops.bn_act_pool_bwd_dy(*abf, dbf, s32, float(B * M), p, pool, act)
    assert torch.equal(ebf, e32.to(BF))
    if ops.bn_relu_pool_sums_supported(F, p, pool, act):
        l32 = ops.bn_relu_pool_sums(o32.to(BF).float(), d32, bias.to(dev))
        lbf = ops.bn_relu_pool_sums(obf, dbf, bias.to(dev))
        assert_close(lbf, l32, 1e-6, 'lean sums')


def _grads_cuda(m, x, labels):
    from deepsphere_tf1_b200 import ops
    P = m._P
    ops.fill(P.g, 0.0)
    m._loss.zero_()
    logits, _ = m._forward(m._to_device(x), training=True, save=True)
    d = m._loss_fwd_bwd(logits, m._labels_to_device(labels), True)
    m._backward(d)
    torch.cuda.synchronize()
    return logits.double().cpu(), float(m._loss), {n: P.gviews[n].detach().double().cpu().clone() for n in P.gviews}


def _flat_err(ga, ge):
    num = sum(float((ga[k].reshape(-1).double() - ge[k].reshape(-1).double()).pow(2).sum()) for k in ge)
    den = sum(float(ge[k].double().pow(2).sum()) for k in ge)
    return (num / den) ** 0.5


@pytest.mark.gpu
def test_bf16_cosmo_like_model_against_the_oracles(dev):
    """config 2 in miniature, dtype='bfloat16': layers 2..n of the encoder run on bfloat16 activations (the fused input layer stays
    fp32).  Forward within 1e-2 of the float64 oracle; gradients within 3x the deviation of the BF16-storage-emulating oracle."""
    from deepsphere_tf1_b200 import models, optimizers, utils
    from oracle.model import OracleCgcnn
    from helpers import copy_weights
    nsides = [64, 32, 16, 8, 8]
    idx = utils.nside2indexes(nsides, 2)
    L, p = utils.build_laplacians(nsides, indexes=idx, new=False)
    kw = dict(F=[16, 32, 64, 2], K=[5] * 4, p=p, batch_norm=[True] * 4, M=[], statistics='mean')
    o64 = OracleCgcnn(L, dtype=torch.float64, **kw)
    oem = OracleCgcnn(L, dtype=torch.float32, **kw)
    oem.bf16_layers = {2, 3, 4}
    model = models.cgcnn(L=L, num_epochs=1, batch_size=4, eval_frequency=10, dir_name='_test_bf16', verbose=False, dtype='bfloat16',
                         optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), scheduler=lambda s: 1e-3, **kw)
    assert [st.bf16 for st in model._encoder] == [False, True, True, True] and model._encoder[0].small
    copy_weights(oem, model)
    o64.params = {k: v.double() for k, v in oem.params.items()}
    rs = np.random.RandomState(0)
    x = (rs.randn(4, L[0].shape[0], 1) * 3.6).astype(np.float32)
    labels = rs.randint(0, 2, size=4).astype(np.int64)
    t64, l64, g64, _ = o64.loss_and_grads(torch.from_numpy(x).double(), torch.from_numpy(labels))
    tem, lem, gem, _ = oem.loss_and_grads(torch.from_numpy(x), torch.from_numpy(labels))
    logits, loss, g = _grads_cuda(model, x, labels.astype(np.int32))
    # the logits of this miniature are means over 64 vertices only (max |logit| ~ 0.5, the activations behind them are O(1)): held
    # to 1e-2 or, where BF16 storage itself costs more, to 3x the deviation of the BF16-storage oracle
    lscale = float(l64.abs().max())
    lfloor = float((lem.double() - l64).abs().max()) / lscale
    lerr = float((logits.reshape(l64.shape) - l64).abs().max()) / lscale
    print('logits error vs float64: CUDA bf16 {:.3e}, BF16-storage oracle {:.3e}'.format(lerr, lfloor))
    assert lerr <= max(RT_FWD, 3 * lfloor), (lerr, lfloor)
    assert abs(loss - float(t64)) <= max(RT_FWD, 3 * abs(float(tem) - float(t64)) / abs(float(t64))) * abs(float(t64))
    floor = _flat_err(gem, g64)                                       # what BF16 storage costs in ANY implementation
    err = _flat_err(g, g64)
    print('flat gradient error vs float64: CUDA bf16 {:.3e}, BF16-storage oracle {:.3e}'.format(err, floor))
    assert err <= max(RT_FWD, 3 * floor), (err, floor)
    # training: three Adam steps stay close to the fp32 oracle's trajectory
    from oracle import optim_tf1
    opt = optim_tf1.Adam(epsilon=1e-8)
    o32 = OracleCgcnn(L, dtype=torch.float32, **kw)
    o32.params = {k: v.clone() for k, v in oem.params.items()}
    for k, v in o32.state.items():
        model.set_var(k, v.numpy())
    for t in range(3):
        lr_ref, _ = o32.train_step(torch.from_numpy(x), torch.from_numpy(labels), opt, 1e-3)
        _, lp = model.train_step(x, labels.astype(np.int32))
        assert float(lp) == pytest.approx(lr_ref, rel=2e-2)
    # inference mode
    ref, _ = o32.predict_logits(torch.from_numpy(x))
    out, _ = model._forward(model._to_device(x), training=False, save=False)
    assert out.dtype == torch.float32
    assert_close(out, ref.reshape(out.shape), 5e-2, 'eval logits after 3 steps')


@pytest.mark.gpu
def test_bf16_encoder_decoder_and_fc_models_run_and_stay_close_to_fp32(dev):
    """Everything outside the ChebConv layers is fp32 between casts: a U-Net (unpool, concat, dense loss) and an fc head in
    dtype='bfloat16' against the same models in float32."""
    from deepsphere_tf1_b200 import models, optimizers, utils
    nsides = [8, 4, 2, 2]
    L, p = utils.build_laplacians(nsides, indexes=utils.nside2indexes(nsides, 0), new=False)
    assert [l.shape[0] for l in L] == [768, 192, 48] and p == [4, 4, 1]
    rs = np.random.RandomState(1)
    common = dict(num_epochs=1, batch_size=2, eval_frequency=10, dir_name='_test_bf16b', verbose=False, seed=3,
                  optimizer=lambda lr: optimizers.MomentumOptimizer(lr, 0.9), scheduler=lambda s: 1e-2)
    cases = [
        (dict(L=L[:3], F=[16, 32, 16], K=[3, 3, 3], p=[4, 4, 1], batch_norm=[True] * 3, M=[], dense=True, Fseg=3, num_feat_in=16,
              pool='average'),
         rs.randn(2, 768, 16).astype(np.float32), rs.randint(0, 3, size=(2, 768)).astype(np.int32)),
        (dict(L=L[:2], F=[16, 32], K=[4, 4], p=[4, 4], batch_norm=[True, True], M=[10, 3], num_feat_in=16, statistics=None),
         rs.randn(2, 768, 16).astype(np.float32), rs.randint(0, 3, size=2).astype(np.int32)),
    ]
    for kw, x, lab in cases:
        a = models.cgcnn(dtype=np.float32, **common, **kw)
        b = models.cgcnn(dtype='bfloat16', **common, **kw)
        assert any(st.bf16 for st in b._encoder)
        b._P.w.copy_(a._P.w)
        la, _, ga = _grads_cuda(a, x, lab)[0], None, None
        lb = _grads_cuda(b, x, lab)[0]
        assert_close(lb, la, 5e-2, 'logits bf16 vs fp32')              # tiny nets: max |logit| ~ 0.4 behind O(1) activations
        for _ in range(2):
            ra, rb = a.train_step(x, lab), b.train_step(x, lab)
        assert float(rb[1]) == pytest.approx(float(ra[1]), rel=3e-2)
