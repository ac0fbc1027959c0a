"""Oracle (CPU) and CUDA path (GPU) against tests/golden/ops_golden.npz: the outputs of the reference's own, unmodified
deepsphere/models.py graph-construction code (chebyshev5, monomials, pool_*, unpool_*, batch_normalization, bias, fc,
learned_histogram, _inference, _decoder, loss, training), executed through the torch-backed TensorFlow stand-in of
tests/golden/tf_shim.py by tests/golden/make_ops_golden.py.  Tolerances: fp32 forward / gradients, rtol 1e-4 of the
largest expected magnitude (north star); the oracle itself is held to 2e-5."""
import json
import os

import numpy as np
import pytest
import torch
from scipy import sparse

from oracle.model import OracleCgcnn

from helpers import assert_close

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ops_golden.npz'), allow_pickle=False)
CASES = sorted({k.split('::')[0] for k in GOLD.files})


class Case(object):
    def __init__(self, name):
        self.name = name
        self.cfg = json.loads(str(self['config']))
        self.var_names = json.loads(str(self['var_names']))
        self.L = []
        for i in range(self.cfg['n_levels']):
            n = int(self['L%d_n' % i])
            self.L.append(sparse.coo_matrix((self['L%d_data' % i], (self['L%d_row' % i], self['L%d_col' % i])), shape=(n, n)))
        self.x, self.labels = self['x'], self['labels']
        self.training = self.cfg['training']

    def __getitem__(self, key):
        return GOLD['{}::{}'.format(self.name, key)]

    def has(self, key):
        return '{}::{}'.format(self.name, key) in GOLD.files

    def model_kwargs(self):
        keys = ('F', 'K', 'batch_norm', 'M', 'num_feat_in', 'conv', 'pool', 'activation', 'statistics', 'Fseg',
                'regularization', 'dropout', 'regression', 'dense', 'weighted')
        return {k: self.cfg[k] for k in keys if k in self.cfg}

    def states(self):
        return {k.split('::state/')[1]: GOLD[k] for k in GOLD.files if k.startswith(self.name + '::state/')}

    def masks(self):
        return [torch.from_numpy(self['dropout_mask%d' % i]) for i in range(int(self['n_dropout_masks']))]

    def split_masks(self):
        """The graph draws the filter masks of the ChebConv layers first (in layer order), then the fc dropout masks."""
        masks = self.masks()
        filt, n = {}, 0
        if self.cfg.get('dropFilt', 1) < 1 and self.training:
            for i in range(len(self.cfg['F'])):
                filt['conv%d' % (i + 1)] = masks[n]
                n += 1
        return filt, masks[n:]


def build_oracle(c):
    o = OracleCgcnn(c.L, p=c.cfg['p'], sampling=c.cfg['sampling'], ratio=c.cfg.get('ratio', 1.), masked=bool(c.cfg.get('mask')),
                    **c.model_kwargs())
    assert list(o.params.keys()) == c.var_names                         # the reference's variable names and creation order
    for k in c.var_names:
        v = torch.from_numpy(c['var/' + k])
        assert tuple(o.params[k].shape) == tuple(v.shape), (k, tuple(o.params[k].shape), tuple(v.shape))
        o.params[k] = v.clone()
    for k, v in c.states().items():
        assert k in o.state, k
        o.state[k] = torch.from_numpy(v).clone()
    return o


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_graph(name):
    c = Case(name)
    o = build_oracle(c)
    x = torch.from_numpy(c.x)
    labels = torch.from_numpy(c.labels)
    filt, drop = c.split_masks()
    tol = 2e-5
    if c.training:
        total, logits, grads, new_state = o.loss_and_grads(x, labels, True, drop_masks=drop or None, filt_masks=filt or None)
        assert_close(logits, c['logits'], tol, name + ' logits')
        assert abs(float(total) - float(c['loss'])) <= tol * max(1.0, abs(float(c['loss']))), (float(total), float(c['loss']))
        gmax = max(np.abs(c['grad/' + k]).max() for k in c.var_names)
        for k in c.var_names:
            assert_close(grads[k], c['grad/' + k], 5 * tol, name + ' grad ' + k, floor=1e-3 * gmax)
        for k in c.states():
            if c.has('update/' + k):
                assert_close(new_state[k], c['update/' + k], tol, name + ' ' + k)
    else:
        logits, desc = o.predict_logits(x)
        assert_close(logits, c['logits'], tol, name + ' logits')
        assert_close(desc, c['descriptor'], tol, name + ' descriptor')
        total, _ = o.loss(logits, labels, o.params)
        assert abs(float(total) - float(c['loss'])) <= tol * max(1.0, abs(float(c['loss'])))


def test_monomial_regulariser_uses_default_stddev():
    """models.py:1108 -> 862-867: `monomials` weights come from `_weight_variable` with stddev 0.1, for the draw and for the
    regulariser l2_loss(W) / stddev^2 (the Chebyshev std would change the loss)."""
    c = Case('monomials')
    o = build_oracle(c)
    assert all(abs(s - 0.1) < 1e-12 for s in o.std.values())
    W = [c['var/' + k] for k in c.var_names if k.endswith('weights')]
    reg = c.cfg['regularization'] * sum((w.astype(np.float64) ** 2).sum() / 2 / 0.1 ** 2 for w in W) / sum(w.size for w in W)
    logits = torch.from_numpy(c['logits'])
    ce = torch.nn.functional.cross_entropy(logits, torch.from_numpy(c.labels).long()).item()
    assert abs(ce + reg - float(c['loss'])) < 1e-4 * float(c['loss'])
    # the shim drew the weights with that stddev too: truncated normal => |w| <= 2 * 0.1
    assert max(np.abs(w).max() for w in W) <= 0.2 + 1e-6


# ------------------------------------------------------------------------------------------ CUDA path
def build_product(c, dev):
    from deepsphere_tf1_b200 import models, optimizers
    kw = c.model_kwargs()
    if c.cfg.get('mask'):                               # only the presence of the argument matters (models.py:1030-1032, 785-787)
        kw['mask'] = [np.ones(c.L[0].shape[0]), np.ones(c.L[0].shape[0])]
    cls = models.cgcnn
    if c.cfg['sampling'] != 'healpix':
        cls = type('cgcnn_' + c.cfg['sampling'], (models.cgcnn,), {'sampling': c.cfg['sampling'], 'ratio': c.cfg.get('ratio', 1.)})
    m = cls(L=c.L, p=c.cfg['p'], num_epochs=1, scheduler=lambda step: 1e-3,
            optimizer=lambda lr: optimizers.GradientDescentOptimizer(lr), batch_size=c.x.shape[0], dir_name='_golden',
            verbose=False, device=dev, **kw)
    assert [s[0] for s in m._P.specs] == c.var_names
    for k in c.var_names:
        m.set_var(k, c['var/' + k])
    for k, v in c.states().items():
        m.set_var(k, v)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize('name', [n for n in CASES if n != 'dropouts'])
def test_cuda_path_matches_reference_graph(name, dev):
    """Same inputs, same weights: logits, loss, every gradient and the batch-norm moving statistics of the CUDA path
    against the reference's own graph code."""
    from deepsphere_tf1_b200 import ops
    c = Case(name)
    m = build_product(c, dev)
    x = m._to_device(c.x)
    labels = m._labels_to_device(c.labels)
    tol = 1e-4
    if not c.training:
        logits, desc = m._forward(x, training=False, save=False)
        assert_close(logits, c['logits'], tol, name + ' logits')
        assert_close(desc, c['descriptor'], tol, name + ' descriptor')
        assert abs(m._loss_only(logits, labels) - float(c['loss'])) <= tol * max(1.0, abs(float(c['loss'])))
        return
    ops.fill(m._P.g, 0.0)
    m._loss.zero_()
    logits, _ = m._forward(x, training=True, save=True)
    dlogits = m._loss_fwd_bwd(logits, labels, True)
    m._backward(dlogits)
    if m._P.has_reg:
        ops.l2_reg(m._P.w, m._P.coef, m._P.g, m._loss)
    assert_close(logits.reshape(c['logits'].shape), c['logits'], tol, name + ' logits')
    assert abs(float(m._loss) - float(c['loss'])) <= tol * max(1.0, abs(float(c['loss']))), (float(m._loss), float(c['loss']))
    gmax = max(np.abs(c['grad/' + k]).max() for k in c.var_names)
    for k in c.var_names:
        assert_close(m._P.gviews[k].reshape(c['grad/' + k].shape), c['grad/' + k], 5 * tol, name + ' grad ' + k, floor=1e-3 * gmax)
    for k in c.states():
        if c.has('update/' + k):
            assert_close(m._state[k], c['update/' + k], tol, name + ' ' + k)
