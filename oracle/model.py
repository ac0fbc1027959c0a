"""Oracle restatement of `cgcnn._inference` / `_decoder` / loss / one training step
(deepsphere/models.py:1219-1344, 777-843) on top of oracle.ops, torch-CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED to the reference's own code: logits, loss, every
gradient and the batch-norm updates of 21 model cases produced by the UNMODIFIED deepsphere/models.py (run through the
TensorFlow stand-in of tests/golden/tf_shim.py) are reproduced to 2e-5 (tests/test_ops_golden.py); also the parameter counts
and level sizes of experiments/cosmo/demo.ipynb:861-889.  Unpinned: the optimizer update rules (oracle/optim_tf1.py).

Variables carry the reference's TF names (`conv1/weights`, `conv1/bias`,
`conv1/batch_normalization/moving_mean`, `fc1/weights`, `logits/weights`, `upconv3/weights`,
`deconv3/weights`, `outconv/weights`; scopes at models.py:1226,1276,1284,1295,1312,1340).
"""
import numpy as np
import torch

from . import ops


class OracleCgcnn:
    def __init__(self, L, F, K, p, batch_norm, M, num_feat_in=1, conv='chebyshev5', pool='max',
                 activation='relu', statistics=None, Fseg=None, regularization=0, dropout=1,
                 regression=False, dense=False, sampling='healpix', dtype=torch.float32, seed=2,
                 finest_vertices=10 * 4 ** 5 + 2, weighted=False, masked=False, ratio=1.):
        if not len(L) == len(F) == len(K) == len(p) == len(batch_norm):          # models.py:937-940
            raise ValueError('Wrong specification of the convolutional layers')
        self.Ls = [ops.sparse_to_torch(l, dtype) for l in L]
        self.Msizes = [l.shape[0] for l in L]
        self.F, self.K, self.p, self.batch_norm, self.M = F, K, p, batch_norm, M
        self.Fseg, self.regularization, self.dropout = Fseg, regularization, dropout
        self.regression, self.dense, self.sampling = regression, dense, sampling
        self.conv = getattr(ops, conv)                                            # models.py:1017
        self.weighted = weighted
        self.ratio = ratio                          # equiangular grids: columns / rows (deepsphere.__init__, models.py:1832-1836)
        self.masked = masked                        # a `mask` argument was given: masked MSE (models.py:785-787, 1030-1032)
        self.pool_kind, self.act, self.stat = pool, ops.ACTIVATIONS[activation], statistics
        self.dtype = dtype
        self.num_feat_in = num_feat_in
        self.finest_vertices = finest_vertices      # hard-coded level-5 size at models.py:1300
        self.bf16_layers = set()                    # encoder layers (1-based) evaluated with BF16 storage emulation (ops.STORE)
        self.params = {}        # trainable, creation order
        self.std = {}           # name -> init stddev for regularised weights (models.py:866)
        self.state = {}         # BN moving statistics
        self._rs = np.random.RandomState(seed)
        self._build()

    # ---- variable creation, mirroring the order of graph construction -------------------
    def _trunc_normal(self, shape, std):
        a = self._rs.randn(*shape)
        bad = np.abs(a) > 2
        while bad.any():                                   # tf.truncated_normal: re-draw beyond 2 sigma
            a[bad] = self._rs.randn(int(bad.sum()))
            bad = np.abs(a) > 2
        return torch.tensor(a * std, dtype=self.dtype)

    def _weights(self, scope, shape, std):
        self.params[scope + '/weights'] = self._trunc_normal(shape, std)
        self.std[scope + '/weights'] = std

    def _cheb(self, scope, Fin, Fout, K):
        if self.conv is ops.monomials:
            self._weights(scope, (Fin * K, Fout), 0.1)                            # models.py:1108 -> 862 (default stddev)
        else:
            self._weights(scope, (Fin * K, Fout), 1 / np.sqrt(Fin * (K + 0.5) / 2))   # models.py:1080-1083

    def _bias(self, scope, F):
        self.params[scope + '/bias'] = torch.zeros(1, 1, F, dtype=self.dtype)

    def _bn(self, scope, F):
        self.state[scope + '/batch_normalization/moving_mean'] = torch.zeros(F, dtype=self.dtype)
        self.state[scope + '/batch_normalization/moving_variance'] = torch.ones(F, dtype=self.dtype)

    def _build(self):
        n = len(self.p)
        Fin = self.num_feat_in
        Mv = self.Msizes[0]
        self.skipF = [Fin]
        for i in range(n):
            sc = 'conv{}'.format(i + 1)
            self._cheb(sc, Fin, self.F[i], self.K[i])
            Fin = self.F[i]
            Mv = self.Msizes[i]
            if i == n - 1 and len(self.M) == 0:
                break
            if self.batch_norm[i]:
                self._bn(sc, Fin)
            self._bias(sc, Fin)
            if self.p[i] > 1:
                Mv = self.p[i] if self.sampling == 'icosahedron' else Mv // self.p[i]
            self.skipF.append(Fin)
        # statistics / fc
        if self.stat is None:
            Min = Mv * Fin
        elif self.stat == 'meanvar':
            Min = 2 * Fin
        elif self.stat == 'histogram':                       # models.py:1263-1266: 20 learned bins per feature
            Min = 20 * Fin
            c, w = ops.histogram_init(Fin, dtype=self.dtype)
            self.params['stat/centers'] = c                  # trainable, not regularised (plain tf.Variable / get_variable)
            self.params['stat/widths'] = w
        else:
            Min = Fin
        for i, Mo in enumerate(self.M[:-1]):
            sc = 'fc{}'.format(i + 1)
            self._weights(sc, (Min, Mo), 1 / np.sqrt(Min))
            self.params[sc + '/bias'] = torch.zeros(Mo, dtype=self.dtype)
            Min = Mo
        if len(self.M) != 0:
            self._weights('logits', (Min, self.M[-1]), 1 / np.sqrt(Min))
        # decoder (models.py:1289-1344)
        if self.dense and (np.asarray(self.p) > 1).any():
            for i in range(1, 1 + n):
                j = n - i
                if self.p[-i] > 1:
                    self._cheb('upconv{}'.format(j), Fin, self.F[-i], self.K[-i])
                    self._bias('upconv{}'.format(j), self.F[-i])
                    Fin = self.F[-i] + self.skipF[-i]
                self._cheb('deconv{}'.format(j), Fin, self.F[-i], self.K[-i])
                if self.batch_norm[-i]:
                    self._bn('deconv{}'.format(j), self.F[-i])
                self._bias('deconv{}'.format(j), self.F[-i])
                Fin = self.F[-i]
            self._cheb('outconv', Fin, self.Fseg, 1)

    def n_params(self):
        return sum(int(np.prod(v.shape)) for v in self.params.values())

    # ---- forward ------------------------------------------------------------------------
    def _bn_apply(self, x, scope, training, new_state):
        mm = scope + '/batch_normalization/moving_mean'
        mv = scope + '/batch_normalization/moving_variance'
        y, nm, nv = ops.batch_normalization(x, self.state[mm], self.state[mv], training)
        new_state[mm], new_state[mv] = nm, nv
        return y

    def forward(self, x, P, training, new_state=None, drop_masks=None, filt_masks=None):
        """`_inference` (+ `_decoder` for dense nets with pooling).  Returns (logits, descriptor)."""
        new_state = {} if new_state is None else new_state
        n = len(self.p)
        pool_layers = [x]                                                    # models.py:1221-1222
        for i in range(n):
            sc = 'conv{}'.format(i + 1)
            fm = filt_masks.get(sc) if (filt_masks and training) else None   # filter dropout: explicit masks (RNG is not parity)
            ops.STORE = True if (i + 1) in self.bf16_layers else None
            try:
                x = self.conv(x, self.Ls[i], P[sc + '/weights'], self.K[i], fm)  # 1229
            finally:
                ops.STORE = None
            if i == n - 1 and len(self.M) == 0:                              # 1232-1234
                break
            if self.batch_norm[i]:
                x = self._bn_apply(x, sc, training, new_state)              # 1235-1236
            x = ops.bias(x, P[sc + '/bias'])                                 # 1238
            x = self.act(x)                                                  # 1239
            x = ops.pool(x, self.p[i], self.pool_kind, self.sampling, self.ratio)        # 1243
            if (i + 1) in self.bf16_layers:
                x = ops._RoundBf16.apply(x)                                  # the product stores this layer's output as bfloat16
            pool_layers.append(x)                                            # 1245
        descriptor = x
        if self.stat == 'histogram':
            x = ops.learned_histogram(x, P['stat/centers'], P['stat/widths'])  # 1263-1266
        else:
            x = ops.statistics(x, self.stat, bool(self.M))                   # 1249-1270
        for i, _ in enumerate(self.M[:-1]):                                  # 1275-1280
            sc = 'fc{}'.format(i + 1)
            x = self.act(ops.fc(x, P[sc + '/weights'], P[sc + '/bias']))
            if training and self.dropout < 1:
                mask = drop_masks[i]                                         # explicit mask: RNG is not part of parity
                x = x * mask / self.dropout
        if len(self.M) != 0:
            x = ops.fc(x, P['logits/weights'])                               # 1283-1285
        if self.dense and (np.asarray(self.p) > 1).any():                    # 687-688
            x = self._decoder(x, P, pool_layers, training, new_state)
        return x, descriptor

    def _decoder(self, x, P, pool_layers, training, new_state):
        n = len(self.p)
        for i in range(1, 1 + n):
            j = n - i
            if self.p[-i] > 1:
                if self.sampling == 'icosahedron':                            # 1297-1301
                    p = self.p[-i - 1] if i + 1 <= n else self.finest_vertices
                else:
                    p = self.p[-i]
                x = ops.unpool(x, p, self.pool_kind, self.sampling, self.ratio)           # 1304
                sc = 'upconv{}'.format(j)
                x = self.conv(x, self.Ls[-i], P[sc + '/weights'], self.K[-i])  # 1307
                x = ops.bias(x, P[sc + '/bias'])                              # 1308
                x = torch.cat([x, pool_layers[-i]], dim=-1)                   # 1311
            sc = 'deconv{}'.format(j)
            x = self.conv(x, self.Ls[-i], P[sc + '/weights'], self.K[-i])      # 1315 / 1327
            if self.batch_norm[-i]:
                x = self._bn_apply(x, sc, training, new_state)
            x = ops.bias(x, P[sc + '/bias'])
            x = self.act(x)
        return self.conv(x, self.Ls[-n], P['outconv/weights'], 1)             # 1340-1341

    # ---- loss / gradients / step ----------------------------------------------------------
    def loss(self, logits, labels, P, data=None):
        wstd = [(P[k], self.std[k]) for k in self.params if k in self.std]
        return ops.loss(logits, labels, self.regression, wstd, self.regularization, self.weighted,
                        data if self.masked else None)

    def loss_and_grads(self, x, labels, training=True, drop_masks=None, filt_masks=None):
        """Forward + loss + autograd gradients w.r.t. every trainable variable."""
        P = {k: v.clone().requires_grad_(True) for k, v in self.params.items()}
        new_state = {}
        logits, _ = self.forward(x, P, training, new_state, drop_masks, filt_masks)
        total, data_loss = self.loss(logits, labels, P, x)
        grads = torch.autograd.grad(total, list(P.values()), allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(P, grads)}
        return total.detach(), logits.detach(), grads, new_state

    def train_step(self, x, labels, optimizer, lr, drop_masks=None):
        """One `sess.run([op_train, op_loss])` (models.py:535, 822-843): gradients, optimizer
        apply, BN moving-average update."""
        total, logits, grads, new_state = self.loss_and_grads(x, labels, True, drop_masks)
        self.params = optimizer.step(dict(self.params), grads, lr)
        self.state.update({k: v.detach() for k, v in new_state.items()})
        return float(total), logits

    def predict_logits(self, x):
        with torch.no_grad():
            logits, desc = self.forward(x, self.params, False)
        return logits, desc
