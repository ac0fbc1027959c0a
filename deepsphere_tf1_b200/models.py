"""`cgcnn` / `deepsphere` models with the reference's constructor and fit / predict / evaluate
API (reference deepsphere/models.py:929-1035 ctor, 432-618 fit, 113-260 predict / probs,
343-430 evaluate, 1827-1839 deepsphere front-end), driving the libdsb200 CUDA kernels.

The reference builds a static TF1 graph once (`build_graph`, models.py:647-726) and then issues
one `sess.run` per batch.  Here the constructor builds a static *layer program* (a list of
ChebConv / head / decoder stages with pre-bound device graphs and parameter views); a training
step runs the program forward, walks it backwards with hand-written adjoint kernels, and applies
the optimizer -- all as libdsb200 launches on one CUDA stream, optionally replayed as one CUDA
graph.  PyTorch supplies device memory, streams and (for data parallelism) NCCL; no torch
operator computes anything on the hot path and there is no CPU fallback.

Differences from the reference that a user sees (also listed in DESIGN.md):
  * `scheduler(step) -> float` and `optimizer(lr) -> deepsphere_tf1_b200.optimizers.*Optimizer`
    replace the lambdas that returned TF objects;
  * `tf_dataset` may be any Python iterator of `(data, labels)` batches;
  * checkpoints are `torch.save` files `checkpoints/<dir_name>/model-<step>.pt` keyed by the
    reference's variable names (`conv1/weights`, `conv1/batch_normalization/moving_mean`, ...);
  * `extra_loss` (triplet loss): NotImplementedError -- upstream passes the rank-3 descriptor to a rank-2 loss (models.py:768-769,
    1247), it cannot run as committed; equiangular grids: any rows x columns the pooling windows divide (`ratio`), on the
    in-repo 8-neighbour graph;
    `conv='monomials'`, `weighted=True`, `statistics='histogram'` and `dropFilt < 1` are supported;
  * `dtype='bfloat16'` (upstream's `dtype` only casts the Laplacians): the ChebConv layers keep their activations, Chebyshev
    planes and activation gradients as bfloat16 (fp32 arithmetic, fp32 weights / statistics); everything else stays float32;
  * TensorFlow checkpoints of the reference can be read / written directly (`load_tf_checkpoint` / `save_tf_checkpoint`).
"""
import glob
import json
import os
import shutil
import time

import numpy as np
import torch

from . import halo, ops, utils
from ._consts import ACT, UNPOOL_REPEAT, UNPOOL_ZEROFILL
from .graph import DeviceGraph, ShardedGraph
from .optimizers import _Optimizer

process_time = time.process_time
perf_counter = time.perf_counter

BN_EPS = 1e-5
BN_MOMENTUM = 0.9


def _truncated_normal(rs, shape, std):
    """tf.truncated_normal_initializer: samples beyond 2 sigma are re-drawn."""
    a = rs.randn(*shape)
    bad = np.abs(a) > 2
    while bad.any():
        a[bad] = rs.randn(int(bad.sum()))
        bad = np.abs(a) > 2
    return (a * std).astype(np.float32)


class _ParamStore(object):
    """All trainable variables in ONE flat fp32 device buffer (plus flat gradient, optimizer
    slots and the per-element regularisation coefficient), so that the regulariser, the
    data-parallel all-reduce and the optimizer are one launch each (models.py:806-808, 831-832)."""

    def __init__(self):
        self.specs = []          # (name, shape, offset, size, std or None)
        self.size = 0

    def add(self, name, shape, init, std=None):
        n = int(np.prod(shape))
        off = (self.size + 3) // 4 * 4                     # 16-byte aligned views
        self.specs.append((name, tuple(shape), off, n, std, init))
        self.size = off + n
        return len(self.specs) - 1

    def materialise(self, device, regularization):
        host = np.zeros(self.size, dtype=np.float32)
        coef = np.zeros(self.size, dtype=np.float32)
        n_weights = sum(s[3] for s in self.specs if s[4] is not None)
        for name, shape, off, n, std, init in self.specs:
            host[off:off + n] = init.reshape(-1)
            if std is not None:
                # regularization * (l2_loss(W) / std^2) / n_weights   (models.py:806-808, 866)
                coef[off:off + n] = regularization / (std * std) / max(n_weights, 1)
        self.w = torch.from_numpy(host).to(device)
        self.coef = torch.from_numpy(coef).to(device)
        self.g = torch.zeros_like(self.w)
        self.views = {}
        self.gviews = {}
        for name, shape, off, n, std, init in self.specs:
            shape2 = shape if len(shape) == 2 else (n,)     # matrices keep their shape, biases are flat
            self.views[name] = self.w[off:off + n].view(shape2)
            self.gviews[name] = self.g[off:off + n].view(shape2)
        self.shapes = {s[0]: s[1] for s in self.specs}
        self.n_trainable = sum(s[3] for s in self.specs)
        self.has_reg = bool(regularization) and n_weights > 0


CLASS_WEIGHTS = (0.34130685, 318.47388343, 14.93759951)     # `weighted=True` (models.py:796)


class _ChebStage(object):
    """filter -> [batch norm] -> [bias] -> [activation] -> [pool]   (models.py:1229-1243;
    decoder up-conv / de-conv / out-conv, models.py:1304-1341, with the unused parts off)."""

    def __init__(self, model, scope, graph, Fin, Fout, K, bn, bias, act, p=1, pool='max', keep=None, first=False):
        self.m, self.scope = model, scope
        self.sg = graph if isinstance(graph, ShardedGraph) else None      # vertex-sharded mode
        self.graph = graph.graph if self.sg is not None else graph
        self.Fin, self.Fout, self.K = Fin, Fout, K
        self.bn, self.has_bias, self.act = bn, bias, act
        self.p, self.pool = p, pool            # HEALPix pooling over p consecutive vertices
        self.keep = keep                       # icosahedral "pooling": keep the first `keep` vertices
        self.mono = getattr(model, 'conv', 'chebyshev5') == 'monomials'       # basis x, Lx, L^2 x, ... (models.py:1085-1120)
        if self.mono and self.sg is not None:
            raise NotImplementedError("conv='monomials' is not available in the vertex-sharded mode")
        # chebyshev5: Xavier-like std (models.py:1080-1083); monomials: `_weight_variable` default 0.1 (models.py:1108, 862)
        std = 0.1 if self.mono else 1 / np.sqrt(Fin * (K + 0.5) / 2)
        model._P.add(scope + '/weights', (Fin * K, Fout), _truncated_normal(model._rs, (Fin * K, Fout), std), std)
        if bn:
            model._add_bn(scope, Fout)
        if bias:
            model._P.add(scope + '/bias', (1, 1, Fout), np.zeros(Fout, np.float32))
        self.post = bn or bias or act != 'identity' or p > 1
        # tiny reduction depth (K*Fin <= 8) on the input layer: never materialise y (csrc/smallconv.cu)
        self.drop = float(getattr(model, 'dropFilt', 1.0))                    # filter dropout (models.py:1068-1076)
        if self.drop < 1 and self.sg is not None:
            raise NotImplementedError('dropFilt < 1 is not available in the vertex-sharded mode')
        self.small = bool(first and self.post and model.fused_small and self.sg is None and self.drop >= 1
                          and ops.smallconv_supported(Fin, K, Fout))
        # BF16 mode: this layer reads, keeps and writes bfloat16 activations (the fused input layer and the vertex-sharded
        # layers stay fp32; tensors are converted where the two kinds of layer meet)
        self.bf16 = bool(getattr(model, 'act_dtype', torch.float32) == torch.bfloat16 and not self.small and self.sg is None)
        self.saved = None

    def forward(self, x, training, save):
        m = self.m
        B, M, Fin = x.shape
        Mg = self.sg.Ml if self.sg is not None else self.graph.M
        if M != Mg or Fin != self.Fin:
            raise ValueError('{}: input [{}, {}] does not fit the Laplacian ({} vertices) / weights ({} features)'
                             .format(self.scope, M, Fin, Mg, self.Fin))
        W = m._P.views[self.scope + '/weights']
        x = ops.to_bf16(x) if self.bf16 else ops.to_f32(x)
        if self.small:
            return self._forward_small(x, W, training, save)
        if self.sg is not None and self.K > 1:
            return self._forward_sharded(x, W, training, save)
        basis = ops.cheb_basis(self.graph, x, self.K, self.mono)           # models.py:1049-1061 / 1094-1103
        umask = None
        if training and self.drop < 1:
            # W * (dropout(ones[Fin, Fout], keep) * keep): a {0, 1} mask shared by the K orders of a filter (1068-1076)
            u = m._dropout_uniform(self.scope + '/filt', W.new_empty(self.Fin, 1, self.Fout))
            umask = u.expand(self.Fin, self.K, self.Fout).contiguous().view(self.Fin * self.K, self.Fout)
            W = ops.dropout(W, umask, self.drop)               # u < keep ? W / keep : 0
            ops.scale(W, self.drop)
        sums = None
        if self.bn and training:
            sums = m._zeros(2 * self.Fout)
        # HEALPix max-pool over 4 vertices fused into the contraction's epilogue where a tensor-core kernel takes the shape: the
        # pooling commutes with batch norm (no scale / shift), bias and a non-decreasing activation.  Used when the
        # un-pooled tensor is not needed afterwards (inference: it is never written).  While training, the batch-norm backward
        # needs y anyway and the epilogues of these kernels are what bounds them: measured inside the cosmo step, the pooled
        # epilogue costs the three forward kernels +35 us and saves 18 us of the normalise pass (profiles/r02_trace_step_pool4.txt)
        fuse_pool = (self.post and self.p == 4 and self.pool == 'max' and self.act in ops.POOL4_ACTS
                     and (not save or ops.POOL4_TRAIN) and ops.contract_pool4_supported(x, self.K, self.Fout))
        ypool = None
        if fuse_pool:
            y, ypool = ops.contract_fwd_pool4(x, basis, W, self.K, self.Fout, sums, want_y=save)
        else:
            y = ops.contract_fwd(x, basis, W, self.K, self.Fout, sums)     # models.py:1062-1078
        mean = rstd = None
        count = float(B * M * m.world_size)
        if self.bn:
            mean, rstd = m._bn_buffers(self.scope, self.Fout)
            mm, mv = m._state[self.scope + '/batch_normalization/moving_mean'], \
                m._state[self.scope + '/batch_normalization/moving_variance']
            if training:
                m._allreduce(sums)                                         # SyncBN over the data-parallel group
                ops.bn_finalize(sums, count, self.Fout, mean, rstd, mm, mv, BN_EPS, BN_MOMENTUM)
            else:
                ops.bn_from_moving(mm, mv, mean, rstd, BN_EPS)
        out = y
        if self.post:
            b = m._P.views[self.scope + '/bias'] if self.has_bias else None
            if ypool is not None:
                out = ops.bn_act_pool_fwd(ypool, mean, rstd, b, 1, self.pool, self.act)     # a quarter of the rows
            else:
                out = ops.bn_act_pool_fwd(y, mean, rstd, b, self.p, self.pool, self.act)
        if save:
            # the pooled output feeds pass 1 of the batch-norm backward (it is the next layer's input anyway)
            lean = self.post and self.bn and training and ops.bn_relu_pool_sums_supported(self.Fout, self.p, self.pool, self.act)
            self.saved = (x, basis, y, mean, rstd, count, training, out if lean else None, umask, W if umask is not None else None)
        if self.keep is not None and self.keep != out.shape[1]:
            out = ops.vertex_resize(ops.to_f32(out), self.keep, 0.0)       # x[:, :p, :]  (models.py:1138,1157)
        return out

    def _forward_small(self, x, W, training, save):
        m = self.m
        B, M, Fin = x.shape
        shape = (B, M, Fin)
        # sample-interleaved layout [M, B, Fin]: the sparse products gather B*Fin contiguous floats
        x = ops.interleave(x)
        basis = ops.cheb_basis(self.graph, x.view(1, M, B * Fin), self.K, self.mono)
        count = float(B * M * m.world_size)
        mean = rstd = gram = None
        if self.bn:
            mean, rstd = m._bn_buffers(self.scope, self.Fout)
            mm, mv = m._state[self.scope + '/batch_normalization/moving_mean'], \
                m._state[self.scope + '/batch_normalization/moving_variance']
            if training:
                gram = ops.cheb_gram(x, basis, self.K, shape)
                m._allreduce(gram)
                sums = ops.bn_sums_from_gram(gram, W, Fin, self.K, self.Fout)
                ops.bn_finalize(sums, count, self.Fout, mean, rstd, mm, mv, BN_EPS, BN_MOMENTUM)
            else:
                ops.bn_from_moving(mm, mv, mean, rstd, BN_EPS)
        b = m._P.views[self.scope + '/bias'] if self.has_bias else None
        want_arg = save and ops.smallconv_has_argmax(self.Fout, self.p, self.pool, self.act)
        out, arg = ops.smallconv_fwd(x, basis, W, self.K, self.Fout, mean, rstd, b, self.p, self.pool, self.act, shape,
                                     want_argmax=want_arg)
        if save:
            self.saved = (x, basis, gram, mean, rstd, count, training, shape, out if want_arg else None, arg)
        if self.keep is not None and self.keep != out.shape[1]:
            out = ops.vertex_resize(out, self.keep, 0.0)
        return out

    def _backward_small(self, dout):
        m = self.m
        x, basis, gram, mean, rstd, count, training, shape, out, arg = self.saved
        self.saved = None
        dout = ops.to_f32(dout)
        B, M, Fin = shape
        if self.keep is not None and dout.shape[1] != M // self.p:
            dout = ops.vertex_resize(dout, M // self.p, 0.0)
        W = m._P.views[self.scope + '/weights']
        b = m._P.views[self.scope + '/bias'] if self.has_bias else None
        acc = ops.smallconv_bwd(x, basis, W, self.K, self.Fout, mean, rstd, b, dout, self.p, self.pool, self.act, shape,
                                out=out, argmax=arg)
        sync = self.bn and training
        if sync:
            m._allreduce(acc)
        gW = m._P.gviews[self.scope + '/weights']
        gb = m._P.gviews[self.scope + '/bias'] if self.has_bias else None
        ops.smallconv_bwd_finalize(acc, gram if sync else None, W, mean if sync else None, rstd if sync else None,
                                   count, Fin, self.K, self.Fout, gW, gb)
        if sync and m.world_size > 1:                      # already global: the flat all-reduce sums them back
            ops.scale(gW, 1.0 / m.world_size)
            if gb is not None:
                ops.scale(gb, 1.0 / m.world_size)
        return None

    # ------------------------------------------------------------------ vertex-sharded mode (one sample)
    def _forward_sharded(self, x, W, training, save):
        """The same layer on this rank's vertex range [1, Ml, Fin].  Every plane of the basis carries the H ghost
        rows of the 1-ring halo behind the Ml owned rows ([1, Me, Fin]); one halo exchange precedes every sparse
        product.  The ghost rows are zeroed afterwards so that the contraction can run over all Me rows without
        touching the batch-norm sums or the weight gradient."""
        m, sg, g, K = self.m, self.sg, self.graph, self.K
        B, Ml, Fin = x.shape
        if B != 1:
            raise ValueError('the vertex-sharded mode takes one sample per step')
        E = torch.empty((K, 1, sg.Me, Fin), dtype=torch.float32, device=x.device)
        ops.vertex_resize(x, sg.Me, 0.0, out=E[0])
        halo.exchange(sg, E[0], m.pg, peer=m._peer_halo)
        wide = sg.rings >= K - 1 and sg.rings > 1      # the halo covers the whole recurrence: no further exchange, all local rows
        for k in range(1, K):
            ops.spmm(g, E[k - 1], E[k - 2] if k >= 2 else None, None, out=E[k], alpha=1.0 if k == 1 else 2.0,
                     beta=0.0 if k == 1 else -1.0, rows=None if wide else Ml)
            if k < K - 1 and not wide:
                halo.exchange(sg, E[k], m.pg, peer=m._peer_halo)
        if sg.H:
            for k in range(K):
                ops.fill(E[k, 0, Ml:], 0.0)
        sums = None
        if self.bn and training:
            sums = m._zeros(2 * self.Fout)
        y = ops.contract_fwd(E[0], E[1:], W, K, self.Fout, sums)           # [1, Me, Fout], ghost rows = 0
        mean = rstd = None
        count = float(Ml * m.world_size)
        if self.bn:
            mean, rstd = m._bn_buffers(self.scope, self.Fout)
            mm, mv = m._state[self.scope + '/batch_normalization/moving_mean'], \
                m._state[self.scope + '/batch_normalization/moving_variance']
            if training:
                m._allreduce(sums)
                ops.bn_finalize(sums, count, self.Fout, mean, rstd, mm, mv, BN_EPS, BN_MOMENTUM)
            else:
                ops.bn_from_moving(mm, mv, mean, rstd, BN_EPS)
        yo = y[:, :Ml]                                                     # owned rows: a contiguous prefix (one sample)
        out = yo
        if self.post:
            b = m._P.views[self.scope + '/bias'] if self.has_bias else None
            out = ops.bn_act_pool_fwd(yo, mean, rstd, b, self.p, self.pool, self.act)
        if save:
            lean = self.post and self.bn and training and ops.bn_relu_pool_sums_supported(self.Fout, self.p, self.pool, self.act)
            self.saved = (E, y, mean, rstd, count, training, out if lean else None)
        return out

    def _backward_sharded(self, dout, need_dx):
        m, sg, g, K = self.m, self.sg, self.graph, self.K
        E, y, mean, rstd, count, training, out = self.saved
        self.saved = None
        dout = ops.to_f32(dout)
        Ml, Fin = sg.Ml, self.Fin
        yo = y[:, :Ml]
        if self.post:
            dy = torch.zeros_like(y)                                        # ghost rows stay zero
            dyo = dy[:, :Ml]
            b = m._P.views[self.scope + '/bias'] if self.has_bias else None
            two_pass = self.bn and training
            if two_pass and out is not None:
                sums = ops.bn_relu_pool_sums(out, dout, b, sums=m._zeros(2 * self.Fout))
            else:
                _, sums = ops.bn_act_pool_bwd(yo, mean, rstd, b, dout, self.p, self.pool, self.act, want_gz=not two_pass,
                                              out=dyo, sums=m._zeros(2 * self.Fout))
            if two_pass:
                m._allreduce(sums)
            if self.has_bias:
                gb = m._P.gviews[self.scope + '/bias']
                ops.sums_to_float(sums, gb)
                if two_pass and m.world_size > 1:
                    ops.scale(gb, 1.0 / m.world_size)
            if two_pass:
                ops.bn_act_pool_bwd_dy(yo, mean, rstd, b, dout, sums, count, self.p, self.pool, self.act, out=dyo)
        else:
            dy = ops.vertex_resize(dout, sg.Me, 0.0)
        ops.contract_bwd_weight(E[0], E[1:], dy, m._P.gviews[self.scope + '/weights'], K)
        if not need_dx:
            return None
        W = m._P.views[self.scope + '/weights']
        wide = sg.rings >= K - 1 and sg.rings > 1
        if wide:
            # dy on the whole halo (one exchange), G_k = dy W_k^T on all local rows, then the adjoint recurrence without further
            # exchanges: A_k is valid on the owned vertices and the rings 1 .. k, A_0 on the owned ones
            if not dy.is_contiguous():
                dy = dy.contiguous()
            halo.exchange(sg, dy, m.pg, peer=m._peer_halo)
        G = ops.contract_bwd_data(dy, W, K, Fin)                           # [K, 1, Me, Fin]
        for k in range(K - 2, -1, -1):                                     # adjoint recurrence (csrc/sparse.cu)
            if not wide:
                halo.exchange(sg, G[k + 1], m.pg, peer=m._peer_halo)
            a2 = G[k + 2] if k + 2 <= K - 1 else None
            ops.spmm(g, G[k + 1], G[k], a2, out=G[k], alpha=2.0 if k >= 1 else 1.0, beta=1.0,
                     gamma=-1.0 if a2 is not None else 0.0, transpose=True, rows=None if wide else Ml)
        return G[0][:, :Ml]

    def backward(self, dout, need_dx=True):
        m = self.m
        if self.sg is not None and self.K > 1:
            return self._backward_sharded(dout, need_dx)
        if self.small:
            if need_dx:
                raise RuntimeError('the fused input layer does not produce an input gradient')
            return self._backward_small(dout)
        x, basis, y, mean, rstd, count, training, out, umask, Wmasked = self.saved
        self.saved = None
        B, M, Fin = x.shape
        if self.keep is not None and dout.shape[1] != M // self.p:
            dout = ops.vertex_resize(ops.to_f32(dout), M // self.p, 0.0)   # adjoint of the slice
        dout = ops.to_bf16(dout) if self.bf16 else ops.to_f32(dout)
        dy = dout
        if self.post:
            b = m._P.views[self.scope + '/bias'] if self.has_bias else None
            two_pass = self.bn and training          # batch norm: sums first, then dy directly (gz never stored)
            if two_pass and out is not None:
                dy, sums = None, ops.bn_relu_pool_sums(out, dout, b, sums=m._zeros(2 * self.Fout))   # pooled tensors only
            else:
                dy, sums = ops.bn_act_pool_bwd(y, mean, rstd, b, dout, self.p, self.pool, self.act, want_gz=not two_pass,
                                               sums=m._zeros(2 * self.Fout))
            if two_pass:
                m._allreduce(sums)
            if self.has_bias:
                    # with synchronised batch norm the sums are already global: pre-divide so that the
                # flat gradient all-reduce (a sum over ranks) leaves the bias gradient unchanged
                gb = m._P.gviews[self.scope + '/bias']
                ops.sums_to_float(sums, gb)
                if self.bn and training and m.world_size > 1:
                    ops.scale(gb, 1.0 / m.world_size)            # the flat all-reduce sums it back
            if two_pass:
                dy = ops.bn_act_pool_bwd_dy(y, mean, rstd, b, dout, sums, count, self.p, self.pool, self.act)
        gW = m._P.gviews[self.scope + '/weights']
        if umask is None:
            ops.contract_bwd_weight(x, basis, dy, gW, self.K)
        else:                                                # gradient w.r.t. the masked weights, times the mask
            tmp = torch.zeros_like(gW)
            ops.contract_bwd_weight(x, basis, dy, tmp, self.K)
            tmp = ops.dropout(tmp, umask, self.drop)
            ops.axpy(tmp, gW, self.drop)                     # gW += keep * (u < keep ? tmp / keep : 0)
        if not need_dx:
            return None
        W = Wmasked if umask is not None else m._P.views[self.scope + '/weights']
        G = ops.contract_bwd_data(dy, W, self.K, Fin)
        return ops.cheb_basis_bwd(self.graph, G, self.mono)


class _FcStage(object):
    """cgcnn.fc (+ activation + dropout) (models.py:1205-1217, 1275-1285)."""

    def __init__(self, model, scope, Min, Mout, bias, act, dropout):
        self.m, self.scope, self.Min, self.Mout = model, scope, Min, Mout
        self.has_bias, self.act, self.dropout = bias, act, dropout
        std = 1 / np.sqrt(Min)
        model._P.add(scope + '/weights', (Min, Mout), _truncated_normal(model._rs, (Min, Mout), std), std)
        if bias:
            model._P.add(scope + '/bias', (Mout,), np.zeros(Mout, np.float32))

    def forward(self, x, training, save):
        m = self.m
        W = m._P.views[self.scope + '/weights']
        b = m._P.views[self.scope + '/bias'] if self.has_bias else None
        y = ops.fc_fwd(x, W, b, self.act)
        out, u = y, None
        if training and self.dropout < 1:
            u = m._dropout_uniform(self.scope, y)
            out = ops.dropout(y, u, self.dropout)
        if save:
            self.saved = (x, y, u)
        return out

    def backward(self, dout, need_dx=True):
        m = self.m
        x, y, u = self.saved
        self.saved = None
        if u is not None:
            dout = ops.dropout(dout, u, self.dropout)
        W = m._P.views[self.scope + '/weights']
        db = m._P.gviews[self.scope + '/bias'] if self.has_bias else None
        return ops.fc_bwd(x, W, y, dout, m._P.gviews[self.scope + '/weights'], db, self.act, need_dx)


class _LossHandle(object):
    """Loss of one training step on its way to the host (`cgcnn.loss_handle`)."""

    def __init__(self, host, event, slot):
        self._host, self._event, self._slot = host, event, slot

    def value(self):
        self._event.synchronize()
        return float(self._host[self._slot])


class base_model(object):
    """Run loops shared by all models (reference `base_model`, models.py:68-879)."""

    sampling = 'healpix'
    icosahedron_finest = 10 * 4 ** 5 + 2       # the reference hard-codes the level-5 mesh (models.py:1300)

    # ------------------------------------------------------------------ high-level interface
    def _batches(self, data, labels, cache):
        """Yield (begin, end, device batch, device labels | None, host labels | None) with the
        tail batch zero-padded to `batch_size` like the reference (models.py:136-160)."""
        size = data.N if cache else data.shape[0]
        bs = self.batch_size
        data_iter = data.iter(bs) if cache else None
        for begin in range(0, size, bs):
            end = min(begin + bs, size)
            host_labels = None
            if cache:
                batch_data, host_labels = self._with_batch_axis(*next(data_iter))
                batch_data = self._as_bmf(np.asarray(batch_data))
            else:
                tmp = data[begin:end]
                if not isinstance(tmp, np.ndarray):
                    tmp = tmp.toarray()
                tmp = self._as_bmf(tmp)
                batch_data = np.zeros((bs,) + tmp.shape[1:], dtype=np.float32)
                batch_data[:end - begin] = tmp
                if labels is not None:
                    lshape = (bs, self.L[0].shape[0]) if self.dense else (bs,)
                    host_labels = np.zeros(lshape, dtype=np.float32 if self.regression else np.int32)
                    host_labels[:end - begin] = labels[begin:end]
            yield begin, end, self._to_device(batch_data), self._labels_to_device(host_labels), host_labels

    @staticmethod
    def _as_bmf(a):
        a = np.asarray(a, dtype=np.float32)
        return a[:, :, None] if a.ndim == 2 else a

    def _with_batch_axis(self, data, labels=None):
        """`LabeledDataset.iter(1)` yields un-batched samples ([M] or [M, F], scalar label) like the reference
        (data.py:58-60); restore the batch axis so that batch_size == 1 (the vertex-sharded mode) goes through fit /
        predict(cache=True)."""
        if self.batch_size != 1:
            return data, labels
        if isinstance(data, torch.Tensor):                 # device-resident datasets (data.DeviceDataset)
            M0 = self.L[0].shape[0]
            if data.dim() == 1 or (data.dim() == 2 and data.shape[0] == M0 and M0 != 1):
                data = data.unsqueeze(0)
                if labels is not None:
                    labels = labels.unsqueeze(0)
            return data, labels
        a = np.asarray(data)
        M0 = self.L[0].shape[0]
        if a.ndim == 1 or (a.ndim == 2 and a.shape[0] == M0 and M0 != 1):
            a = a[None]
            if labels is not None:
                labels = np.asarray(labels)[None]
        return a, labels

    def get_descriptor(self, data, sess=None, cache=True):
        """Output of the last convolutional stage for every sample (models.py:77-111)."""
        self._get_session(sess)
        size = data.N if cache else data.shape[0]
        out = None
        for begin, end, x, _, _ in self._batches(data, None, cache):
            _, desc = self._forward(x, training=False, save=False)
            desc = desc.cpu().numpy()
            if out is None:
                out = np.empty((size,) + desc.shape[1:])
            out[begin:end] = desc[:end - begin]
        return out

    def _run_eval(self, data, labels, sess, cache, want):
        self._get_session(sess)
        size = data.N if cache else data.shape[0]
        loss, out, have_labels = 0.0, None, False
        for begin, end, x, dev_labels, host_labels in self._batches(data, labels, cache):
            logits, _ = self._forward(x, training=False, save=False)
            if dev_labels is not None:
                have_labels = True
                loss += self._loss_only(logits, dev_labels)
            if want == 'prediction':
                if self.regression:
                    batch = logits.reshape(logits.shape[0], -1).squeeze(-1) if not self.dense else logits.squeeze(-1)
                    batch = batch.cpu().numpy()
                else:
                    batch = ops.softmax_argmax(logits, False)[1].cpu().numpy()
            else:
                batch = ops.softmax_argmax(logits, True)[0].cpu().numpy()
            if out is None:
                out = np.empty((size,) + batch.shape[1:])
            out[begin:end] = batch[:end - begin]
        if have_labels:
            return out, loss * self.batch_size / size
        return out

    def predict(self, data, labels=None, sess=None, cache=False):
        """Predicted classes (or regressed values); with labels also the loss (models.py:113-181)."""
        return self._run_eval(data, labels, sess, cache, 'prediction')

    def probs(self, data, nb_class, labels=None, sess=None, cache=False):
        """Class probabilities; with labels also the loss (models.py:183-260)."""
        return self._run_eval(data, labels, sess, cache, 'probabilities')

    def evaluate(self, data, labels, sess=None, cache=False):
        """One evaluation pass; returns (string, accuracy, f1, loss, metrics) (models.py:343-430)."""
        import sklearn.metrics
        import sklearn.preprocessing
        t_cpu, t_wall = process_time(), time.time()
        if self.dense and not self.regression:
            probabilities, loss = self.probs(data, 3, labels, sess, cache=cache)
            predictions = np.argmax(probabilities, axis=-1)
        else:
            predictions, loss = self.predict(data, labels, sess, cache=cache)
        if cache:
            labels = data.get_labels()
        if hasattr(self, 'val_mask'):
            predictions = predictions * data[..., -2]
            labels = labels * data[..., -1]
        if self.regression:
            exp_var = sklearn.metrics.explained_variance_score(labels, predictions)
            r2 = sklearn.metrics.r2_score(labels, predictions)
            mae = sklearn.metrics.mean_absolute_error(labels, predictions)
            string = 'explained variance: {:.4f}, r2: {:.4f}, loss (MSE): {:.3e}, loss (MAE): {:.3e}'.format(
                exp_var, r2, loss, mae)
            accuracy, f1 = exp_var, r2
            mre = np.mean(np.abs((labels - predictions) / np.clip(labels, 1, None))) * 100
            metrics = mae, mre
        elif self.dense:
            # the reference reports the per-class average precision and zeroes accuracy / f1
            # for the 3-class climate task (models.py:390-418)
            labels = np.asarray(labels).flatten()
            true = sklearn.preprocessing.label_binarize(labels, classes=[0, 1, 2])
            AP = sklearn.metrics.average_precision_score(true, probabilities.reshape(-1, 3), average=None)
            ncorrects, class_acc, accuracy, f1 = 0, [0, 0, 0], 0, [0, 0, 0]
            string = 'accuracy: {:.2f} ({:d} / {:d}), f1 (TC): {:.2f}, f1 (AR): {:.2f}, loss: {:.2e}'.format(
                accuracy, ncorrects, len(labels), 100 * f1[1], 100 * f1[2], loss)
            metrics = AP, class_acc
        else:
            metrics = None
            labels = np.asarray(labels).flatten()
            predictions = predictions.flatten()
            ncorrects = int(np.sum(predictions == labels))
            accuracy = 100 * sklearn.metrics.accuracy_score(labels, predictions)
            f1 = 100 * sklearn.metrics.f1_score(labels, predictions, average='weighted')
            string = 'accuracy: {:.2f} ({:d} / {:d}), f1 (weighted): {:.2f}, loss: {:.2e}'.format(
                accuracy, ncorrects, len(labels), f1, loss)
        if sess is None:
            string += '\nCPU time: {:.0f}s, wall time: {:.0f}s'.format(process_time() - t_cpu, time.time() - t_wall)
        return string, accuracy, f1, loss, metrics

    def fit(self, train_dataset, val_dataset, use_tf_dataset=False, restore=True, verbose=True, cache=False):
        """Training loop (models.py:432-618).  Returns (accuracies_validation, losses_validation,
        losses_training, t_step, mean time per batch)."""
        t_cpu, t_wall = process_time(), time.time()
        ckpt_dir = self._get_path('checkpoints')
        start_step = 1
        # Data parallel (torchrun): rank 0 alone resets / writes the run directories; the others wait at a barrier before
        # they look at the checkpoint directory, so that every rank takes the same decision.
        restored = restore and self._has_checkpoint()         # own .pt files, or a run directory written by the reference (TF files)
        if not restored and self.rank == 0:
            shutil.rmtree(self._get_path('summaries'), ignore_errors=True)
            shutil.rmtree(ckpt_dir, ignore_errors=True)
            os.makedirs(ckpt_dir, exist_ok=True)
        self._barrier()
        if restored and self._restore_latest():
            start_step = self.global_step + 1
            print('training from last checkpoint')
        else:
            print('training from scratch')
            self._reset_variables()
        summary = writer = None
        if self.rank == 0:
            os.makedirs(self._get_path('summaries'), exist_ok=True)
            summary = open(os.path.join(self._get_path('summaries'), 'scalars.jsonl'), 'a')
            from .summary import FileWriter                  # TensorBoard event file (tf.summary.FileWriter, models.py:463)
            writer = FileWriter(self._get_path('summaries'))

        accuracies_validation, losses_validation, losses_training = [], [], []
        num_steps = int(self.num_epochs * train_dataset.N / self.batch_size)
        if use_tf_dataset:
            if self.tf_dataset is None:
                raise ValueError('use_tf_dataset=True needs the tf_dataset constructor argument')
            train_iter = iter(self.tf_dataset)
        else:
            train_iter = train_dataset.iter(self.batch_size)
        if not cache:
            val_data, val_labels = val_dataset.get_all_data()
            if len(val_data.shape) == 2:
                val_data = np.expand_dims(val_data, axis=2)

        times = []
        loss = float('nan')
        pending = None

        def fetch():
            d, l = next(train_iter)
            if not isinstance(d, (np.ndarray, torch.Tensor)):
                d = d.toarray()
            return self._with_batch_axis(d, l)

        coming = fetch() if num_steps >= start_step else None
        for step in range(start_step, num_steps + 1):
            t_begin = perf_counter()
            batch_data, batch_labels = coming
            coming = fetch() if step < num_steps else None          # handed to train_step: its copy overlaps this step
            evaluate = (step % self.eval_frequency == 0) or (step == num_steps)
            trace = None
            if evaluate and self.profile and self.rank == 0:
                from .summary import StepTrace               # tf.RunOptions(FULL_TRACE) + RunMetadata, models.py:514-519
                trace = StepTrace()
                self._trace_eager = True                    # launch by launch instead of one graph replay: the host side of
                trace.start()                               # every kernel is part of the trace, as in a FULL_TRACE run
            t_begin_load = perf_counter()
            learning_rate, loss_dev = self.train_step(batch_data, batch_labels, next_batch=coming)
            if evaluate:
                loss = float(loss_dev)                      # device -> host read of the step's loss
                pending = None
                if trace is not None:
                    trace.stop()
                    self._trace_eager = False
            elif self.sync_every_step:
                # the loss of every step reaches the host, one step behind: wait for step i - 1 now that step i is enqueued
                handle = self.loss_handle()
                if pending is not None:
                    loss = pending.value()
                pending = handle
            t_end = perf_counter()
            times.append(t_end - t_begin)
            if not evaluate:
                continue
            epoch = step * self.batch_size / train_dataset.N
            if verbose:
                print('step {} / {} (epoch {:.2f} / {}):'.format(step, num_steps, epoch, self.num_epochs))
                print('  learning_rate = {:.2e}, training loss = {:.2e}'.format(learning_rate, loss))
            losses_training.append(loss)
            record = {'step': step, 'learning_rate': learning_rate, 'loss/total': loss}
            if cache:
                string, accuracy, f1, vloss, metrics = self.evaluate(val_dataset, None, self, cache=cache)
            else:
                string, accuracy, f1, vloss, metrics = self.evaluate(val_data, val_labels, self)
            if not self.regression:
                accuracies_validation.append(accuracy)
            losses_validation.append(vloss)
            record.update({'validation/accuracy': float(np.mean(accuracy)), 'validation/loss': float(vloss)})
            if writer is not None:                           # the tags of models.py:575-603, 815-819, 838, 868
                sc = {'loss/total': loss, 'learning_rate': learning_rate, 'validation/loss': float(vloss)}
                if self.regression:
                    sc.update({'validation/exp_variance': float(accuracy), 'validation/r2': float(f1),
                               'validation/mae': float(metrics[0]), 'validation/mre': float(metrics[1])})
                elif self.dense:
                    AP, class_acc = metrics
                    names = ('BG', 'TC', 'AR')
                    sc['validation/accuracy'] = float(accuracy)
                    sc.update({'validation/accuracy_' + n: float(class_acc[i]) for i, n in enumerate(names)})
                    sc.update({'validation/f1_' + n: float(f1[i]) for i, n in enumerate(names)})
                    sc.update({'validation/AP_' + n: float(AP[i]) for i, n in enumerate(names)})
                else:
                    sc.update({'validation/accuracy': float(accuracy), 'validation/f1': float(f1)})
                hist = {}
                for name in self._P.views:                   # tf.summary.histogram(var.op.name, var) and its gradient
                    hist[name] = self._P.views[name].detach().cpu().numpy()
                    hist[name + '/gradients'] = self._P.gviews[name].detach().cpu().numpy()
                if trace is not None:                        # writer.add_run_metadata(run_metadata, 'step%d'), models.py:604-605
                    sc['profile/step_kernel_us'] = trace.write(self._get_path('summaries'), step)
                writer.add_summary(step, sc, hist)
                writer.flush()
            if verbose:
                print('  validation {}'.format(string))
                print('  CPU time: {:.0f}s, wall time: {:.0f}s, perf_time_load: {:.3f}s, perf_time: {:.3f}s'.format(
                    process_time() - t_cpu, time.time() - t_wall, t_end - t_begin_load, times[-1]))
            if summary is not None:
                summary.write(json.dumps(record) + '\n')
                summary.flush()
            self.save_checkpoint(step)
            self._barrier()                                  # nobody reads a checkpoint rank 0 is still writing

        if verbose and not self.regression and accuracies_validation:
            print('validation accuracy: best = {:.2f}, mean = {:.2f}'.format(
                max(accuracies_validation), np.mean(accuracies_validation[-10:])))
        if summary is not None:
            summary.close()
            writer.close()
        if verbose and times:
            print('time per batch: mean = {:.5f}, var = {:.5f}'.format(np.mean(times), np.var(times)))
        t_step = (time.time() - t_wall) / max(num_steps, 1)
        return accuracies_validation, losses_validation, losses_training, t_step, (np.mean(times) if times else 0.0)

    def _barrier(self):
        if self.world_size > 1:
            import torch.distributed as dist
            dist.barrier(group=self.pg)

    def get_var(self, name):
        """Value of a variable by its reference name, from the latest checkpoint if one exists
        (models.py:620-625)."""
        self._get_session(None)
        if name in self._P.views:
            return self._P.views[name].cpu().numpy().reshape(self._P.shapes[name])
        return self._state[name].cpu().numpy()

    def get_nbr_var(self):
        total = 0
        for name, shape, off, n, std, init in self._P.specs:
            print(name + ':0', n)
            total += n
        return total

    # ------------------------------------------------------------------ checkpoints
    def _get_path(self, folder):
        path = os.path.dirname(os.path.realpath(__file__))
        return os.path.join(path, '..', folder, self.dir_name)

    def state_dict(self):
        sd = {name: self._P.views[name].cpu().reshape(self._P.shapes[name]).clone() for name in self._P.views}
        sd.update({k: v.cpu().clone() for k, v in self._state.items()})
        sd['global_step'] = self.global_step
        for i, s in enumerate(self._slots):
            sd['optimizer/slot{}'.format(i)] = s.cpu().clone()
        return sd

    def load_state_dict(self, sd):
        for name, view in self._P.views.items():
            view.copy_(torch.as_tensor(sd[name], dtype=torch.float32).reshape(view.shape))
        for k in self._state:
            self._state[k].copy_(torch.as_tensor(sd[k], dtype=torch.float32))
        self.global_step = int(sd.get('global_step', 0))
        for i, s in enumerate(self._slots):
            key = 'optimizer/slot{}'.format(i)
            if key in sd:
                s.copy_(sd[key])

    # TensorFlow checkpoints (tf.train.Saver files of the reference, models.py:723, 847-860): same variable names
    def tf_variables(self):
        """{TF variable name: array} as `tf.train.Saver` would store them: trainable variables, batch-norm moving statistics,
        `training/global_step`, and the Adam / Momentum / RMSProp slots under TF's slot names."""
        out = {name: self._P.views[name].detach().cpu().numpy().reshape(self._P.shapes[name]) for name in self._P.views}
        out.update({k: v.detach().cpu().numpy() for k, v in self._state.items()})
        out['training/global_step'] = np.asarray(self.global_step, dtype=np.int64)
        for s, slot in enumerate(self._slots):
            suffix = self._opt.tf_slot_names[s] if hasattr(self._opt, 'tf_slot_names') else 'slot{}'.format(s)
            for name, shape, off, n, std, init in self._P.specs:
                out['{}/{}'.format(name, suffix)] = slot[off:off + n].detach().cpu().numpy().reshape(shape)
        return out

    def save_tf_checkpoint(self, prefix):
        from . import tf_checkpoint
        tf_checkpoint.write_checkpoint(prefix, self.tf_variables())

    def load_tf_checkpoint(self, prefix, strict=True):
        """Restore from a TensorFlow V2 checkpoint written by the reference (`checkpoints/<dir_name>/model-<step>`) or by
        `save_tf_checkpoint`.  Returns the names found in the file that this model does not use."""
        from . import tf_checkpoint
        ck = tf_checkpoint.read_checkpoint(prefix)
        used = set()
        for name, view in self._P.views.items():
            if name not in ck:
                if strict:
                    raise KeyError('checkpoint {} has no variable {}'.format(prefix, name))
                continue
            if int(np.prod(ck[name].shape)) != view.numel():
                raise ValueError('{}: checkpoint shape {} does not match {}'.format(name, ck[name].shape, tuple(view.shape)))
            view.copy_(torch.from_numpy(np.ascontiguousarray(ck[name], dtype=np.float32)).reshape(view.shape))
            used.add(name)
        for k in self._state:
            if k in ck:
                self._state[k].copy_(torch.from_numpy(np.ascontiguousarray(ck[k], dtype=np.float32)))
                used.add(k)
            elif strict:
                raise KeyError('checkpoint {} has no variable {}'.format(prefix, k))
        for key in ('training/global_step', 'global_step'):
            if key in ck:
                self.global_step = int(ck[key])
                used.add(key)
        for s, slot in enumerate(self._slots):
            suffix = self._opt.tf_slot_names[s] if hasattr(self._opt, 'tf_slot_names') else 'slot{}'.format(s)
            for name, shape, off, n, std, init in self._P.specs:
                key = '{}/{}'.format(name, suffix)
                if key in ck:
                    slot[off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(ck[key], dtype=np.float32)).reshape(-1))
                    used.add(key)
        self._graph_exec = None
        return sorted(set(ck) - used)

    def save_checkpoint(self, step):
        if self.rank != 0:
            return
        d = self._get_path('checkpoints')
        os.makedirs(d, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(d, 'model-{}.pt'.format(step)))
        kept = sorted(glob.glob(os.path.join(d, 'model-*.pt')), key=lambda f: int(f.split('-')[-1][:-3]))
        for f in kept[:-5]:                                  # tf.train.Saver(max_to_keep=5)
            os.remove(f)

    def _latest_checkpoint(self):
        files = glob.glob(os.path.join(self._get_path('checkpoints'), 'model-*.pt'))
        if not files:
            return None
        return max(files, key=lambda f: int(f.split('-')[-1][:-3]))

    def _has_checkpoint(self):
        if self._latest_checkpoint() is not None:
            return True
        from . import tf_checkpoint
        return tf_checkpoint.latest_checkpoint(self._get_path('checkpoints')) is not None

    def _restore_latest(self):
        f = self._latest_checkpoint()
        if f is None:
            # a run directory written by the reference itself: TensorFlow checkpoint files
            from . import tf_checkpoint
            prefix = tf_checkpoint.latest_checkpoint(self._get_path('checkpoints'))
            if prefix is None:
                return False
            self.load_tf_checkpoint(prefix, strict=False)
            return True
        self.load_state_dict(torch.load(f, map_location='cpu'))
        return True

    def _get_session(self, sess=None):
        """The reference restores the latest checkpoint when no session is passed
        (models.py:851-860); a live model (`sess` = anything not None) keeps its weights."""
        if sess is None:
            self._restore_latest()
        return self


class cgcnn(base_model):
    """Graph CNN with Chebyshev filters (reference `cgcnn`, models.py:882-1419).

    Same hyper-parameters as the reference: L, F, K, p, batch_norm (one entry per graph
    convolution), M (fully connected sizes, last = classes), conv / pool / activation /
    statistics choices, Fseg, regularization, dropout, batch_size, eval_frequency, regression,
    dense, dir_name.  Extra keyword arguments: `device` (default current CUDA device), `seed`,
    `process_group` (torch.distributed group for data parallelism; default: the world group if
    initialised), `cuda_graph` (replay the training step as one CUDA graph), `sync_every_step` (fit reads the loss of every step,
    one step behind, instead of only at the evaluation steps).  `profile=True` traces the training step of every evaluation as
    upstream (models.py:514-519, 604-605): `summaries/<dir_name>/trace-step<N>.json` (Chrome trace) and `kernels-step<N>.txt`
    (summary.StepTrace); `debug` (the TF debugger session, models.py:471-472) has no counterpart and is ignored.
    """

    def __init__(self, L, F, K, p, batch_norm, M,
                 num_epochs, scheduler, optimizer, num_feat_in=1, tf_dataset=None, restore=False,
                 conv='chebyshev5', pool='max', activation='relu', statistics=None, Fseg=None, dtype=np.float32,
                 regularization=0, dropout=1, batch_size=128, eval_frequency=200, regression=False, dense=False,
                 weighted=False, mask=None, extra_loss=False, dropFilt=1, dir_name='', profile=False, debug=False,
                 device=None, seed=None, process_group=None, cuda_graph=False, sync_every_step=True, verbose=True,
                 fused_small=True, vertex_shard=False):
        # Verify the consistency w.r.t. the number of layers (models.py:937-950).
        if not len(L) == len(F) == len(K) == len(p) == len(batch_norm):
            raise ValueError('Wrong specification of the convolutional layers: '
                             'parameters L, F, K, p, batch_norm, must have the same length.')
        if not np.all(np.array(p) >= 1):
            raise ValueError('Down-sampling factors p should be greater or equal to one.')
        p_log2 = np.where(np.array(p) > 1, np.log2(p), 0)
        if not np.all(np.mod(p_log2, 1) == 0) and self.sampling != 'icosahedron':
            raise ValueError('Down-sampling factors p should be powers of two.')
        if len(M) == 0 and p[-1] != 1 and self.sampling != 'icosahedron' and not dense:
            raise ValueError('Down-sampling should not be used in the last '
                             'layer if no fully connected layer follows.')
        if mask and not isinstance(mask, list):
            raise ValueError('Must provide a list of mask for training and validation.')
        if conv not in ('chebyshev5', 'monomials'):                        # getattr(self, conv), models.py:1017
            raise NotImplementedError("conv='{}' is outside the ChebConv hot path".format(conv))
        self.conv = conv
        if pool not in ('max', 'average'):
            raise AttributeError("'cgcnn' object has no attribute 'pool_{}'".format(pool))
        if activation not in ACT:
            raise AttributeError("unknown activation '{}'".format(activation))
        if statistics not in (None, 'mean', 'var', 'meanvar', 'max', 'histogram'):
            raise ValueError('Unknown statistical layer {}'.format(statistics))
        if extra_loss:
            raise NotImplementedError('extra_loss (triplet loss) is outside the hot path')
        self.dropFilt = float(dropFilt)
        self.weighted = bool(weighted)
        if self.sampling == 'equiangular':
            # 2-D pooling of the (colatitude, longitude) grid with a sqrt(p) x sqrt(p) window (models.py:1131-1136, 1150-1155).
            # Here the vertices of every level are kept in Z (Morton) order of their grid position, so that a window is p
            # CONSECUTIVE vertices and every kernel of the HEALPix path (pooling, fused epilogues, tile plans) applies unchanged;
            # inputs, dense outputs and descriptors are permuted at the boundary (`_equi_gather`).
            # `ratio` = columns / rows of the grid (an attribute, as upstream: deepsphere.__init__ sets it, models.py:1832-1836).
            # Every level that is pooled by p needs sides divisible by sqrt(p) at that point of the chain.
            rows = int(round((L[0].shape[0] / float(getattr(self, 'ratio', 1.))) ** 0.5))
            cols = L[0].shape[0] // max(rows, 1)
            for pi in (p[:-1] if len(M) == 0 else p):          # the last layer of a model without fc head does not pool
                s = int(round(pi ** 0.5))
                if s * s != pi or rows % s or cols % s:
                    raise ValueError('equiangular pooling by {} needs a square window that divides the {} x {} grid'
                                     .format(pi, rows, cols))
                rows, cols = rows // s, cols // s
            if vertex_shard:
                raise NotImplementedError('vertex_shard=True is a HEALPix mode')
            if dense and pool == 'max' and (np.asarray(p) > 1).any():
                raise NotImplementedError('unpooling max with equiangular sampling is not yet implemented')   # models.py:1367
        # dtype: upstream it only casts the Laplacians (models.py:1047).  Here 'bfloat16' selects the BF16 mode of the ChebConv
        # layers (activations stored as bfloat16, fp32 arithmetic and accumulation; include/dsb200.h, "BF16 mode"); the
        # Laplacians, the weights and the statistics stay float32 either way.
        if (isinstance(dtype, str) and dtype in ('bfloat16', 'bf16')) or dtype is torch.bfloat16:
            self.act_dtype = torch.bfloat16
        elif np.dtype(dtype) == np.float32:
            self.act_dtype = torch.float32
        else:
            raise NotImplementedError("dtype must be float32 or 'bfloat16'")
        if not torch.cuda.is_available():
            raise RuntimeError('deepsphere_tf1_b200 needs a CUDA device (sm_100a); there is no CPU fallback')

        self.L, self.F, self.K, self.p, self.M = L, F, K, p, M
        self.Fseg = Fseg
        self.num_epochs = num_epochs
        self.scheduler, self.optimizer = scheduler, optimizer
        self.regularization, self.dropout = regularization, dropout
        self.batch_size, self.eval_frequency = batch_size, eval_frequency
        self.batch_norm = batch_norm
        self.dir_name = dir_name
        self.pool_kind, self.activation_name, self.statistics = pool, activation, statistics
        self.profile, self.debug = profile, debug
        self.regression, self.dense = regression, dense
        self.restore, self.tf_dataset = restore, tf_dataset
        self.num_feat_in = num_feat_in
        self.cuda_graph, self.sync_every_step = cuda_graph, sync_every_step
        self._trace_eager = False
        self.fused_small = fused_small
        self.vertex_shard = bool(vertex_shard)
        if self.vertex_shard:
            # contiguous nested-index vertex ranges, 1-ring halo exchange per sparse product
            if batch_size != 1 or len(M) or dense or statistics != 'mean' or self.sampling != 'healpix':
                raise NotImplementedError('vertex_shard=True supports one sample per step, HEALPix pooling, '
                                          "statistics='mean', no fully connected layers, no decoder")
        if mask:
            # models.py:1030-1032: only the PRESENCE of the attribute switches the masked loss on (models.py:785-787): the last two
            # input features are the masks of the predictions and of the labels
            self.train_mask, self.val_mask = mask[0], mask[1]
            if not (regression and dense):
                raise ValueError('masks apply to the dense regression loss (models.py:785-787)')

        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        import torch.distributed as dist
        self.pg = process_group
        if dist.is_available() and dist.is_initialized():
            self.world_size = dist.get_world_size(self.pg)
            self.rank = dist.get_rank(self.pg)
        else:
            self.world_size, self.rank = 1, 0

        if verbose:
            self._print_architecture(num_feat_in)
        self._build_program(seed)
        self.global_step = 0
        if restore:
            self._restore_latest()

    # ------------------------------------------------------------------ construction
    def _print_architecture(self, num_feat_in):
        L, F, K, p, M = self.L, self.F, self.K, self.p, self.M
        print('NN architecture')
        print('  input: M_0 = {}'.format(L[0].shape[0]))
        for i in range(len(p)):
            print('  layer {0}: cgconv{0}'.format(i + 1))
            print('    representation: M_{0} * F_{1} / p_{1} = {2} * {3} / {4} = {5}'.format(
                i, i + 1, L[i].shape[0], F[i], p[i], L[i].shape[0] * F[i] // p[i]))
            F_last = F[i - 1] if i > 0 else num_feat_in
            print('    weights: F_{0} * F_{1} * K_{1} = {2} * {3} * {4} = {5}'.format(
                i, i + 1, F_last, F[i], K[i], F_last * F[i] * K[i]))
            if not (i == len(p) - 1 and len(M) == 0):
                print('    biases: F_{} = {}'.format(i + 1, F[i]))
            if self.batch_norm[i]:
                print('    batch normalization')
        if self.statistics is not None:
            print('  Statistical layer: {}'.format(self.statistics))

    def _add_bn(self, scope, F):
        self._state[scope + '/batch_normalization/moving_mean'] = torch.zeros(F, device=self.device)
        self._state[scope + '/batch_normalization/moving_variance'] = torch.ones(F, device=self.device)

    def _bn_buffers(self, scope, F):
        if scope not in self._bn_buf:
            self._bn_buf[scope] = (torch.empty(F, device=self.device), torch.empty(F, device=self.device))
        return self._bn_buf[scope]

    def _equi_perm(self, M):
        """(perm, inv) of an equiangular grid of M vertices, [sqrt(M / ratio), sqrt(M * ratio)] as upstream reshapes it
        (models.py:1134): perm[k] = row-major index of the k-th vertex in the blocked Z order of
        utils.equiangular_block_order (plain Morton order on a square power-of-two grid); inv[perm[k]] = k.  Cached per size,
        with int32 device copies."""
        cache = self.__dict__.setdefault('_equi_perms', {})
        if M not in cache:
            ratio = float(getattr(self, 'ratio', 1.))
            rows, cols = int(round((M / ratio) ** 0.5)), int(round((M * ratio) ** 0.5))
            if rows * cols != M:
                raise ValueError('{} vertices are not a grid with {} times as many columns as rows'.format(M, ratio))
            perm = utils.equiangular_block_order(rows, cols)
            inv = np.empty(M, np.int64)
            inv[perm] = np.arange(M, dtype=np.int64)
            cache[M] = (perm, inv, {})
        return cache[M]

    def _equi_gather(self, x, inverse=False):
        """x [B, M, F] -> out[b, k] = x[b, perm[k]] (row-major -> Z order) or, inverse, out[b, v] = x[b, inv[v]] (back)."""
        B, M, F = x.shape
        perm, inv, dev = self._equi_perm(M)
        key = (B, bool(inverse))
        if key not in dev:
            src = inv if inverse else perm
            idx = (np.arange(B, dtype=np.int64)[:, None] * M + src[None, :]).reshape(-1)
            dev[key] = torch.from_numpy(idx.astype(np.int32)).to(self.device)
        return ops.gather_rows(x.reshape(B * M, F), dev[key]).view(B, M, F)

    def _graph_for(self, i):
        """Device copy of Laplacian level i; identical consecutive levels share one copy."""
        if i in self._graphs:
            return self._graphs[i]
        for j, g in list(self._graphs.items()):
            Lj = self.L[j]
            if Lj.shape == self.L[i].shape and Lj.nnz == self.L[i].nnz and (Lj != self.L[i]).nnz == 0:
                self._graphs[i] = g
                return g
        pre = getattr(self, '_prebuilt_graphs', None)
        if self.vertex_shard:
            # (K - 1)-ring halos: one exchange per layer and direction instead of one per sparse product (graph.ShardedGraph);
            # DSB200_SHARD_RINGS=1 restores the 1-ring halo with an exchange before every product
            rings = int(os.environ.get('DSB200_SHARD_RINGS', 0)) or max(1, max(self.K) - 1)
            self._graphs[i] = ShardedGraph(self.L[i], self.rank, self.world_size, self.device, rings=rings)
        elif pre is not None and pre[i] is not None and pre[i].device == self.device:
            self._graphs[i] = pre[i]                          # built on the device by graph_build.py
        else:
            # HEALPix geometry of the level (nside, nested pixel numbers) when `deepsphere` built the graphs: lets the
            # fused tile recurrence (csrc/cheb_tile.cu) address the vertices as pixel squares
            nested = getattr(self, '_nested', None)
            Li = self.L[i]
            if self.sampling == 'equiangular':               # vertices in Z order (see __init__)
                perm = self._equi_perm(Li.shape[0])[0]
                Li = Li.tocsr()[perm][:, perm].tocoo()
            self._graphs[i] = DeviceGraph(Li, self.device, nested=nested[i] if nested else None)
        return self._graphs[i]

    def _build_program(self, seed):
        self._rs = np.random.RandomState(seed)
        self._P = _ParamStore()
        self._state, self._bn_buf, self._graphs = {}, {}, {}
        n = len(self.p)
        icos = self.sampling == 'icosahedron'
        act, pool = self.activation_name, self.pool_kind
        self._encoder = []
        Fin, Mv = self.num_feat_in, self.L[0].shape[0]
        skipF = [Fin]
        for i in range(n):
            scope = 'conv{}'.format(i + 1)
            g = self._graph_for(i)
            Mv = g.Ml if self.vertex_shard else g.M
            if i == n - 1 and len(self.M) == 0:            # linear layer before the softmax (models.py:1232-1234)
                self._encoder.append(_ChebStage(self, scope, g, Fin, self.F[i], self.K[i], False, False, 'identity'))
                Fin = self.F[i]
                break
            pi = self.p[i]
            if icos:
                st = _ChebStage(self, scope, g, Fin, self.F[i], self.K[i], self.batch_norm[i], True, act,
                                keep=(pi if pi > 1 else None), first=(i == 0))
                if pi > 1:
                    Mv = pi
            else:
                if Mv % pi:
                    raise ValueError('layer {}: {} vertices cannot be pooled by {}'.format(i + 1, Mv, pi))
                st = _ChebStage(self, scope, g, Fin, self.F[i], self.K[i], self.batch_norm[i], True, act, p=pi, pool=pool,
                                first=(i == 0))
                Mv //= pi
            self._encoder.append(st)
            Fin = self.F[i]
            skipF.append(Fin)
        # statistics + fully connected head (models.py:1249-1285)
        self._fc = []
        if self.statistics == 'histogram':                 # learned histogram: trainable centers / widths (models.py:1173-1185)
            if self.vertex_shard:
                raise NotImplementedError("statistics='histogram' is not available in the vertex-sharded mode")
            nb, rg = ops.HIST_BINS, ops.HIST_RANGE
            self._P.add('stat/centers', (1, 1, Fin, nb), np.tile(np.linspace(-rg, rg, nb, dtype=np.float32), (Fin, 1)))
            self._P.add('stat/widths', (1, 1, Fin, nb), np.full((Fin, nb), 4 * rg / nb, np.float32))
        if self.M:
            Min = {None: Mv * Fin, 'meanvar': 2 * Fin, 'histogram': ops.HIST_BINS * Fin}.get(self.statistics, Fin)
            for i, Mo in enumerate(self.M[:-1]):
                self._fc.append(_FcStage(self, 'fc{}'.format(i + 1), Min, Mo, True, act, self.dropout))
                Min = Mo
            self._fc.append(_FcStage(self, 'logits', Min, self.M[-1], False, 'identity', 1))
        # decoder (models.py:1289-1344)
        self._decoder = []
        self._has_decoder = bool(self.dense and (np.asarray(self.p) > 1).any())
        if self._has_decoder:
            for i in range(1, n + 1):
                j = n - i
                g = self._graph_for(j)
                up = None
                if self.p[-i] > 1:
                    if icos:
                        target = self.p[-i - 1] if i + 1 <= n else self.icosahedron_finest  # models.py:1297-1301
                        unpool = ('pad', target)
                    else:
                        if pool == 'max':
                            unpool = ('zerofill', self.p[-i])
                        else:
                            unpool = ('repeat', self.p[-i])
                    up = (unpool, _ChebStage(self, 'upconv{}'.format(j), g, Fin, self.F[-i], self.K[-i],
                                             False, True, 'identity'))
                    Fin = self.F[-i] + skipF[-i]
                de = _ChebStage(self, 'deconv{}'.format(j), g, Fin, self.F[-i], self.K[-i],
                                self.batch_norm[-i], True, act)
                Fin = self.F[-i]
                self._decoder.append((up, de, len(skipF) - i))        # pool_layers[-i]  (models.py:1311)
            self._outconv = _ChebStage(self, 'outconv', self._graph_for(0), Fin, self.Fseg, 1, False, False, 'identity')
        self._P.materialise(self.device, self.regularization)
        self._init_w = self._P.w.clone()
        # optimizer state
        self._opt = self.optimizer(self.scheduler(0)) if self.optimizer is not None else None
        if self._opt is not None and not isinstance(self._opt, _Optimizer):
            raise TypeError('optimizer(lr) must return a deepsphere_tf1_b200.optimizers.*Optimizer')
        self._slots = []
        if self._opt is not None:
            for s in range(self._opt.slots):
                init = self._opt.slot1_init if s == 0 else 0.0
                self._slots.append(torch.full_like(self._P.w, init))
        self._hyper = torch.zeros(8, device=self.device)
        self._loss = torch.zeros(1, device=self.device)
        self._dropout_u = {}
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(0 if seed is None else int(seed))
        self._graph_exec = None
        self._prefetched = None
        self._copy_stream = None
        self._peer = None
        self._peer_halo = None
        if self.vertex_shard and self.world_size > 1 and os.environ.get('DSB200_PEER_HALO', '0') == '1':
            # opt-in: one-kernel halo exchange over NVLink peer memory (csrc/peer_halo.cu).  Measured on 2 GPUs inside the step's
            # CUDA graph it is no faster than NCCL's grouped send / recv (2.75 vs 2.48 ms at Nside 512): the ~57 rank
            # synchronisations of a sharded step cost their dependency chain, not the transport
            cap = max([st.sg.H * max(st.Fin, st.Fout) for st in self._encoder if st.sg is not None] + [4])
            try:
                from .peer import PeerHalo
                ph = PeerHalo(cap, self.pg)
                self._peer_halo = ph if ph.enabled else None
            except Exception:
                self._peer_halo = None
        self._arena = torch.zeros(16384, dtype=torch.float64, device=self.device)
        self._arena_off = None

    def _reset_variables(self):
        """`sess.run(op_init)`: back to the initial values (models.py:476-479)."""
        self._P.w.copy_(self._init_w)
        for k, v in self._state.items():
            v.fill_(0.0 if k.endswith('moving_mean') else 1.0)
        for s, buf in enumerate(self._slots):
            buf.fill_(self._opt.slot1_init if s == 0 else 0.0)
        self.global_step = 0
        self._graph_exec = None

    def set_var(self, name, value):
        """Overwrite a variable / batch-norm statistic by its reference name."""
        t = torch.as_tensor(np.asarray(value), dtype=torch.float32).to(self.device)
        dst = self._P.views[name] if name in self._P.views else self._state[name]
        dst.copy_(t.reshape(dst.shape))

    def get_filter_coeffs(self, layer, ind_in=None, ind_out=None):
        """Chebyshev coefficients of a layer as K x Fout x Fin (models.py:1377-1401)."""
        K, Fout = self.K[layer - 1], self.F[layer - 1]
        w = self.get_var('conv{}/weights'.format(layer)).reshape((-1, K, Fout)).transpose([1, 2, 0])
        if ind_in:
            w = w[:, :, ind_in]
        if ind_out:
            w = w[:, ind_out, :]
        return w

    # ------------------------------------------------------------------ device plumbing
    def _to_device(self, a):
        if isinstance(a, torch.Tensor):
            t = a.to(self.device, dtype=torch.float32, non_blocking=True)
        else:
            t = torch.from_numpy(np.ascontiguousarray(self._as_bmf(a))).to(self.device, non_blocking=True)
        if t.dim() == 2:
            t = t.unsqueeze(2)
        if self.vertex_shard:                              # this rank's contiguous vertex range
            Ml = t.shape[1] // self.world_size
            t = t[:, self.rank * Ml:(self.rank + 1) * Ml]
        return t.contiguous()

    def _labels_to_device(self, labels):
        if labels is None:
            return None
        if isinstance(labels, torch.Tensor):
            t = labels.to(self.device, non_blocking=True)
        else:
            t = torch.from_numpy(np.ascontiguousarray(labels)).to(self.device, non_blocking=True)
        t = t.to(torch.float32 if self.regression else torch.int32)
        return t.contiguous()

    def _zeros(self, n):
        """n zeroed doubles for the per-feature sums of this step: slices of one arena that `_step_device` clears
        with a single memset (ten separate fills per step otherwise).  Outside a training step: a fresh tensor."""
        if self._arena_off is None or self._arena_off + n > self._arena.numel():
            return torch.zeros(n, dtype=torch.float64, device=self.device)
        t = self._arena[self._arena_off:self._arena_off + n]
        self._arena_off += (n + 1) // 2 * 2
        return t

    def _allreduce(self, t):
        if self.world_size > 1:
            import torch.distributed as dist
            # small fp64 vectors (SyncBN sums): one single-CTA kernel over NVLink peer memory instead of an NCCL launch
            # (csrc/peer_allreduce.cu; ~10 us instead of ~17-30 us per call, 13 calls per cosmo step).  Default on where CUDA IPC
            # between the ranks works (one node); DSB200_PEER_ALLREDUCE=0 keeps everything on NCCL.
            if self._peer is None and os.environ.get('DSB200_PEER_ALLREDUCE', '1') == '1':
                from .peer import PeerAllReduce
                try:
                    self._peer = PeerAllReduce(self.pg)
                    if not self._peer.enabled:
                        self._peer = False
                except Exception:                         # no IPC / not one node: NCCL only
                    self._peer = False
            elif self._peer is None:
                self._peer = False
            if self._peer and self._peer.allreduce(t):    # small fp64 vectors: one kernel over NVLink peer memory
                return
            dist.all_reduce(t, group=self.pg)

    def _dropout_uniform(self, scope, like):
        u = self._dropout_u.get(scope)
        if u is None or u.shape != like.shape:
            u = torch.empty_like(like)
            self._dropout_u[scope] = u
        u.uniform_(generator=self._gen)
        return u

    # ------------------------------------------------------------------ the layer program
    def _forward(self, x, training, save):
        """`_inference` (+ `_decoder`): returns (logits, descriptor) (models.py:1219-1344)."""
        self._input = x                                    # the masked loss reads its last two channels
        equi = self.sampling == 'equiangular'
        if equi:
            x = self._equi_gather(ops.to_f32(x))
        skips = [x]
        for st in self._encoder:
            x = st.forward(x, training, save)
            if st.post:
                skips.append(x)
        x = ops.to_f32(x)                                  # BF16 mode: everything outside the ChebConv layers is fp32
        descriptor = x
        if equi:
            descriptor = self._equi_gather(x, inverse=True)     # the reference's vertex order
            if not self._has_decoder:
                x = descriptor                                  # ... also for the statistics / the flattened head
        B = x.shape[0]
        self._stat_saved = None
        if self.statistics == 'histogram':
            s = ops.hist_fwd(x, self._P.views['stat/centers'], self._P.views['stat/widths'])
            if save:
                self._stat_saved = (x,)
            x = s
        elif self.statistics is not None:
            s = ops.stat_fwd(x, self.statistics)
            if self.vertex_shard and self.world_size > 1:
                # mean over the whole sphere = mean of the equally sized per-rank means
                self._allreduce(s)
                ops.scale(s, 1.0 / self.world_size)
            if save:
                self._stat_saved = (x,)
            x = s
        elif self.M:
            if save:
                self._stat_saved = (x.shape,)
            x = x.reshape(B, -1)
        for fc in self._fc:
            x = fc.forward(x, training, save)
        if self._has_decoder:
            for up, de, j in self._decoder:
                if up is not None:
                    (kind, arg), stage = up
                    x = ops.to_f32(x)
                    if kind == 'pad':
                        xin = ops.vertex_resize(x, arg, 1.0)                 # tf.pad(..., constant_values=1)
                    else:
                        xin = ops.unpool_fwd(x, arg, UNPOOL_REPEAT if kind == 'repeat' else UNPOOL_ZEROFILL)
                    if xin.shape[1] != stage.graph.M:
                        raise ValueError('Down-sampling should not be used in the first layers if training '
                                         'for segmentation task')
                    stage.prev_M = x.shape[1]
                    x = stage.forward(xin, training, save)
                    x = ops.concat_f(ops.to_f32(x), ops.to_f32(skips[j]))
                x = de.forward(x, training, save)
            x = ops.to_f32(self._outconv.forward(x, training, save))
            if equi:
                x = self._equi_gather(x, inverse=True)
        return x, descriptor

    def _backward(self, dlogits):
        """Adjoint of `_forward`, accumulating into the flat gradient buffer."""
        n_enc = len(self._encoder)
        skip_grads = {}                                    # encoder skip index -> gradient from the decoder
        d = dlogits
        equi = self.sampling == 'equiangular'
        if self._has_decoder:
            if equi:
                d = self._equi_gather(d.contiguous())
            d = self._outconv.backward(d)
            for up, de, j in reversed(self._decoder):
                d = de.backward(d)
                if up is not None:
                    (kind, arg), stage = up
                    Fa = stage.Fout
                    da, db = ops.split_f(ops.to_f32(d), Fa)
                    if j in skip_grads:
                        ops.axpy(db, skip_grads[j], 1.0)
                    else:
                        skip_grads[j] = db
                    dxin = ops.to_f32(stage.backward(da))
                    if kind == 'pad':
                        Mprev = stage.prev_M
                        d = ops.vertex_resize(dxin, Mprev, 0.0) if dxin.shape[1] != Mprev else dxin
                    else:
                        d = ops.unpool_bwd(dxin, arg, UNPOOL_REPEAT if kind == 'repeat' else UNPOOL_ZEROFILL)
        d = ops.to_f32(d)
        for fc in reversed(self._fc):
            d = fc.backward(d)
        if self.statistics == 'histogram':
            P = self._P
            d = ops.hist_bwd(self._stat_saved[0], P.views['stat/centers'], P.views['stat/widths'], d,
                             P.gviews['stat/centers'], P.gviews['stat/widths'])
        elif self.statistics is not None:
            d = ops.stat_bwd(self._stat_saved[0], d, self.statistics)
        elif self.M:
            d = d.reshape(self._stat_saved[0])
        if equi and not self._has_decoder:
            d = self._equi_gather(ops.to_f32(d).contiguous())    # adjoint of the un-permutation in _forward
        # encoder: stage k produced skips[k + 1] (when it has a post chain)
        for k in range(n_enc - 1, -1, -1):
            st = self._encoder[k]
            if st.post and (k + 1) in skip_grads:
                d = ops.to_f32(d)
                ops.axpy(skip_grads.pop(k + 1), d, 1.0)
            d = st.backward(d, need_dx=k > 0)
        return d

    def _loss_fwd_bwd(self, logits, labels, need_grad):
        """Data term of `base_model.loss` (models.py:777-805) into self._loss; returns dlogits."""
        gs = 1.0 / self.world_size
        if self.regression:
            pred = logits.reshape(-1)
            tgt = labels.reshape(-1)
            if pred.numel() != tgt.numel():
                raise ValueError('regression: logits {} do not match labels {}'.format(tuple(logits.shape),
                                                                                       tuple(labels.shape)))
            masked = hasattr(self, 'train_mask')
            if masked:                                     # predictions * data[..., -2], labels * data[..., -1]
                xin = self._input
                pred, tgt = ops.mul_channel(pred, xin, -2), ops.mul_channel(tgt.contiguous(), xin, -1)
            d = ops.mse(pred, tgt, self._loss, gs, need_grad)
            if masked and need_grad:
                d = ops.mul_channel(d, xin, -2)
            return d.reshape(logits.shape) if need_grad else None
        cw = None
        if self.weighted:                                  # models.py:795-802: hard-coded weights of 3 classes
            if logits.shape[-1] != 3:
                raise ValueError('weighted cross-entropy is defined for 3 classes (models.py:793-796)')
            if getattr(self, '_class_w', None) is None:
                self._class_w = torch.tensor(CLASS_WEIGHTS, dtype=torch.float32, device=self.device)
            cw = self._class_w
        return ops.softmax_ce(logits, labels, self._loss, gs, need_grad, cw)

    def _loss_only(self, logits, labels):
        self._loss.zero_()
        self._loss_fwd_bwd(logits, labels, False)
        if self._P.has_reg:
            ops.l2_reg(self._P.w, self._P.coef, None, self._loss)
        return float(self._loss)

    # ------------------------------------------------------------------ one training step
    def _step_device(self, x, labels):
        """forward + loss + backward + all-reduce + regulariser + optimizer, all on the stream.
        The loss of the step (data term averaged over the ranks + regulariser) lands in self._loss."""
        P = self._P
        ops.fill(P.g, 0.0)
        self._loss.zero_()
        self._arena.zero_()
        self._arena_off = 0
        try:
            logits, _ = self._forward(x, training=True, save=True)
            dlogits = self._loss_fwd_bwd(logits, labels, True)
            self._backward(dlogits)
        finally:
            self._arena_off = None
        if self.world_size > 1:
            self._allreduce(P.g)                           # one flat NCCL all-reduce for every gradient
            ops.scale(self._loss, 1.0 / self.world_size)
            self._allreduce(self._loss)
        if P.has_reg:
            ops.l2_reg(P.w, P.coef, P.g, self._loss)
        ops.optimizer_step(self._opt.kind, P.w, P.g, self._slots[0] if self._slots else None,
                           self._slots[1] if len(self._slots) > 1 else None, self._hyper)

    def _advance_hyper(self):
        """Learning rate of this step (`scheduler(global_step)`, models.py:826-827) and the
        optimizer's scalars -> device; increments global_step.  The 32-byte copy comes from
        pageable memory so the host may run ahead of the stream without overwriting it."""
        step = self.global_step
        lr = float(self.scheduler(step))
        self._hyper.copy_(torch.tensor(self._opt.hyper(lr, step + 1), dtype=torch.float32), non_blocking=True)
        self.global_step = step + 1
        return lr

    def train_step(self, batch_data, batch_labels, next_batch=None):
        """One `sess.run([op_train, op_loss])` (models.py:535-536, 822-843).  Returns
        (learning_rate, loss) with the loss still on the device (a 1-element tensor).
        `next_batch` = (data, labels) of the FOLLOWING step, if the caller has it (fit does): its host -> device copy is
        enqueued on a copy stream right after this step's kernels, so it overlaps them (the feed_dict copy of the
        reference, models.py:535, is 40 % of its step); the next call must then pass the same two objects."""
        if self._opt is None:
            raise RuntimeError('model was built without an optimizer')
        pre = self._prefetched
        if pre is not None and pre[0] is batch_data and pre[1] is batch_labels:
            x, labels = pre[2], pre[3]
            torch.cuda.current_stream().wait_event(pre[4])
        else:
            x = self._to_device(batch_data)
            labels = self._labels_to_device(batch_labels)
        self._prefetched = None
        if x.shape[0] != self.batch_size:
            raise ValueError('batch of {} samples, model built for batch_size={}'.format(x.shape[0], self.batch_size))
        lr = self._advance_hyper()
        if self.cuda_graph and not self._trace_eager:
            self._step_graphed(x, labels)
        else:
            self._step_device(x, labels)                    # also the traced steps of fit(profile=True): every launch visible
        if pre is not None and x is pre[2]:
            pre[5].record()                                 # staging pair consumed (copied into the graph's inputs / used)
        if next_batch is not None:
            self._prefetch(next_batch[0], next_batch[1])
        return lr, self._loss

    def loss_handle(self):
        """Start the device -> host copy of the loss of the step just enqueued (4 bytes into one of two pinned slots, on the
        compute stream, before the next step clears the accumulator) and return a handle whose `.value()` waits for that copy
        only.  fit() reads the loss of step i - 1 through such a handle AFTER it has enqueued step i, so the host never leaves
        the GPU without work between two steps (the `sess.run` of the reference returns the loss synchronously,
        models.py:535-536: there the device idles while Python prepares the next feed)."""
        if getattr(self, '_loss_slots', None) is None:
            self._loss_slots = (torch.zeros(2, dtype=torch.float32).pin_memory(), [torch.cuda.Event(), torch.cuda.Event()], [0])
        host, events, turn = self._loss_slots
        i = turn[0]
        turn[0] ^= 1
        host[i:i + 1].copy_(self._loss, non_blocking=True)
        events[i].record()
        return _LossHandle(host, events[i], i)

    def _prefetch(self, batch_data, batch_labels):
        """Start the host -> device copy of a coming batch on the copy stream into one of two staging pairs."""
        if self.vertex_shard or batch_labels is None:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [None, None]
            self._stage_i = 0
        hx = batch_data if isinstance(batch_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(self._as_bmf(batch_data)))
        hl = batch_labels if isinstance(batch_labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(batch_labels))
        if hx.is_cuda or hx.dtype != torch.float32:
            return
        if hx.dim() == 2:
            hx = hx.unsqueeze(2)
        ldt = torch.float32 if self.regression else torch.int32
        i = self._stage_i
        self._stage_i ^= 1
        st = self._stage[i]
        if st is None or st[0].shape != hx.shape or st[1].shape != hl.shape:
            st = (torch.empty(hx.shape, dtype=torch.float32, device=self.device),
                  torch.empty(hl.shape, dtype=ldt, device=self.device),
                  torch.cuda.Event(), torch.cuda.Event(),          # (x, labels, ready, consumed,
                  torch.empty(hx.shape, dtype=torch.float32).pin_memory(),     #  pinned x, pinned labels, copied)
                  torch.empty(hl.shape, dtype=ldt).pin_memory(), torch.cuda.Event())
            st[3].record()
            st[6].record()
            self._stage[i] = st
        cs = self._copy_stream
        hl = hl.to(ldt)
        if not (hx.is_pinned() and hl.is_pinned()):
            # pageable batches (numpy arrays of the reference's dataset protocol): through this pair's pinned buffers, so that
            # the device copy is a real asynchronous DMA; the previous DMA out of them (two steps ago) must have finished
            st[6].synchronize()
            st[4].copy_(hx)
            st[5].copy_(hl)
            hx, hl = st[4], st[5]
        cs.wait_event(st[3])                                # the step that last read this pair has taken its copy
        with torch.cuda.stream(cs):
            st[0].copy_(hx.contiguous(), non_blocking=True)
            st[1].copy_(hl.contiguous(), non_blocking=True)
            st[2].record(cs)
            st[6].record(cs)
        self._prefetched = (batch_data, batch_labels, st[0], st[1], st[2], st[3])

    def _capture(self, x, labels):
        self._gx = x.clone()
        self._gl = labels.clone()
        # one eager warm-up on a side stream (allocator, NCCL), undone afterwards, then capture
        snap = (self._P.w.clone(), [s.clone() for s in self._slots],
                {k: v.clone() for k, v in self._state.items()})
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step_device(self._gx, self._gl)
        torch.cuda.current_stream().wait_stream(side)
        self._P.w.copy_(snap[0])
        for s, c in zip(self._slots, snap[1]):
            s.copy_(c)
        for k, v in snap[2].items():
            self._state[k].copy_(v)
        g = torch.cuda.CUDAGraph()
        if self._dropout_u and hasattr(g, 'register_generator_state'):
            g.register_generator_state(self._gen)
        with torch.cuda.graph(g):
            self._step_device(self._gx, self._gl)
        self._graph_exec = g

    def _step_graphed(self, x, labels):
        """Replay the captured training step on a new batch."""
        if self._graph_exec is None:
            self._capture(x, labels)
        self._gx.copy_(x, non_blocking=True)
        self._gl.copy_(labels, non_blocking=True)
        self._graph_exec.replay()

    def _step_graph_resident(self, x=None, labels=None):
        """Replay on the batch already sitting in the graph's input buffers."""
        if self._graph_exec is None:
            self._capture(x, labels)
        self._graph_exec.replay()


class deepsphere(cgcnn):
    """Spherical CNN front-end: Laplacians and pooling factors from a list of nsides
    (reference `deepsphere`, models.py:1784-1839).  `new=True` (the upstream default) builds the kNN graph of the PyGSP
    fork's `SphereHealpix` (graphs.py); `new=False` the in-repo 8-neighbour graph that the cosmo experiments use."""

    def __init__(self, nsides, indexes=None, use_4=False, sampling='healpix', std=None, full=False, new=True,
                 n_neighbors=8, lmax=None, device_build=None, **kwargs):
        # `device_build`: build the in-repo 8-neighbour graphs (new=False) with the CUDA builder (graph_build.py: same bits as
        # the numpy path for the same lmax, lmax itself by Lanczos instead of ARPACK); default: levels of 2^16 vertices and more
        on_device = (sampling == 'healpix' and not new and not use_4
                     and not (full if not isinstance(full, list) else any(full)) and torch.cuda.is_available())
        if device_build is None:
            device_build = on_device and max(int(n if not isinstance(n, tuple) else n[0]) for n in nsides) >= 128 \
                and (indexes is None or max(len(ix) for ix in indexes if ix is not None) >= (1 << 16))
        if device_build and on_device:
            from . import graph_build
            dev = kwargs.get('device')
            L, p, graphs = graph_build.build_laplacians(nsides, indexes=indexes, device=dev, std=std, lmax=lmax)
            # vertex-sharded mode: every rank keeps its own vertex range only (graph.ShardedGraph, cut from the scipy matrices)
            self._prebuilt_graphs = None if kwargs.get('vertex_shard', False) else graphs
            del graphs
        else:
            L, p = utils.build_laplacians(nsides, indexes=indexes, use_4=use_4, sampling=sampling,
                                          std=std, full=full, new=new, n_neighbors=n_neighbors, lmax=lmax)
        self.sampling = sampling
        if sampling == 'equiangular':
            self.ratio = nsides[0][1] / nsides[0][0] if isinstance(nsides[0], tuple) else 1.
        self.nsides = nsides
        self.pygsp_graphs = [None] * len(nsides)
        if sampling == 'healpix' and not full and not new:
            idx = indexes if indexes is not None else [None] * len(nsides)
            self._nested = [(int(n), idx[i]) for i, n in enumerate(nsides[:len(L)])]
        super(deepsphere, self).__init__(L=L, p=p, **kwargs)
