"""Tensor-level wrappers of the libdsb200 kernels (allocation + argument plumbing only).

Every function takes / returns contiguous fp32 CUDA tensors in the reference's [B, M, F]
layout and enqueues on torch's current stream.  The arithmetic lives in csrc/*.cu; PyTorch is
used for device memory and streams, nothing else.  Reference call sites are cited per function.
"""
import os

import torch

from ._consts import ACT, POOL, STAT, UNPOOL_REPEAT, UNPOOL_ZEROFILL  # noqa: F401
from ._lib import call


BF16 = torch.bfloat16


def _new(shape, like, dtype=None):
    return torch.empty(shape, dtype=like.dtype if dtype is None else dtype, device=like.device)


# ---- BF16 mode (cgcnn(..., dtype='bfloat16'); include/dsb200.h, "BF16 mode").  The ChebConv operators below take bfloat16
# activations natively where a bf16 kernel covers the shape; everywhere else the tensor is converted with these two kernels
# and the fp32 operator runs (models.py casts at the boundaries of the ChebConv chain, so the other operators only see fp32).
def to_f32(x):
    if x is None or x.dtype == torch.float32:
        return x
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call('cast_bf16_to_f32', x, out, x.numel())
    return out


def to_bf16(x, out=None):
    if x is None or (x.dtype == BF16 and out is None):
        return x
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    call('cast_f32_to_bf16', x, out, x.numel())
    return out


def _lib_int(name, *args):
    from ._lib import LIB
    return int(getattr(LIB.load(), 'dsb200_' + name)(*args))


def spmm(graph, x, y=None, z=None, out=None, alpha=1.0, beta=0.0, gamma=0.0, transpose=False, rows=None, reverse=False):
    """out = alpha * L x + beta * y + gamma * z  (tf.sparse_tensor_dense_matmul, models.py:1056,1059).
    `rows`: compute only the first `rows` rows (vertex-sharded mode, one sample: the trailing ghost rows of x are
    inputs only).  `reverse` (pipelined kernel only; same result): walk memory from its end to its start (`ZIGZAG`)."""
    B, M, F = x.shape
    if rows is not None:
        if B != 1:
            raise ValueError('rows= needs a single sample')
        M = rows
    rowptr, col, val, width = graph.bwd if transpose else graph.fwd
    which = 'bwd' if transpose else 'fwd'
    if x.dtype == BF16:
        pp = _pipe(graph, which, B, M, F, esize=2, min_items=1) if rows is None else None
        if pp is None:                                   # no bf16 gather kernels: fp32 operator between two casts
            res = spmm(graph, to_f32(x), to_f32(y), to_f32(z), None, alpha, beta, gamma, transpose, rows, reverse)
            return to_bf16(res, out)
        if out is None:
            out = _new((B, M, F), x)
        call('spmm_pipe_bf16', *pp.args(USE_DENSE_GROUPS, reverse), B, M, F, x, y, z, out, alpha, beta, gamma)
        return out
    if out is None:
        out = _new((B, M, F), x)
    pp = _pipe(graph, which, B, M, F) if rows is None else None
    if pp is not None:
        call('spmm_pipe', *pp.args(USE_DENSE_GROUPS, reverse), B, M, F, x, y, z, out, alpha, beta, gamma)
    elif rowptr is None:
        rec = _rec(graph, which) if rows is None else None         # the record stride is the graph's M
        if rec is not None:
            call('spmm_ellp', rec, col, val, width, B, M, F, x, y, z, out, alpha, beta, gamma)
        else:
            call('spmm_ell', col, val, width, B, M, F, x, y, z, out, alpha, beta, gamma)
    else:
        call('spmm_csr', rowptr, col, val, B, M, F, x, y, z, out, alpha, beta, gamma)
    return out


USE_PIPE = True           # bulk-asynchronous pipelined sparse product (csrc/sparse_pipe.cu) where a tile plan exists
USE_DENSE_GROUPS = True   # pipelined kernel: 4-row groups with dense 4 x 16 weight blocks (each distinct neighbour read once)
PIPE_MIN_ITEMS = int(os.environ.get('DSB200_PIPE_MIN_ITEMS', 4 * 296))  # (row tiles x samples) below which the pipeline has
                                                                        # nothing to overlap: gather kernels


# Zigzag recurrence: consecutive sparse products of a layer sweep memory in opposite directions, so that each one starts on the
# rows its predecessor touched last (still in the 126 MB L2) instead of the rows it touched first (evicted long ago: one step at
# the large levels moves 200-270 MB).  The last forward step runs end -> start because the contraction that follows walks
# start -> end; the first adjoint step runs end -> start because the data-gradient contraction before it stopped at the end.
# 0: every step start -> end.  (Results do not depend on the order: every item writes its own rows.)
ZIGZAG = int(os.environ.get('DSB200_ZIGZAG', 0))


def _rev_fwd(k, K):
    """Direction of forward step k (1..K-1): step K-1 reversed, alternating below it (ZIGZAG = 2: the other parity)."""
    return bool(ZIGZAG) and ((K - 1 - k) % 2 == 0) == (ZIGZAG == 1)


def _rev_bwd(k, K):
    """Direction of adjoint step k (K-2..0): step K-2 (the first to run) reversed, alternating after it."""
    return bool(ZIGZAG) and ((K - 2 - k) % 2 == 0) == (ZIGZAG == 1)


def set_pipe(on, dense=None):
    global USE_PIPE, USE_DENSE_GROUPS
    USE_PIPE = bool(on)
    if dense is not None:
        USE_DENSE_GROUPS = bool(dense)


def _pipe(graph, which, B, M, F, esize=4, min_items=None):
    if not USE_PIPE or not hasattr(graph, 'pipe_plan'):
        return None
    plan = graph.pipe_plan(which, F, esize)
    if plan is None or plan.T * B < (PIPE_MIN_ITEMS if min_items is None else min_items):
        return None
    return plan


USE_RECORDS = True        # packed slot-major ELL records (csrc/sparse.cu: spmm_rec_kernel) instead of row-major col[] / val[]


def set_records(on):
    global USE_RECORDS
    USE_RECORDS = bool(on)


def _rec(graph, which):
    return getattr(graph, which + '_rec', None) if USE_RECORDS else None


USE_TILE = os.environ.get('DSB200_TILE', '1') != '0'   # fused tile recurrence (csrc/cheb_tile.cu): all K - 1 sparse products of
                                                       # a layer in one launch
# Where the fused kernel is used.  Measured inside the cosmo training step (profiles/r02_trace_step_*.txt): the per-step
# pipelined kernels find the plane the previous step has just written in the 126 MB L2 (28 us per step at layer 2 instead of 42 us
# with cold operands), so the HBM traffic the fusion removes is already L2 traffic there and the halo recomputation (1.43x) and
# idle lanes (0.88) of the tile kernel decide: it wins where there are few (tile, sample, 16-feature chunk) items (layer 1: 54 vs
# 112 us, layers 4-5), loses at 4096 items (layer 2: 156 vs 111 us).  With operands that exceed L2 (tools/bench_tile.py) it wins
# the forward direction everywhere.  Items = nt * B * F / 16.
TILE_MIN_ITEMS = int(os.environ.get('DSB200_TILE_MIN_ITEMS', 128))
TILE_FWD_MAX_ITEMS = int(os.environ.get('DSB200_TILE_FWD_MAX_ITEMS', 1024))
TILE_BWD_MAX_ITEMS = int(os.environ.get('DSB200_TILE_BWD_MAX_ITEMS', 256))


def set_tile(on):
    global USE_TILE
    USE_TILE = bool(on)


def _tile(graph, which, K, B, F):
    if not USE_TILE or K < 2 or K > 8 or F % 16 or not hasattr(graph, 'stencil_plan'):
        return None
    plan = graph.stencil_plan(which, K - 1)
    if plan is None:
        return None
    items = plan.nt * B * (F // 16)
    if items < TILE_MIN_ITEMS or items > (TILE_BWD_MAX_ITEMS if which == 'bwd' else TILE_FWD_MAX_ITEMS):
        return None
    return plan


def cheb_basis(graph, x, K, monomial=False):
    """Orders 1..K-1 of the Chebyshev basis of x (models.py:1049-1061) as [K-1, B, M, F]; `monomial`: the basis
    x, Lx, L^2 x, ... of `cgcnn.monomials` instead (models.py:1094-1103)."""
    if K == 1:
        return None
    B, M, F = x.shape
    if x.dtype == BF16:
        if _pipe(graph, 'fwd', B, M, F, esize=2, min_items=1) is None:
            return to_bf16(cheb_basis(graph, to_f32(x), K, monomial))
        basis = _new((K - 1, B, M, F), x)
        for k in range(1, K):
            if monomial or k == 1:
                spmm(graph, x if k == 1 else basis[k - 2], out=basis[k - 1], reverse=_rev_fwd(k, K))
            else:
                spmm(graph, basis[k - 2], x if k == 2 else basis[k - 3], None, out=basis[k - 1], alpha=2.0, beta=-1.0,
                     reverse=_rev_fwd(k, K))
        return basis
    basis = _new((K - 1, B, M, F), x)
    tp = _tile(graph, 'fwd', K, B, F)
    if tp is not None:
        call('cheb_tile', tp.p2v, tp.wts, tp.nt, tp.T, K - 1, B, M, F, 0, 1 if monomial else 0, x, basis)
        return basis
    if monomial:
        for k in range(1, K):                                                  # X_k = L X_(k-1)
            spmm(graph, x if k == 1 else basis[k - 2], out=basis[k - 1])
        return basis
    if _pipe(graph, 'fwd', B, M, F) is not None:
        spmm(graph, x, out=basis[0], reverse=_rev_fwd(1, K))                   # X1 = L X0
        for k in range(2, K):                                                  # Xk = 2 L X(k-1) - X(k-2)
            spmm(graph, basis[k - 2], x if k == 2 else basis[k - 3], None, out=basis[k - 1], alpha=2.0, beta=-1.0,
                 reverse=_rev_fwd(k, K))
        return basis
    rowptr, col, val, width = graph.fwd
    rec = _rec(graph, 'fwd')
    if rowptr is None and rec is not None:
        call('cheb_basis_ellp', rec, col, val, width, B, M, F, K, x, basis)
    else:
        call('cheb_basis', rowptr, col, val, width, B, M, F, K, x, basis)
    return basis


def cheb_basis_bwd(graph, G, monomial=False):
    """Adjoint of the recurrence: G [K, B, M, F] -> dLoss/dx [B, M, F] (G is overwritten on the gather path).
    `monomial`: adjoint of X_k = L X_(k-1), i.e. A_k = G_k + L^T A_(k+1)."""
    K, B, M, F = G.shape
    if G.dtype == BF16:
        if K >= 2 and _pipe(graph, 'bwd', B, M, F, esize=2, min_items=1) is None:
            return to_bf16(cheb_basis_bwd(graph, to_f32(G), monomial))
        for k in range(K - 2, -1, -1):
            a2 = G[k + 2] if (k + 2 <= K - 1 and not monomial) else None
            spmm(graph, G[k + 1], G[k], a2, out=G[k], alpha=2.0 if (k >= 1 and not monomial) else 1.0, beta=1.0,
                 gamma=-1.0 if a2 is not None else 0.0, transpose=True, reverse=_rev_bwd(k, K))
        return G[0]
    tp = _tile(graph, 'bwd', K, B, F)
    if tp is not None:
        dx = _new((B, M, F), G)
        call('cheb_tile', tp.p2v, tp.wts, tp.nt, tp.T, K - 1, B, M, F, 1, 1 if monomial else 0, G, dx)
        return dx
    if monomial:
        for k in range(K - 2, -1, -1):
            spmm(graph, G[k + 1], G[k], None, out=G[k], alpha=1.0, beta=1.0, transpose=True)
        return G[0]
    if _pipe(graph, 'bwd', B, M, F) is not None and K >= 2:
        # A_(K-1) = G_(K-1);  A_k = G_k + c_k L^T A_(k+1) - A_(k+2), in place on the planes of G (see csrc/sparse.cu)
        for k in range(K - 2, -1, -1):
            a2 = G[k + 2] if k + 2 <= K - 1 else None
            spmm(graph, G[k + 1], G[k], a2, out=G[k], alpha=2.0 if k >= 1 else 1.0, beta=1.0,
                 gamma=-1.0 if a2 is not None else 0.0, transpose=True, reverse=_rev_bwd(k, K))
        return G[0]
    rowptr, col, val, width = graph.bwd
    rec = _rec(graph, 'bwd')
    if rowptr is None and rec is not None:
        call('cheb_basis_bwd_ellp', rec, col, val, width, B, M, F, K, G)
    else:
        call('cheb_basis_bwd', rowptr, col, val, width, B, M, F, K, G)
    return G[0]


def _workspace(Fin, K, Fout, device):
    """Scratch for the split-weight image of the tensor-core contraction (one allocation per call from
    torch's caching allocator: stream-ordered, CUDA-graph safe)."""
    from ._lib import LIB
    n = int(LIB.load().dsb200_contract_workspace_bytes(Fin, K, Fout))
    return torch.empty(n, dtype=torch.uint8, device=device)


def _workspace_bf16(R, Fin, K, Fout, device):
    return torch.empty(_lib_int('contract_workspace_bytes_bf16', R, Fin, K, Fout), dtype=torch.uint8, device=device)


def contract_fwd(x, basis, W, K, Fout, bn_sums=None):
    """Y = [X_0 .. X_{K-1}] . W  (tf.matmul, models.py:1062-1078), optional batch-norm sums."""
    B, M, Fin = x.shape
    if x.dtype == BF16:
        if not _lib_int('contract_bf16_supported', 0, B * M, Fin, K, Fout):
            return to_bf16(contract_fwd(to_f32(x), to_f32(basis), W, K, Fout, bn_sums))
        y = _new((B, M, Fout), x)
        call('cheb_contract_fwd_bf16', x, basis, W, y, B * M, Fin, K, Fout, bn_sums, _workspace_bf16(B * M, Fin, K, Fout, x.device))
        return y
    y = _new((B, M, Fout), x)
    call('cheb_contract_fwd', x, basis, W, y, B * M, Fin, K, Fout, bn_sums, _workspace(Fin, K, Fout, x.device))
    return y


POOL4_ACTS = ('relu', 'identity', 'leaky_relu', 'relu6')   # non-decreasing in floating point too: pooling first is bit-identical
USE_POOL4 = os.environ.get('DSB200_POOL4', '1') != '0'
POOL4_TRAIN = os.environ.get('DSB200_POOL4_TRAIN', '0') == '1'   # also while training (slower: see models._ChebStage.forward)


def contract_pool4_supported(x, K, Fout):
    """The forward contraction can emit the max over every 4 consecutive vertices itself."""
    B, M, Fin = x.shape
    if not USE_POOL4 or M % 4:
        return False
    if x.dtype == BF16:
        return bool(_lib_int('contract_bf16_supported', 3, B * M, Fin, K, Fout))
    return bool(_lib_int('contract_pool4_supported', B * M, Fin, K, Fout))


def contract_fwd_pool4(x, basis, W, K, Fout, bn_sums=None, want_y=True):
    """(Y or None, max over 4 consecutive vertices of Y [B, M/4, Fout]) in one launch: the HEALPix max-pool fused into the
    contraction's epilogue (models.py:1062-1078 + 1139-1142).  want_y=False (inference): Y is never written."""
    B, M, Fin = x.shape
    y = _new((B, M, Fout), x) if want_y else None
    yp = _new((B, M // 4, Fout), x)
    if x.dtype == BF16:
        call('cheb_contract_fwd_pool4_bf16', x, basis, W, y, yp, B * M, Fin, K, Fout, bn_sums, _workspace_bf16(B * M, Fin, K, Fout, x.device))
    else:
        call('cheb_contract_fwd_pool4', x, basis, W, y, yp, B * M, Fin, K, Fout, bn_sums, _workspace(Fin, K, Fout, x.device))
    return y, yp


def contract_bwd_data(dy, W, K, Fin):
    B, M, Fout = dy.shape
    if dy.dtype == BF16:
        if not _lib_int('contract_bf16_supported', 1, B * M, Fin, K, Fout):
            return to_bf16(contract_bwd_data(to_f32(dy), W, K, Fin))
        G = _new((K, B, M, Fin), dy)
        call('cheb_contract_bwd_data_bf16', dy, W, G, B * M, Fin, K, Fout, _workspace_bf16(B * M, Fin, K, Fout, dy.device))
        return G
    G = _new((K, B, M, Fin), dy)
    call('cheb_contract_bwd_data', dy, W, G, B * M, Fin, K, Fout, _workspace(Fin, K, Fout, dy.device))
    return G


def contract_bwd_weight(x, basis, dy, dW, K):
    B, M, Fin = x.shape
    if x.dtype == BF16:
        if _lib_int('contract_bf16_supported', 2, B * M, Fin, K, dy.shape[2]):
            call('cheb_contract_bwd_weight_bf16', x, basis, dy, dW, B * M, Fin, K, dy.shape[2])
            return
        x, basis, dy = to_f32(x), to_f32(basis), to_f32(dy)
    call('cheb_contract_bwd_weight', x, basis, dy, dW, B * M, Fin, K, dy.shape[2])


def _sfx(t):
    return '_bf16' if t.dtype == BF16 else ''


def bn_finalize(sums, count, F, mean, rstd, moving_mean=None, moving_var=None, eps=1e-5, momentum=0.9):
    call('bn_finalize', sums, float(count), F, eps, momentum, mean, rstd, moving_mean, moving_var)


def bn_from_moving(moving_mean, moving_var, mean, rstd, eps=1e-5):
    call('bn_from_moving', moving_mean, moving_var, moving_mean.numel(), eps, mean, rstd)


def bn_act_pool_fwd(y, mean, rstd, bias, p, pool, act):
    B, M, F = y.shape
    out = _new((B, M // p, F), y)
    call('bn_act_pool_fwd' + _sfx(y), y, mean, rstd, bias, out, B, M, F, p, POOL[pool], ACT[act])
    return out


def bn_act_pool_bwd(y, mean, rstd, bias, dout, p, pool, act, want_gz=True, out=None, sums=None):
    """gz = dLoss/d(normalised + bias) and the per-feature sums (sum gz, sum gz*xhat); with
    want_gz=False only the sums (pass 1 of the batch-norm backward)."""
    B, M, F = y.shape
    gz = (_new((B, M, F), y) if out is None else out) if want_gz else None
    if sums is None:
        sums = torch.zeros(2 * F, dtype=torch.float64, device=y.device)
    call('bn_act_pool_bwd' + _sfx(y), y, mean, rstd, bias, dout, gz, sums, B, M, F, p, POOL[pool], ACT[act])
    return gz, sums


def bn_relu_pool_sums_supported(F, p, pool, act):
    return act == 'relu' and (p == 1 or pool == 'max') and F % 4 == 0 and F <= 1024


def bn_relu_pool_sums(out, dout, bias, sums=None):
    """Pass 1 of the batch-norm backward for relu (+ max pool) from the pooled tensors only."""
    F = out.shape[-1]
    if sums is None:
        sums = torch.zeros(2 * F, dtype=torch.float64, device=out.device)
    call('bn_relu_pool_sums' + _sfx(out), out, dout, bias, sums, out.numel() // F, F)
    return sums


def gather_rows(x, idx):
    """out[i] = x[idx[i]] for a [M, F] matrix (send buffer of the halo exchange)."""
    F = x.shape[-1]
    out = _new((idx.numel(), F), x)
    call('gather_rows', x, idx, idx.numel(), F, out)
    return out


def bn_act_pool_bwd_dy(y, mean, rstd, bias, dout, sums, count, p, pool, act, out=None):
    """Pass 2 of the batch-norm backward: dy from y, dout and the (all-reduced) sums."""
    B, M, F = y.shape
    dy = _new((B, M, F), y) if out is None else out
    call('bn_act_pool_bwd_dy' + _sfx(y), y, mean, rstd, bias, dout, sums, float(count), dy, B, M, F, p, POOL[pool], ACT[act])
    return dy


def bn_bwd_apply(y, mean, rstd, sums, count, gz):
    B, M, F = y.shape
    call('bn_bwd_apply', y, mean, rstd, sums, float(count), gz, B * M, F)


def sums_to_float(sums, out):
    call('sums_to_float', sums, out, out.numel())


def vertex_resize(x, Mout, fill, out=None):
    B, Min, F = x.shape
    if out is None:
        out = _new((B, Mout, F), x)
    call('vertex_resize', x, out, B, Min, Mout, F, float(fill))
    return out


def unpool_fwd(x, p, mode):
    B, M, F = x.shape
    out = _new((B, M * p, F), x)
    call('unpool_fwd', x, out, B, M, F, p, mode)
    return out


def unpool_bwd(dout, p, mode):
    B, Mp, F = dout.shape
    dx = _new((B, Mp // p, F), dout)
    call('unpool_bwd', dout, dx, B, Mp // p, F, p, mode)
    return dx


def concat_f(a, b):
    B, M, Fa = a.shape
    Fb = b.shape[2]
    out = _new((B, M, Fa + Fb), a)
    call('concat_f', a, b, out, B * M, Fa, Fb)
    return out


def split_f(dout, Fa, db=None):
    """da = dout[..., :Fa]; db (+)= dout[..., Fa:] (accumulates when db is given)."""
    B, M, F = dout.shape
    da = _new((B, M, Fa), dout)
    acc = 1 if db is not None else 0
    if db is None:
        db = _new((B, M, F - Fa), dout)
    call('split_f', dout, da, db, B * M, Fa, F - Fa, acc)
    return da, db


def stat_fwd(x, kind):
    B, M, F = x.shape
    out = _new((B, 2 * F if kind == 'meanvar' else F), x)
    call('stat_fwd', x, out, B, M, F, STAT[kind])
    return out


def stat_bwd(x, dout, kind):
    B, M, F = x.shape
    dx = _new((B, M, F), x)
    call('stat_bwd', x, dout, dx, B, M, F, STAT[kind])
    return dx


HIST_BINS, HIST_RANGE = 20, 2.0          # hard-coded in the reference (models.py:1166, 1264)


def hist_fwd(x, centers, widths, bins=HIST_BINS, initial_range=HIST_RANGE):
    """cgcnn.learned_histogram (models.py:1166-1192) -> [B, F*bins]."""
    B, M, F = x.shape
    out = _new((B, F * bins), x)
    call('hist_fwd', x, centers, widths, out, B, M, F, bins, bins / initial_range / 4)
    return out


def hist_bwd(x, centers, widths, dout, dcenters, dwidths, bins=HIST_BINS, initial_range=HIST_RANGE):
    """dx of the learned histogram; the gradients of the centers / widths are accumulated into dcenters / dwidths."""
    B, M, F = x.shape
    dx = _new((B, M, F), x)
    call('hist_bwd', x, centers, widths, dout, dx, dcenters, dwidths, B, M, F, bins, bins / initial_range / 4)
    return dx


def fc_fwd(x, W, b, act):
    N, Min = x.shape
    y = _new((N, W.shape[1]), x)
    call('fc_fwd', x, W, b, y, N, Min, W.shape[1], ACT[act])
    return y


def fc_bwd(x, W, y, dy, dW, db, act, need_dx=True):
    N, Min = x.shape
    dx = _new((N, Min), x) if need_dx else None
    call('fc_bwd', x, W, y, dy, dx, dW, db, N, Min, W.shape[1], ACT[act])
    return dx


def softmax_ce(logits, labels, loss_out, grad_scale=1.0, need_grad=True, class_weights=None):
    """Mean sparse-softmax cross-entropy into loss_out (models.py:790-804); `class_weights` [C]: the `weighted` variant."""
    C = logits.shape[-1]
    R = logits.numel() // C
    dlogits = torch.empty_like(logits) if need_grad else None
    if class_weights is not None:
        call('softmax_ce_weighted', logits, labels, class_weights, dlogits, loss_out, R, C, grad_scale)
    else:
        call('softmax_ce', logits, labels, dlogits, loss_out, R, C, grad_scale)
    return dlogits


def mse(pred, target, loss_out, grad_scale=1.0, need_grad=True):
    dpred = torch.empty_like(pred) if need_grad else None
    call('mse', pred, target, dpred, loss_out, pred.numel(), grad_scale)
    return dpred


def softmax_argmax(logits, want_probs=False):
    C = logits.shape[-1]
    R = logits.numel() // C
    probs = torch.empty_like(logits) if want_probs else None
    pred = torch.empty(logits.shape[:-1], dtype=torch.int32, device=logits.device)
    call('softmax_argmax', logits, probs, pred, R, C)
    return probs, pred


def l2_reg(w, coef, g, loss_out):
    call('l2_reg', w, coef, g, w.numel(), loss_out)


def optimizer_step(kind, w, g, s1, s2, hyper):
    call('optimizer_step', kind, w, g, s1, s2, hyper, w.numel())


def fill(x, v):
    call('fill', x, float(v), x.numel())


def axpy(x, y, a=1.0):
    call('axpy', x, y, float(a), x.numel())


def mul_channel(a, x, ch):
    """a * x[..., ch] for a [B, M] (or flat) and x [B, M, F] (masked regression, models.py:785-787)."""
    F = x.shape[-1]
    out = torch.empty_like(a)
    call('mul_channel', a, x, F, ch % F, out, a.numel())
    return out


def scale(x, a):
    call('scale', x, float(a), x.numel())


def dropout(x, u, keep):
    """tf.nn.dropout with an explicit uniform tensor (models.py:1279-1280); also its own backward."""
    out = torch.empty_like(x)
    call('dropout', x, u, float(keep), out, x.numel())
    return out


def set_tensor_cores(on):
    """Enable / disable the tcgen05 contraction kernels (FP32 FFMA fallback); returns the previous setting."""
    from ._lib import LIB
    return bool(LIB.load().dsb200_set_tensor_cores(1 if on else 0))


def smallconv_supported(Fin, K, Fout):
    from ._lib import LIB
    return bool(LIB.load().dsb200_smallconv_supported(Fin, K, Fout))


def cheb_gram(x, basis, K, shape=None):
    """Column sums and Gram matrix of the Chebyshev basis (double[Q + Q*Q], Q = K*Fin)."""
    B, M, Fin = shape if shape is not None else x.shape
    Q = K * Fin
    gram = torch.zeros(Q + Q * Q, dtype=torch.float64, device=x.device)
    call('cheb_gram', x, basis, B * M, Fin, K, gram)
    return gram


def bn_sums_from_gram(gram, W, Fin, K, Fout):
    sums = torch.empty(2 * Fout, dtype=torch.float64, device=W.device)
    call('bn_sums_from_gram', gram, W, Fin, K, Fout, sums)
    return sums


def interleave(x):
    """[B, M, F] -> [M, B, F] (sample-interleaved layout for the sparse products of a narrow input layer)."""
    B, M, F = x.shape
    out = _new((M, B, F), x)
    call('interleave', x, out, B, M, F)
    return out


def smallconv_has_argmax(Fout, p, pool, act):
    from ._lib import LIB
    return bool(LIB.load().dsb200_smallconv_has_argmax(Fout, p, POOL[pool], ACT[act]))


def smallconv_fwd(x, basis, W, K, Fout, mean, rstd, bias, p, pool, act, shape=None, want_argmax=False):
    """`shape` = (B, M, Fin) when x / basis are in the sample-interleaved layout [M, B, Fin].
    Returns (out, argmax | None)."""
    B, M, Fin = shape if shape is not None else x.shape
    out = _new((B, M // p, Fout), x)
    arg = torch.empty((B, M // p, Fout), dtype=torch.uint8, device=x.device) if want_argmax else None
    call('smallconv_fwd', x, basis, W, mean, rstd, bias, out, B, M, Fin, K, Fout, p, POOL[pool], ACT[act],
         1 if shape is not None else 0, arg)
    return out, arg


def smallconv_bwd(x, basis, W, K, Fout, mean, rstd, bias, dout, p, pool, act, shape=None, out=None, argmax=None):
    B, M, Fin = shape if shape is not None else x.shape
    acc = torch.zeros((K * Fin + 2) * Fout, dtype=torch.float64, device=x.device)
    call('smallconv_bwd', x, basis, W, mean, rstd, bias, dout, acc, B, M, Fin, K, Fout, p, POOL[pool], ACT[act],
         1 if shape is not None else 0, out, argmax)
    return acc


def smallconv_bwd_finalize(acc, gram, W, mean, rstd, count, Fin, K, Fout, dW, dbias):
    call('smallconv_bwd_finalize', acc, gram, W, mean, rstd, float(count), Fin, K, Fout, dW, dbias)
