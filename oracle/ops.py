"""Oracle restatement of the reference's TF1 graph operators, in torch-CPU (autograd gives
the gradients TF1's `optimizer.compute_gradients` would, deepsphere/models.py:831).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function follows the cited lines of
deepsphere/models.py statement by statement (same transposes / reshapes / concatenations, so
the memory order of every intermediate is the reference's).  PINNED through the model cases of
tests/golden/ops_golden.npz: the reference's own, unmodified operator code (chebyshev5, monomials,
pool_*, unpool_*, batch_normalization, bias, fc, learned_histogram, loss) executed through the
TensorFlow stand-in tests/golden/tf_shim.py; tests/test_ops_golden.py holds these functions to 2e-5
of its outputs.  Unpinned: the arithmetic inside a TF op (TensorFlow 1.10 cannot run here).
"""
import numpy as np
import torch
import torch.nn.functional as TF


# ---- emulation of the BF16 storage mode of the product (cgcnn(dtype='bfloat16')): where the product STORES an activation
# (layer input, Chebyshev planes, the contraction's output, the pooled output) or its gradient as bfloat16, the oracle rounds
# the same tensor -- and, in the backward pass, its gradient -- to bfloat16 (round to nearest even) and keeps computing in
# its own dtype.  STORE is None (no rounding) unless OracleCgcnn.bf16_layers names the layer being evaluated.  This is the
# noise floor of BF16 storage itself, against which the CUDA kernels are held (tests/test_bf16.py).
STORE = None


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundBf16Fwd(torch.autograd.Function):          # weights: the contraction reads them rounded, their gradient stays fp32
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _st(t):
    return t if STORE is None else _RoundBf16.apply(t)


def _stw(t):
    return t if STORE is None else _RoundBf16Fwd.apply(t)


def sparse_to_torch(L, dtype=torch.float32):
    """models.py:1044-1047: COO triplets -> SparseTensor, reordered row-major, cast to dtype."""
    L = L.tocoo()
    order = np.lexsort((L.col, L.row))                       # tf.sparse_reorder
    idx = torch.from_numpy(np.vstack([L.row[order], L.col[order]]).astype(np.int64))
    val = torch.from_numpy(L.data[order].astype(np.float64)).to(dtype)
    return torch.sparse_coo_tensor(idx, val, L.shape).coalesce().to_sparse_csr()


def drop_filters(W, Fin, K, mask):
    """models.py:1068-1076 / 1110-1118: W [Fin*K, Fout] times a {0, 1} mask [Fin, Fout] shared by the K orders (the reference
    draws it as tf.nn.dropout(ones, keep) * keep while training; ones otherwise)."""
    if mask is None:
        return W
    Fout = W.shape[1]
    return (W.reshape(Fin, K, Fout) * mask.reshape(Fin, 1, Fout).to(W.dtype)).reshape(Fin * K, Fout)


def chebyshev5(x, L, W, K, mask=None):
    """models.py:1039-1078.  x [N, M, Fin], L torch sparse [M, M], W [Fin*K, Fout]; mask: filter dropout."""
    N, M, Fin = x.shape
    Fout = W.shape[1]
    W = drop_filters(W, Fin, K, mask)
    x0 = x.permute(1, 2, 0)                                   # M x Fin x N        (1049)
    x0 = _st(x0.reshape(M, Fin * N))                          # M x Fin*N          (1050)
    xs = x0.unsqueeze(0)                                      # 1 x M x Fin*N      (1051)
    if K > 1:
        x1 = _st(torch.sparse.mm(L, x0))                      #                    (1056)
        xs = torch.cat([xs, x1.unsqueeze(0)], dim=0)          #                    (1057)
    for k in range(2, K):
        x2 = _st(2 * torch.sparse.mm(L, x1) - x0)             #                    (1059)
        xs = torch.cat([xs, x2.unsqueeze(0)], dim=0)
        x0, x1 = x1, x2                                       #                    (1061)
    xs = xs.reshape(K, M, Fin, N)                             #                    (1062)
    xs = xs.permute(3, 1, 2, 0)                               # N x M x Fin x K    (1063)
    xs = xs.reshape(N * M, Fin * K)                           #                    (1064)
    y = _st(xs @ _stw(W))                                     #                    (1077)
    return y.reshape(N, M, Fout)                              #                    (1078)


def monomials(x, L, W, K, mask=None):
    """models.py:1085-1120 (monomial basis x, Lx, L^2x, ...)."""
    N, M, Fin = x.shape
    Fout = W.shape[1]
    W = drop_filters(W, Fin, K, mask)
    x0 = x.permute(1, 2, 0).reshape(M, Fin * N)
    xs = x0.unsqueeze(0)
    for k in range(1, K):
        x1 = torch.sparse.mm(L, x0)
        xs = torch.cat([xs, x1.unsqueeze(0)], dim=0)
        x0 = x1
    xs = xs.reshape(K, M, Fin, N).permute(3, 1, 2, 0).reshape(N * M, Fin * K)
    return (xs @ W).reshape(N, M, Fout)


def bias(x, b):
    """models.py:1122-1126: one bias per filter, b of shape [1, 1, F]."""
    return x + b.reshape(1, 1, -1)


def batch_normalization(x, moving_mean, moving_var, training, momentum=0.9, eps=1e-5):
    """models.py:1194-1203 -> tf.layers.batch_normalization(axis=-1, center=False,
    scale=False): rank-3 input => non-fused path: tf.nn.moments over axes (0, 1) (biased
    variance), y = (x - mean) * rsqrt(var + eps); moving <- moving*momentum + batch*(1-momentum)
    through UPDATE_OPS (models.py:840-842).  Returns (y, new_moving_mean, new_moving_var)."""
    if training:
        mean = x.mean(dim=(0, 1))
        var = ((x - mean) ** 2).mean(dim=(0, 1))
        new_mean = moving_mean * momentum + mean.detach() * (1 - momentum)
        new_var = moving_var * momentum + var.detach() * (1 - momentum)
    else:
        mean, var = moving_mean, moving_var
        new_mean, new_var = moving_mean, moving_var
    y = (x - mean) * torch.rsqrt(var + eps)
    return y, new_mean, new_var


def pool(x, p, kind, sampling, ratio=1.):
    """models.py:1128-1163: healpix = max/avg over p consecutive vertices (ksize [1,p,1,1],
    SAME padding, M divisible by p); icosahedron = keep the first p vertices."""
    if p > 1:
        if sampling == 'equiangular':
            # models.py:1131-1136, 1150-1155: the vertices as a [sqrt(M/ratio), sqrt(M*ratio)] grid, 2-D pooling with a
            # sqrt(p) x sqrt(p) window, SAME padding (sizes divisible by the window)
            N, M, F = x.shape
            H, W, s = int(round((M / ratio) ** 0.5)), int(round((M * ratio) ** 0.5)), int(round(p ** 0.5))
            xr = x.reshape(N, H // s, s, W // s, s, F)
            return (xr.amax(dim=(2, 4)) if kind == 'max' else xr.mean(dim=(2, 4))).reshape(N, -1, F)
        if sampling == 'icosahedron':
            return x[:, :p, :]
        N, M, F = x.shape
        xr = x.reshape(N, M // p, p, F)
        return xr.max(dim=2).values if kind == 'max' else xr.mean(dim=2)
    return x


def unpool(x, p, kind, sampling, ratio=1.):
    """models.py:1347-1375.  average: healpix repeats each vertex p times, icosahedron pads to
    p vertices with the constant 1; max: healpix puts the value first and p-1 zeros."""
    N, M, F = x.shape
    if kind == 'average':
        if sampling == 'equiangular':                                                  # (1348-1355): repeat_elements on both axes
            H, W, s = int(round((M / ratio) ** 0.5)), int(round((M * ratio) ** 0.5)), int(round(p ** 0.5))
            xr = x.reshape(N, H, W, F).repeat_interleave(s, dim=1).repeat_interleave(s, dim=2)
            return xr.reshape(N, -1, F)
        if sampling == 'icosahedron':
            return TF.pad(x, (0, 0, 0, int(p - M)), mode='constant', value=1.0)   # (1358)
        return x.unsqueeze(2).repeat(1, 1, p, 1).reshape(N, M * p, F)             # (1363-1364)
    if sampling in ('equiangular', 'icosahedron'):
        raise NotImplementedError('unpooling max with {} sampling is not yet implemented'.format(sampling))
    zeros = torch.zeros(N, M, p - 1, F, dtype=x.dtype)
    return torch.cat([x.unsqueeze(2), zeros], dim=2).reshape(N, M * p, F)         # (1372-1375)


def statistics(x, kind, has_fc):
    """models.py:1249-1270 (tf.nn.moments => biased variance)."""
    N, M, F = x.shape
    if kind is None:
        return x.reshape(N, M * F) if has_fc else x
    if kind == 'mean':
        return x.mean(dim=1)
    if kind == 'var':
        return ((x - x.mean(dim=1, keepdim=True)) ** 2).mean(dim=1)
    if kind == 'meanvar':
        mean = x.mean(dim=1)
        var = ((x - x.mean(dim=1, keepdim=True)) ** 2).mean(dim=1)
        return torch.cat([mean, var], dim=1)
    if kind == 'max':
        return x.max(dim=1).values
    raise ValueError('Unknown statistical layer {}'.format(kind))


def learned_histogram(x, centers, widths, bins=20, initial_range=2):
    """models.py:1166-1192; centers / widths are the trainable [1, 1, F, bins] variables `stat/centers`, `stat/widths`."""
    N, M, F = x.shape
    x4 = x.unsqueeze(3)                                                   # 1186
    w = widths.reshape(1, 1, F, bins).abs()                               # 1188
    dist = (x4 - centers.reshape(1, 1, F, bins)).abs()                    # 1189
    hist = torch.relu(1 - dist * w).mean(dim=1) * (bins / initial_range / 4)   # 1190
    return hist.reshape(N, bins * F)                                      # 1266


def histogram_init(F, bins=20, initial_range=2, dtype=torch.float32):
    """Initial values of `stat/centers` (linspace per feature, 1173-1179) and `stat/widths` (4 range / bins, 1181-1185)."""
    centers = torch.linspace(-float(initial_range), float(initial_range), bins, dtype=dtype).repeat(F, 1).reshape(1, 1, F, bins)
    widths = torch.full((1, 1, F, bins), 4 * initial_range / bins, dtype=dtype)
    return centers, widths


def fc(x, W, b=None):
    """models.py:1205-1212."""
    y = x @ W
    if b is not None:
        y = y + b
    return y


ACTIVATIONS = {
    'relu': torch.relu, 'elu': TF.elu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid,
    'softplus': TF.softplus, 'leaky_relu': lambda x: TF.leaky_relu(x, 0.2),   # tf.nn.leaky_relu alpha=0.2
    'relu6': TF.relu6,
}


CLASS_WEIGHTS = (0.34130685, 318.47388343, 14.93759951)     # models.py:796 (hard-coded, 3 classes)


def loss(logits, labels, regression, weights_and_std, regularization, weighted=False, mask_data=None):
    """models.py:777-820 (cross-entropy / MSE + regularization * sum_i(l2_loss(W_i)/std_i^2)/sum_i size_i).
    `weights_and_std` = [(W_i, std_i)] in creation order (models.py:862-869)."""
    if regression:
        predictions = logits.squeeze()
        if mask_data is not None:                                          # models.py:785-787 (`hasattr(self, 'train_mask')`)
            predictions = predictions * mask_data[..., -2]
            labels = labels * mask_data[..., -1]
        data_loss = ((labels - predictions) ** 2).mean()                   # tf.losses.mean_squared_error
    else:
        C = logits.shape[-1]
        ce = TF.cross_entropy(logits.reshape(-1, C), labels.reshape(-1).long(), reduction='none')
        if weighted:                                                       # models.py:795-802: one_hot(labels, 3) @ weights^T
            if C != 3:
                raise ValueError('weighted cross-entropy is defined for 3 classes')
            ce = ce * torch.tensor(CLASS_WEIGHTS, dtype=ce.dtype)[labels.reshape(-1).long()]
        data_loss = ce.mean()
    n_weights = sum(int(np.prod(W.shape)) for W, _ in weights_and_std)
    reg = sum(((W ** 2).sum() / 2) / std ** 2 for W, std in weights_and_std) / n_weights
    return data_loss + regularization * reg, data_loss
