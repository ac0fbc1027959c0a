"""A minimal eager stand-in for the `tensorflow` 1.10 module, backed by torch-CPU tensors.

TEST INFRASTRUCTURE ONLY.  Purpose: execute the UNMODIFIED /root/reference/deepsphere/models.py (graph
construction code of `cgcnn` / `deepsphere`: chebyshev5, monomials, pool_*, unpool_*, batch_normalization, bias, fc,
learned_histogram, _inference, _decoder, loss, training) in this container, where TensorFlow 1.10 cannot be installed,
so that its outputs can be committed as golden vectors (tests/golden/make_ops_golden.py -> ops_golden.npz).

How a static TF1 graph becomes an eager run: the feeds (`data`, `labels`, `training`) are given to the shim BEFORE the
reference constructor runs (`feed(...)`); `tf.placeholder*` return them, every `tf.*` call computes its value at once,
`tf.get_variable` returns a leaf tensor (preset value if one was registered with `preset(...)`, else the initializer's
draw), and the optimizer handed to the reference (`GradientRecorder`) records d loss / d variable with torch autograd
-- what `optimizer.compute_gradients` differentiates in TF1.

What this file decides itself (TF internals, NOT reference code -- they stay unpinned): the arithmetic of
`tf.layers.batch_normalization` (non-fused rank-3 path: biased batch moments, moving <- momentum * moving +
(1 - momentum) * batch), `tf.nn.max_pool / avg_pool` windows, `sparse_softmax_cross_entropy_with_logits`,
`tf.losses.mean_squared_error`, the summation order of `sparse_tensor_dense_matmul`, the truncated-normal draw.
Everything else -- which ops are called, in which order, on which axes, with which shapes, scopes, names,
standard deviations and regularisers -- is the reference's own code.
"""
import collections
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

float32, float64, int32, int64 = torch.float32, torch.float64, torch.int32, torch.int64
bool = torch.bool                                                   # noqa: A001  (tf.bool)

_S = types.SimpleNamespace(feeds={}, presets={}, scope=[], variables=collections.OrderedDict(), state=collections.OrderedDict(),
                           updates=collections.OrderedDict(), grads=collections.OrderedDict(), rs=np.random.RandomState(0),
                           dropout_masks=[], uniq={})


def reset(seed=0):
    _S.feeds, _S.presets, _S.scope = {}, {}, []
    _S.variables, _S.state = collections.OrderedDict(), collections.OrderedDict()
    _S.updates, _S.grads = collections.OrderedDict(), collections.OrderedDict()
    _S.rs = np.random.RandomState(seed)
    _S.dropout_masks, _S.uniq = [], {}


def feed(data, labels, training):
    _S.feeds = {'data': _t(data), 'labels': torch.as_tensor(np.asarray(labels)), 'training': training}


def preset(values):
    _S.presets = dict(values)


def variables():
    return _S.variables


def state():
    return _S.state


def updates():
    return _S.updates


def gradients():
    return _S.grads


def dropout_masks():
    return _S.dropout_masks


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    a = np.asarray(x)
    if dtype is None and a.dtype == np.float64:
        dtype = torch.float32                                      # python floats are float32 constants in TF1 graphs
    return torch.as_tensor(a, dtype=dtype)


# torch.Tensor has no get_shape(); the reference calls it on every tensor (models.py:1041 ...)
torch.Tensor.get_shape = lambda self: self.shape


# ---------------------------------------------------------------- graph / session scaffolding (no-ops)
class Graph(object):
    @contextlib.contextmanager
    def as_default(self):
        yield self

    def finalize(self):
        pass


class _Iterator(object):
    def get_next(self):
        return None, None


class _Dataset(object):
    @staticmethod
    def from_generator(gen, output_types=None):
        return _Dataset()

    def prefetch(self, n):
        return self

    def make_initializable_iterator(self):
        return _Iterator()

    def make_one_shot_iterator(self):
        return _Iterator()


data = types.SimpleNamespace(Dataset=_Dataset)


def placeholder_with_default(default, shape, name=None):
    v = _S.feeds[name]
    if isinstance(shape, tuple):
        assert tuple(v.shape) == tuple(int(s) for s in shape), (name, tuple(v.shape), shape)
    else:                                                          # `(self.batch_size)` is an int (models.py:677)
        assert v.shape[0] == int(shape), (name, tuple(v.shape), shape)
    return v


def placeholder(dtype, shape, name=None):
    return _S.feeds[name]


@contextlib.contextmanager
def name_scope(name):
    yield


@contextlib.contextmanager
def variable_scope(name):
    _S.scope.append(name)
    try:
        yield
    finally:
        _S.scope.pop()


@contextlib.contextmanager
def control_dependencies(deps):
    yield


def identity(x, name=None):
    return x


def _scoped(name):
    return '/'.join(_S.scope + [name])


def _register(name, value, trainable):
    full = _scoped(name)
    if full in _S.variables or full in _S.state:                   # tf.Variable uniquifies: name, name_1, ...
        n = _S.uniq.get(full, 0) + 1
        _S.uniq[full] = n
        full = '{}_{}'.format(full, n)
    v = _t(_S.presets[full], torch.float32).clone() if full in _S.presets else value
    v = v.detach().clone()
    v.op = types.SimpleNamespace(name=full)
    if trainable:
        v.requires_grad_(True)
        _S.variables[full] = v
    else:
        _S.state[full] = v
    return v


def truncated_normal_initializer(mean=0.0, stddev=1.0):
    def init(shape):
        a = _S.rs.randn(*shape)
        bad = np.abs(a) > 2
        while bad.any():
            a[bad] = _S.rs.randn(int(bad.sum()))
            bad = np.abs(a) > 2
        return torch.tensor(mean + a * stddev, dtype=torch.float32)
    return init


def constant_initializer(value=0, dtype=None):
    return lambda shape: torch.full(tuple(shape), float(value), dtype=torch.float32)


initializers = types.SimpleNamespace(constant=constant_initializer)


def get_variable(name, shape=None, dtype=None, initializer=None):
    return _register(name, initializer([int(s) for s in shape]), True)


def Variable(initial_value, name=None, dtype=None, trainable=True):
    v = _t(initial_value)
    if not trainable:                                              # global_step
        v = torch.as_tensor(np.asarray(initial_value))
    return _register(name, v, trainable)


def get_collection(key, scope=None):
    return []


GraphKeys = types.SimpleNamespace(UPDATE_OPS='update_ops', LOCAL_VARIABLES='local_variables')


def global_variables_initializer():
    return None


def local_variables_initializer():
    return None


def trainable_variables():
    return list(_S.variables.values())


class _Saver(object):
    def __init__(self, *a, **k):
        pass


train = types.SimpleNamespace(latest_checkpoint=lambda path: None, Saver=_Saver)
summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None, merge_all=lambda: None,
                                FileWriter=lambda *a, **k: None)
metrics = types.SimpleNamespace(mean_relative_error=lambda *a, **k: (None, None),
                                average_precision_at_k=lambda *a, **k: (None, None),
                                accuracy=lambda *a, **k: (None, None))
logging = types.SimpleNamespace(set_verbosity=lambda *a: None, WARN=0, DEBUG=0)


class GradientRecorder(object):
    """Stands where a tf.train.*Optimizer would: `compute_gradients` = d loss / d trainable variable (autograd);
    `apply_gradients` does nothing (the update rules of TF's optimizers are TF internals)."""

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate

    def compute_gradients(self, loss):
        vs = list(_S.variables.items())
        gs = torch.autograd.grad(loss, [v for _, v in vs], allow_unused=True)
        out = []
        for (name, v), g in zip(vs, gs):
            _S.grads[name] = g
            out.append((g, v))
        return out

    def apply_gradients(self, grads, global_step=None):
        return None


# ---------------------------------------------------------------- tensor ops
class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = np.asarray(indices), np.asarray(values), tuple(dense_shape)


def sparse_reorder(sp):
    order = np.lexsort((sp.indices[:, 1], sp.indices[:, 0]))      # canonical row-major order
    return SparseTensor(sp.indices[order], sp.values[order], sp.dense_shape)


def cast(x, dtype):
    if isinstance(x, SparseTensor):
        return SparseTensor(x.indices, x.values.astype(np.float64 if dtype == torch.float64 else np.float32), x.dense_shape)
    return x.to(dtype)


def sparse_tensor_dense_matmul(sp, x):
    idx = torch.from_numpy(np.ascontiguousarray(sp.indices.T).astype(np.int64))
    A = torch.sparse_coo_tensor(idx, torch.from_numpy(np.ascontiguousarray(sp.values)).to(x.dtype), sp.dense_shape)
    return torch.sparse.mm(A.coalesce().to_sparse_csr(), x)


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim())))


def reshape(x, shape):
    return x.reshape([int(s) for s in shape])


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def squeeze(x, axis=None):
    if axis is None:
        return x.squeeze()
    for a in sorted(axis, reverse=True):
        x = x.squeeze(a)
    return x


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def matmul(a, b):
    return a @ b


def multiply(a, b):
    return a * b


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def constant(value, dtype=None):
    return _t(value, dtype)


def zeros(shape, dtype=None):
    return torch.zeros([int(s) for s in shape], dtype=torch.float32)


def ones(shape, dtype=None):
    return torch.ones([int(s) for s in shape], dtype=torch.float32)


def zeros_like(x):
    return torch.zeros_like(x)


def linspace(start, stop, num, name=None):
    return torch.from_numpy(np.linspace(np.float32(start), np.float32(stop), int(num), dtype=np.float32))


def pad(x, paddings, mode='CONSTANT', constant_values=0):
    assert mode == 'CONSTANT'
    flat = []
    for lo, hi in reversed(np.asarray(paddings).tolist()):
        flat += [int(lo), int(hi)]
    return F.pad(x, flat, value=float(constant_values))


def abs(x):                                                        # noqa: A001
    return x.abs()


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(dim=axis)


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_max(x, axis=None):
    return x.max() if axis is None else x.max(dim=axis).values


def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


def to_int64(x):
    return x.to(torch.int64)


def one_hot(x, depth):
    return F.one_hot(x.to(torch.int64), depth).to(torch.float32)


def argmax(x, axis=None):
    return x.argmax(dim=axis)


def cond(pred, true_fn, false_fn):
    return true_fn() if pred else false_fn()


def is_nan(x):
    return torch.isnan(x)


def where(c, a, b):
    return torch.where(c, a, b)


def clip_by_value(x, lo, hi):
    return x.clamp(lo, hi)


# ---------------------------------------------------------------- tf.nn / tf.layers / tf.losses
def _pool_windows(x, ksize, strides, padding, fill):
    """NHWC, window and stride `p` over H only (the reference's HEALPix pooling, models.py:1141,1160)."""
    assert padding == 'SAME' and list(ksize) == list(strides) and ksize[0] == ksize[2] == ksize[3] == 1
    p = int(ksize[1])
    N, H, Wd, C = x.shape
    Hp = -(-H // p) * p
    if Hp != H:
        x = F.pad(x, [0, 0, 0, 0, 0, Hp - H], value=fill)
    return x.reshape(N, Hp // p, p, Wd, C), H, p


def _pool2d(x, ksize, strides, padding, kind):
    """NHWC, square window s x s with stride s over H and W, sizes divisible by s (the reference's equiangular pooling,
    models.py:1131-1136, 1150-1155: ksize=[1, p**0.5, p**0.5, 1])."""
    assert padding == 'SAME' and list(ksize) == list(strides) and ksize[0] == ksize[3] == 1 and ksize[1] == ksize[2]
    s = int(ksize[1])
    N, H, Wd, C = x.shape
    assert H % s == 0 and Wd % s == 0
    w = x.reshape(N, H // s, s, Wd // s, s, C)
    return w.amax(dim=(2, 4)) if kind == 'max' else w.mean(dim=(2, 4))


def _max_pool(x, ksize, strides, padding):
    if ksize[2] != 1:
        return _pool2d(x, ksize, strides, padding, 'max')
    w, _, _ = _pool_windows(x, ksize, strides, padding, float('-inf'))
    return w.max(dim=2).values


def _avg_pool(x, ksize, strides, padding):
    if ksize[2] != 1:
        return _pool2d(x, ksize, strides, padding, 'avg')
    w, H, p = _pool_windows(x, ksize, strides, padding, 0.0)
    n = torch.full((w.shape[1],), float(p))
    if H % p:
        n[-1] = H % p                                              # SAME padding: padded cells are not counted
    return w.sum(dim=2) / n.reshape(1, -1, 1, 1)


def _moments(x, axes):
    axes = [axes] if isinstance(axes, int) else list(axes)
    mean = x.mean(dim=axes)
    var = ((x - x.mean(dim=axes, keepdim=True)) ** 2).mean(dim=axes)
    return mean, var


def _dropout(x, keep_prob):
    if keep_prob == 1.0:
        return x
    mask = torch.from_numpy((_S.rs.uniform(size=tuple(x.shape)) < keep_prob).astype(np.float32))
    _S.dropout_masks.append(mask)
    return x * mask / keep_prob


def _sparse_ce(logits=None, labels=None):
    C = logits.shape[-1]
    return F.cross_entropy(logits.reshape(-1, C), labels.reshape(-1).to(torch.int64), reduction='none').reshape(labels.shape)


nn = types.ModuleType('tensorflow.nn')
nn.relu = torch.relu
nn.elu = F.elu
nn.leaky_relu = lambda x, alpha=0.2: F.leaky_relu(x, alpha)
nn.sigmoid = torch.sigmoid
nn.tanh = torch.tanh
nn.softplus = F.softplus
nn.softmax = lambda x: torch.softmax(x, dim=-1)
nn.max_pool = _max_pool
nn.avg_pool = _avg_pool
nn.moments = _moments
nn.dropout = _dropout
nn.l2_loss = lambda t: (t * t).sum() / 2
nn.l2_normalize = lambda x, axis=-1: F.normalize(x, dim=axis)
nn.sparse_softmax_cross_entropy_with_logits = _sparse_ce
nn.bias_add = lambda x, b: x + b


def _batch_normalization(x, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False):
    assert axis == -1 and not center and not scale
    Fdim = x.shape[-1]
    with variable_scope('batch_normalization'):
        mm = _register('moving_mean', torch.zeros(Fdim), False)
        mv = _register('moving_variance', torch.ones(Fdim), False)
    if training:
        mean, var = _moments(x, list(range(x.dim() - 1)))
        _S.updates[mm.op.name] = (mm * momentum + mean.detach() * (1 - momentum))
        _S.updates[mv.op.name] = (mv * momentum + var.detach() * (1 - momentum))
    else:
        mean, var = mm, mv
    return (x - mean) * torch.rsqrt(var + epsilon)


layers = types.SimpleNamespace(batch_normalization=_batch_normalization)
losses = types.SimpleNamespace(mean_squared_error=lambda labels, predictions: ((predictions - labels.to(predictions.dtype)) ** 2).mean())


def install():
    """Register this module as `tensorflow` (and the sub-modules models.py imports, models.py:16-19)."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    sys.modules['tensorflow.nn'] = nn
    py = types.ModuleType('tensorflow.python')
    py.debug = types.ModuleType('tensorflow.python.debug')
    sys.modules['tensorflow.python'] = py
    sys.modules['tensorflow.python.debug'] = py.debug
    names = ['tensorflow.contrib', 'tensorflow.contrib.losses', 'tensorflow.contrib.losses.python',
             'tensorflow.contrib.losses.python.metric_learning']
    for n in names:
        sys.modules[n] = types.ModuleType(n)
    sys.modules[names[-1]].triplet_semihard_loss = None
    # tensorflow.keras.backend.repeat_elements (equiangular unpooling, models.py:1349-1355)
    keras = types.ModuleType('tensorflow.keras')
    keras.backend = types.ModuleType('tensorflow.keras.backend')
    keras.backend.repeat_elements = lambda t, rep, axis: torch.repeat_interleave(t, int(rep), dim=axis)
    me.keras = keras
    sys.modules['tensorflow.keras'] = keras
    sys.modules['tensorflow.keras.backend'] = keras.backend
    return me
