"""torch.autograd / torch.library face of the hot-path kernels ("torch ops ... autograd via
torch.autograd.Function").  The models in models.py run their own static layer program and never touch autograd; this module is
for users who want the B200 ChebConv inside an ordinary PyTorch model:

    conv = functional.ChebConv(L, Fin, Fout, K)          # nn.Module, weights in the reference layout [Fin*K, Fout]
    y = conv(x)                                          # [B, M, Fin] -> [B, M, Fout], differentiable

    y = functional.chebconv(x, graph, W, K)              # the same as a function (graph = graph.DeviceGraph)
    torch.ops.dsb200.spmm_ell(col, val, x, y, z, a, b, c) # raw registered operators on ELL tensors (no autograd)

`chebconv` is cgcnn.chebyshev5 (reference deepsphere/models.py:1039-1078); its backward is the hand-written adjoint the
models use: dW = Xcat^T dY (weight-gradient contraction), dX_k = dY W_k^T (data-gradient contraction), dX = adjoint recurrence
with L^T.  PyTorch supplies tensors, the stream and the autograd tape; every number is computed by libdsb200.
"""
import torch

from . import ops
from ._lib import call
from .graph import DeviceGraph


class _ChebConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, graph, K, monomial):
        x = x.contiguous()
        W = W.contiguous()
        basis = ops.cheb_basis(graph, x, K, monomial)
        y = ops.contract_fwd(x, basis, W, K, W.shape[1], None)
        ctx.save_for_backward(x, basis if basis is not None else x.new_empty(0), W)
        ctx.graph, ctx.K, ctx.monomial = graph, K, monomial
        return y

    @staticmethod
    def backward(ctx, dy):
        x, basis, W = ctx.saved_tensors
        K = ctx.K
        basis = basis if K > 1 else None
        dy = dy.contiguous()
        dW = dx = None
        if ctx.needs_input_grad[1]:
            dW = torch.zeros_like(W)
            ops.contract_bwd_weight(x, basis, dy, dW, K)
        if ctx.needs_input_grad[0]:
            G = ops.contract_bwd_data(dy, W, K, x.shape[2])
            dx = ops.cheb_basis_bwd(ctx.graph, G, ctx.monomial) if K > 1 else G[0]
        return dx, dW, None, None, None


def chebconv(x, graph, W, K, monomial=False):
    """y = [T_0(L) x, ..., T_{K-1}(L) x] . W   (x [B, M, Fin] fp32 CUDA, W [Fin*K, Fout] with row fin*K + k)."""
    if x.dim() != 3 or W.dim() != 2 or W.shape[0] != x.shape[2] * K:
        raise ValueError('chebconv: x [B, M, Fin] and W [Fin*K, Fout] expected')
    if x.shape[1] != graph.M:
        raise ValueError('chebconv: {} vertices, the graph has {}'.format(x.shape[1], graph.M))
    return _ChebConvFn.apply(x, W, graph, int(K), bool(monomial))


class ChebConv(torch.nn.Module):
    """Graph convolution with Chebyshev filters on a fixed Laplacian (scipy sparse, already rescaled to [-1, 1])."""

    def __init__(self, L, Fin, Fout, K, device=None, nested=None):
        super().__init__()
        device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.graph = L if isinstance(L, DeviceGraph) else DeviceGraph(L, device, nested=nested)
        self.Fin, self.Fout, self.K = Fin, Fout, K
        std = 1.0 / (Fin * (K + 0.5) / 2) ** 0.5                               # models.py:1080-1083
        w = torch.nn.init.trunc_normal_(torch.empty(Fin * K, Fout), std=std, a=-2 * std, b=2 * std)
        self.weight = torch.nn.Parameter(w.to(device))

    def forward(self, x):
        return chebconv(x, self.graph, self.weight, self.K)


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, kind):
        x = x.contiguous()
        out = ops.bn_act_pool_fwd(x, None, None, None, p, kind, 'identity')
        ctx.save_for_backward(x)
        ctx.p, ctx.kind = p, kind
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        gz, _ = ops.bn_act_pool_bwd(x, None, None, None, dout.contiguous(), ctx.p, ctx.kind, 'identity')
        return gz, None, None


def pool(x, p, kind='max'):
    """HEALPix pooling over p consecutive (nested) vertices (cgcnn.pool_max / pool_average, models.py:1128-1163)."""
    return x if p == 1 else _PoolFn.apply(x, int(p), kind)


# ---------------------------------------------------------------------------------------------- torch.library registration
_LIB = None


def register_ops():
    """Register the raw kernels as `torch.ops.dsb200.*` (idempotent).  ELL graphs as tensors: col int32 [M, w], val fp32 [M, w]."""
    global _LIB
    if _LIB is not None:
        return _LIB
    lib = torch.library.Library('dsb200', 'DEF')
    lib.define('spmm_ell(Tensor col, Tensor val, Tensor x, Tensor? y, Tensor? z, float alpha, float beta, float gamma) -> Tensor')
    lib.define('cheb_basis_ell(Tensor col, Tensor val, Tensor x, int K) -> Tensor')
    lib.define('cheb_contract_fwd(Tensor x0, Tensor? xk, Tensor W, int K) -> Tensor')
    lib.define('cheb_contract_bwd_data(Tensor dY, Tensor W, int K, int Fin) -> Tensor')
    lib.define('cheb_contract_bwd_weight(Tensor x0, Tensor? xk, Tensor dY, int K) -> Tensor')

    def spmm_ell(col, val, x, y, z, alpha, beta, gamma):
        B, M, F = x.shape
        out = torch.empty_like(x)
        call('spmm_ell', col, val, col.shape[1], B, M, F, x, y, z, out, alpha, beta, gamma)
        return out

    def cheb_basis_ell(col, val, x, K):
        B, M, F = x.shape
        basis = x.new_empty((max(K - 1, 0), B, M, F))
        if K > 1:
            call('cheb_basis', None, col, val, col.shape[1], B, M, F, K, x, basis)
        return basis

    def contract_fwd(x0, xk, W, K):
        return ops.contract_fwd(x0, xk, W, K, W.shape[1], None)

    def contract_bwd_data(dY, W, K, Fin):
        return ops.contract_bwd_data(dY, W, K, Fin)

    def contract_bwd_weight(x0, xk, dY, K):
        dW = x0.new_zeros((x0.shape[2] * K, dY.shape[2]))
        ops.contract_bwd_weight(x0, xk, dY, dW, K)
        return dW

    lib.impl('spmm_ell', spmm_ell, 'CUDA')
    lib.impl('cheb_basis_ell', cheb_basis_ell, 'CUDA')
    lib.impl('cheb_contract_fwd', contract_fwd, 'CUDA')
    lib.impl('cheb_contract_bwd_data', contract_bwd_data, 'CUDA')
    lib.impl('cheb_contract_bwd_weight', contract_bwd_weight, 'CUDA')
    _LIB = lib
    return lib
