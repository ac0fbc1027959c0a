import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from deepsphere_tf1_b200 import ops
dev = torch.device('cuda', 0)
MODE = sys.argv[1] if len(sys.argv) > 1 else 'normal'
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())
for (B, M, Fin, Fout, K) in [(16, 65536, 16, 32, 5), (16, 16384, 32, 64, 5), (16, 4096, 64, 64, 5), (4, 4096, 64, 64, 5), (16, 1024, 64, 64, 5), (8, 10242, 48, 32, 4), (8, 10242, 16, 32, 4), (8, 10242, 32, 32, 4), (8, 10242, 32, 3, 1), (8, 2562, 96, 64, 4), (8, 2562, 64, 128, 4), (8, 642, 192, 128, 4), (8, 162, 384, 256, 4), (8, 42, 768, 512, 4), (8, 12, 1024, 512, 4), (8, 12, 512, 512, 4)]:
    R = B * M
    rs = np.random.RandomState(Fin)
    planes = torch.from_numpy(rs.standard_normal((K, B, M, Fin)).astype(np.float32))
    if MODE == 'relu':          # what the layers really see: non-negative activations with a large mean, zero-mean gradients
        planes = planes.abs() + 1.0
    W = torch.from_numpy((rs.randn(Fin * K, Fout) / np.sqrt(Fin * K)).astype(np.float32))
    dy = torch.from_numpy(rs.standard_normal((B, M, Fout)).astype(np.float32))
    if MODE == 'relu':
        dy = dy - dy.mean(dim=(0, 1), keepdim=True)
    pl = planes.reshape(K, R, Fin).double().to(dev)
    W3 = W.double().reshape(Fin, K, Fout).to(dev)
    dy2 = dy.reshape(R, Fout).double().to(dev)
    y_ref = sum(pl[k] @ W3[:, k, :] for k in range(K))
    dW_ref = torch.stack([pl[k].t() @ dy2 for k in range(K)], dim=1).reshape(Fin * K, Fout)
    dX_ref = torch.stack([dy2 @ W3[:, k, :].t() for k in range(K)])
    x0 = planes[0].contiguous().to(dev); xk = planes[1:].contiguous().to(dev) if K > 1 else None
    out = []
    for tc in (True, False):
        prev = ops.set_tensor_cores(tc)
        y = ops.contract_fwd(x0, xk, W.to(dev), K, Fout, None)
        G = ops.contract_bwd_data(dy.to(dev), W.to(dev), K, Fin)
        dW = torch.zeros(Fin * K, Fout, device=dev)
        ops.contract_bwd_weight(x0, xk, dy.to(dev), dW, K)
        ops.set_tensor_cores(prev)
        out.append('tc=%d fwd %.1e dX %.1e dW %.1e' % (tc, rel(y.reshape(R, Fout), y_ref), rel(G.reshape(K, R, Fin), dX_ref), rel(dW, dW_ref)))
    # fp32 torch on GPU for scale
    y32 = sum(pl[k].float() @ W3[:, k, :].float() for k in range(K))
    torch.backends.cuda.matmul.allow_tf32 = False
    dW32 = torch.stack([pl[k].float().t() @ dy2.float() for k in range(K)], dim=1).reshape(Fin * K, Fout)
    print((B, M, Fin, Fout, K), ' | '.join(out), '| torch fp32: fwd %.1e dW %.1e' % (rel(y32, y_ref), rel(dW32, dW_ref)), flush=True)
