"""Where do two runs of the cosmo training step differ?  Same weights and batch, fused tile recurrence on / off: per stage the
saved forward tensors (input, basis planes, y) and every gradient, max difference relative to the largest magnitude."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import models, ops, optimizers, utils  # noqa: E402

torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = models.deepsphere(
    nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
    F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
    batch_size=B, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False,
    scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_diag')
x, labels = bench.synthetic_batch(B, 7)
xd, ld = model._to_device(x), model._labels_to_device(labels)


def run(tile):
    ops.set_tile(tile)
    ops.fill(model._P.g, 0.0)
    model._loss.zero_()
    logits, _ = model._forward(xd, training=True, save=True)
    saved = []
    for st in model._encoder:
        saved.append([t.clone() if isinstance(t, torch.Tensor) else t for t in st.saved])
    d = model._loss_fwd_bwd(logits, ld, True)
    model._backward(d)
    torch.cuda.synchronize()
    return logits.clone(), saved, {k: v.clone() for k, v in model._P.gviews.items()}


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


a = run(True)
b = run(False)
c = run(False)
print('logits tile vs steps {:.1e}   steps vs steps {:.1e}'.format(rel(a[0], b[0]), rel(c[0], b[0])))
for i, (sa, sb) in enumerate(zip(a[1], b[1])):
    parts = []
    for j, (ta, tb) in enumerate(zip(sa, sb)):
        if isinstance(ta, torch.Tensor) and isinstance(tb, torch.Tensor) and ta.shape == tb.shape and ta.dtype.is_floating_point:
            parts.append('[{}] {:.1e}'.format(j, rel(ta, tb)))
    print('stage', i + 1, 'saved tensors:', ' '.join(parts))
for k in a[2]:
    print('grad {:16s} tile vs steps {:.1e}   steps vs steps {:.1e}'.format(k, rel(a[2][k], b[2][k]), rel(c[2][k], b[2][k])))

# ---- weight gradients: kernel accuracy against the float64 product of ITS OWN inputs, and sensitivity to the inputs
rec = {}
orig = ops.contract_bwd_weight


def spy(x, basis, dy, dW, K):
    before = dW.clone()
    orig(x, basis, dy, dW, K)
    rec.setdefault(tag, []).append((x.clone(), None if basis is None else basis.clone(), dy.clone(), (dW - before).clone(), K))


ops.contract_bwd_weight = spy
import deepsphere_tf1_b200.models as M_  # noqa: E402
for tag, tile in (('tile', True), ('steps', False)):
    run(tile)


def dw64(x, basis, dy, K):
    Bq, Mq, Fin = x.shape
    R = Bq * Mq
    planes = [x.reshape(R, Fin).double()] + ([basis[k].reshape(R, Fin).double() for k in range(K - 1)] if basis is not None else [])
    d = dy.reshape(R, -1).double()
    return torch.stack([p.t() @ d for p in planes], dim=1).reshape(Fin * K, -1)


for i, (ta, tb) in enumerate(zip(rec['tile'], rec['steps'])):
    xa, ba, dya, dWa, K = ta
    xb, bb, dyb, dWb, _ = tb
    ra, rb = dw64(xa, ba, dya, K), dw64(xb, bb, dyb, K)
    cond = float((xa.abs().reshape(-1, xa.shape[-1]).double().t() @ dya.abs().reshape(-1, dya.shape[-1]).double()).max() / ra.abs().max())
    print('wgrad call {} (Fin {} K {}): dy tile vs steps {:.1e} | kernel vs float64 of its own inputs: tile {:.1e} steps {:.1e} | '
          'float64(tile inputs) vs float64(steps inputs) {:.1e} | sum|x||dy| / |dW| = {:.0f}'.format(
              i, xa.shape[-1], K, rel(dya, dyb), rel(dWa, ra), rel(dWb, rb), rel(ra, rb), cond))
