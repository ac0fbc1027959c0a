"""BF16 mode at the model level: a cosmo-style miniature (and optionally the benchmark configuration) in dtype='bfloat16' against
the float32 product and the CPU oracle: logits, loss, every gradient.  Prints, does not assert.
    python tools/diag_bf16_model.py [full]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from deepsphere_tf1_b200 import models, ops, optimizers, utils  # noqa: E402


def build(nsides, order, F, K, dtype, B, seed=0):
    idx = utils.nside2indexes(nsides, order)
    kw = dict(F=F, K=K, batch_norm=[True] * len(F), M=[], statistics='mean')
    return models.deepsphere(nsides=nsides, indexes=idx, new=False, num_epochs=1, batch_size=B, verbose=False,
                             scheduler=lambda step: 1e-3, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                             dir_name='_diag_bf16', dtype=dtype, seed=seed, **kw), kw


def grads(m, x, labels):
    P = m._P
    ops.fill(P.g, 0.0)
    m._loss.zero_()
    logits, _ = m._forward(m._to_device(x), training=True, save=True)
    d = m._loss_fwd_bwd(logits, m._labels_to_device(labels), True)
    m._backward(d)
    torch.cuda.synchronize()
    return logits.float().cpu(), float(m._loss), {n: P.gviews[n].detach().cpu().clone() for n in P.gviews}


def main(full):
    if full:
        nsides, order, F, B = [1024, 512, 256, 128, 64, 32, 32], 2, [16, 32, 64, 64, 64, 2], 16
    else:
        nsides, order, F, B = [64, 32, 16, 8, 8], 2, [16, 32, 64, 2], 4
    K = [5] * len(F)
    m32, kw = build(nsides, order, F, K, np.float32, B)
    mbf, _ = build(nsides, order, F, K, 'bfloat16', B)
    mbf._P.w.copy_(m32._P.w)
    print('bf16 stages:', [(st.scope, st.bf16, st.small) for st in mbf._encoder])
    M0 = m32.L[0].shape[0]
    rs = np.random.RandomState(0)
    x = rs.randn(B, M0, 1).astype(np.float32)
    labels = rs.randint(0, 2, size=B).astype(np.int32)
    l32, loss32, g32 = grads(m32, x, labels)
    lbf, lossbf, gbf = grads(mbf, x, labels)
    print('loss fp32 {:.6f}  bf16 {:.6f}  rel {:.3e}'.format(loss32, lossbf, abs(loss32 - lossbf) / abs(loss32)))
    print('logits max err / max|logit| {:.3e}'.format(float((l32 - lbf).abs().max() / l32.abs().max())))
    num = den = 0.0
    for n in g32:
        a, e = gbf[n].double(), g32[n].double()
        num += float((a - e).pow(2).sum()); den += float(e.pow(2).sum())
        print('  grad {:16s} max err / max|g| {:.3e}   relL2 {:.3e}'.format(n, float((a - e).abs().max() / max(float(e.abs().max()), 1e-30)),
                                                                        float((a - e).norm() / max(float(e.norm()), 1e-30))))
    print('flat gradient relL2 {:.3e}'.format((num / den) ** 0.5))
    # three optimizer steps
    for _ in range(3):
        a = m32.train_step(x, labels); b = mbf.train_step(x, labels)
    print('after 3 steps: loss fp32 {:.6f} bf16 {:.6f}; weights relL2 {:.3e}'.format(float(a[1]), float(b[1]),
          float((m32._P.w - mbf._P.w).norm() / m32._P.w.norm())))
    pa, pb = m32.predict(x), mbf.predict(x)
    print('predict agree:', float((np.asarray(pa) == np.asarray(pb)).mean()))
    if full:
        for name, m in (('fp32', m32), ('bf16', mbf)):
            xd, ld = m._to_device(x), m._labels_to_device(labels)
            for _ in range(3):
                m._step_device(xd, ld)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                m._step_device(xd, ld)
            e1.record(); torch.cuda.synchronize()
            print('{} eager step: {:.3f} ms'.format(name, e0.elapsed_time(e1) / 10))


if __name__ == '__main__':
    main(len(sys.argv) > 1 and sys.argv[1] == 'full')
