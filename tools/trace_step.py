"""Per-kernel time of ONE eager cosmo training step in situ (torch.profiler / CUPTI: kernels run back to back with the
caches as the step leaves them -- unlike an ncu launch list, which replays every kernel cold and alone).  Attribution only;
bench.py's numbers are never taken under a profiler."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepsphere_tf1_b200 import models, optimizers, utils  # noqa: E402

torch.cuda.set_device(0)
CFG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if CFG != 2:
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench_configs
    c = bench_configs.config(CFG)
    model = bench_configs.build_model(c, c['B'], False, '_trace')
    x, labels = c['x'], c['labels']
else:
  model = models.deepsphere(
    nsides=bench.NSIDES, indexes=utils.nside2indexes(bench.NSIDES, bench.ORDER), new=False, lmax=bench.LMAX,
    F=bench.F, K=[bench.KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
    batch_size=16, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False,
    dtype='bfloat16' if os.environ.get('DSB200_TRACE_DTYPE', 'f32') == 'bf16' else 'float32',
    scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_trace')
if CFG == 2:
    x, labels = bench.synthetic_batch(16, 0)
xd, ld = model._to_device(x), model._labels_to_device(labels)
for _ in range(3):
    model._advance_hyper()
    model._step_device(xd, ld)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        model._advance_hyper()
        model._step_device(xd, ld)
    torch.cuda.synchronize()
tot = collections.OrderedDict()
order = []
for ev in prof.events():
    if ev.device_type.name != 'CUDA' or not ev.name:
        continue
    name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '').replace('dsb200::', '').split('(')[0]
    d = tot.setdefault(name, [0.0, 0])
    d[0] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    d[1] += 1
total = sum(v[0] for v in tot.values()) / 3
print('{:>10s} {:>5s} {:>6s}  kernel'.format('us/step', 'n', 'share'))
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print('{:10.1f} {:5.1f} {:5.1f}%  {}'.format(us / 3, n / 3, 100 * us / 3 / total, name[:110]))
print('{:10.1f} total kernel time per step'.format(total))
