#!/usr/bin/env python
"""Benchmark of the DeepSphere ChebConv training hot path on B200 (BASELINE.json metric:
"train samples/s, cosmo Nside=1024 K=5 at 1/2/4/8 B200; ChebConv HBM GB/s").

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      the restated reference on the host CPU

A step = one full training step (forward + loss + backward + SyncBN / gradient all-reduce +
TF1-Adam update + batch-norm moving averages) of the cosmo model of
experiments/cosmo/experiment_deepsphere.py:39-47,94-98 at order 2: HEALPix Nside=1024 patches of
262,144 px, nsides [1024,512,256,128,64,32,32], F=[16,32,64,64,64,2], K=5, BN, max-pool 4,
statistics='mean', batch 16 per GPU, synthetic maps ~ N(0, 1+3.5^2), fp32.

Prints ONE JSON line (rank 0).  `value` = samples/s with the batch resident in HBM; `e2e` = the
same through the public API (`model.train_step(host batch, next_batch=...)` + `model.loss_handle()`, the loop fit()
runs: pinned host -> device copy of a batch and its labels on a copy stream overlapping the step's kernels,
device -> host read of every step's loss, one step behind); `roofline` = the ChebConv
recurrence SpMM step timed alone with CUDA events (back to back over operand sets >> L2) against the
measured HBM copy bandwidth; `cpu_baseline` = the oracle (restated reference) timed on this box's
host cores; `parity` = one step of this very configuration against the float64 oracle (fp32 path, and
`parity.bf16` for the BF16 mode; the run exits 3 if it fails); `bf16_mode` = the same measurement with
dtype='bfloat16' (a second line of evidence: the headline `dtype` stays f32).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDES = [1024, 512, 256, 128, 64, 32, 32]
ORDER = 2
F = [16, 32, 64, 64, 64, 2]
KCHEB = 5
NPIX = 262144
# 1.02 x largest eigenvalue of each level's normalised Laplacian, as produced by
# deepsphere_tf1_b200.utils.estimate_lmax (ARPACK, ~15 s for the six levels); --recompute-lmax
# re-derives them.  Skipping the eigen-solve only shortens start-up; it is outside the timed region.
LMAX = [1.5169168710708618, 1.5167698860168457, 1.5166294574737549, 1.5161875486373901, 1.5148117542266846,
        1.511078119277954]
WORKLOAD = 'cosmo Nside=1024 order-2 patches (262,144 px), K=5, F=[16,32,64,64,64,2], statistics=mean'


def build_laplacians(recompute=False):
    from deepsphere_tf1_b200 import utils
    idx = utils.nside2indexes(NSIDES, ORDER)
    return utils.build_laplacians(NSIDES, indexes=idx, new=False, lmax=None if recompute else LMAX)


def synthetic_batch(B, seed):
    rs = np.random.RandomState(seed)
    x = (rs.standard_normal((B, NPIX, 1)) * np.sqrt(1 + 3.5 ** 2)).astype(np.float32)
    labels = np.random.RandomState(seed + 1).randint(0, 2, size=B).astype(np.int32)
    return x, labels


def bench_config(global_batch):
    """`config` of the JSON line, the same for both arms (--impl b200 / reference)."""
    return {'workload': WORKLOAD, 'global_batch': global_batch,
            'l2': 'no flush between steps: one step streams > 2 GB of activations per GPU at batch 16, >> 126 MB L2'}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference_steps(L, p, B, steps, warmup):
    """The restated reference (oracle/, torch-CPU autograd + TF1 Adam) on the host cores."""
    import warnings
    import torch
    from oracle.model import OracleCgcnn
    from oracle import optim_tf1
    warnings.filterwarnings('ignore')
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    m = OracleCgcnn(L, F=F, K=[KCHEB] * 6, p=p, batch_norm=[True] * 6, M=[], statistics='mean')
    opt = optim_tf1.Adam(epsilon=1e-8)
    x, labels = synthetic_batch(B, 0)
    xt, lt = torch.from_numpy(x), torch.from_numpy(labels)
    for _ in range(warmup):
        m.train_step(xt, lt, opt, 2e-4)
    times = []
    for s in range(steps):
        t = time.perf_counter()
        m.train_step(xt, lt, opt, 2e-4 * 0.999 ** s)
        times.append(time.perf_counter() - t)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    L, p = build_laplacians(args.recompute_lmax)
    B = 16
    # K steps and W warm-ups as asked (a step of 16 samples takes ~1.3 s on the GPU box's 16 cores); capped so that the run
    # always ends within a few minutes
    steps, warm = max(1, min(args.steps, 60)), max(1, min(args.warmup, 5))
    world = max(1, args.gpus)
    times, cores = cpu_reference_steps(L, p, B, steps, warm)
    ms = float(np.mean(times)) * 1e3
    v = B / (ms / 1e3)
    sample = ('{} full training steps of {} samples each (after {} warm-up) of the same workload, torch-CPU oracle, {} threads; '
              'samples/s does not depend on how many such batches make the global batch').format(steps, B, warm, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': 'train samples/s', 'value': v, 'unit': 'samples/s', 'n_gpus': world,
        'steps': steps, 'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(B * world if args.scaling == 'weak' else B),
        'details': {'parallelism': 'host CPU of rank 0, {} threads; no GPU is used'.format(cores)},
        'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def spmm_roofline(model, B, iters=60, nsets=6, bf16=False):
    """The ChebConv recurrence step X_k = 2 L X_{k-1} - X_{k-2} at the layer with the largest
    feature slab (layer 2: M=65,536, Fin=16), timed alone with CUDA events on the launching stream:
    `iters` launches back to back over `nsets` rotating operand sets (6 x 201 MB, ten times the 126 MB
    L2, so no launch finds its operands cached -- "inputs larger than L2"), average per launch.
    Single launches cannot be timed on this box: the event resolution is ~2 us, and a flush kernel that
    REWRITES a buffer leaves an L2 full of dirty lines whose write-back is charged to the timed kernel
    (tools/exp_timing_method.py).  Algorithmic bytes per launch: 3*M*C*4 + 8*nnz."""
    import torch
    from deepsphere_tf1_b200 import ops
    g = model._graphs[1]
    M, Fin = g.M, F[0]
    dev = model.device
    dt, es = (torch.bfloat16, 2) if bf16 else (torch.float32, 4)
    xs = [torch.randn(B, M, Fin, device=dev).to(dt) for _ in range(nsets)]
    ys = [torch.randn(B, M, Fin, device=dev).to(dt) for _ in range(nsets)]
    outs = [torch.empty(B, M, Fin, device=dev, dtype=dt) for _ in range(nsets)]
    bytes_alg = 3 * B * M * Fin * es + 8 * g.nnz
    for i in range(nsets):
        ops.spmm(g, xs[i], ys[i], None, out=outs[i], alpha=2.0, beta=-1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.spmm(g, xs[i % nsets], ys[i % nsets], None, out=outs[i % nsets], alpha=2.0, beta=-1.0)
    e1.record()
    e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    if bf16:
        peak, how = peaks()
        ach = bytes_alg / t / 1e9
        return {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': None,
                'kernel': 'spmm_pipe_kernel<9,4,true,false,true,BF16> (ChebConv recurrence step, layer 2: B={} M={} Fin={}, bfloat16)'.format(B, M, Fin),
                'algorithmic_bytes_per_launch': bytes_alg, 'avg_launch_us': t * 1e6, 'peak_source': how,
                'timing': '{} launches back to back over {} rotating operand sets, CUDA events (the {} MB of the sets exceed L2 '
                          'about five times)'.format(iters, nsets, nsets * 3 * B * M * Fin * es // 1000000)}
    # for comparison, the round's earlier method: single launches, each after a 512 MB buffer has been rewritten
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    singles = []
    for _ in range(10):
        ops.fill(flush, 0.0)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        ops.spmm(g, xs[0], ys[0], None, out=outs[0], alpha=2.0, beta=-1.0)
        s1.record()
        s1.synchronize()
        singles.append(s0.elapsed_time(s1) * 1e3)
    del flush
    pipelined = ops._pipe(g, 'fwd', B, M, Fin) is not None
    peak, how = peaks()
    ach = bytes_alg / t / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('spmm_kernel_dram_bytes_per_launch')
    kname = ('spmm_pipe_kernel<9,4,true,false,true>' if pipelined else 'spmm_rec_kernel<1,9,1>')
    return {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic,
            'kernel': '{} (ChebConv recurrence step, layer 2: B={} M={} Fin={})'.format(kname, B, M, Fin),
            'algorithmic_bytes_per_launch': bytes_alg, 'avg_launch_us': t * 1e6, 'peak_source': how,
            'timing': '{} launches back to back over {} rotating operand sets ({} MB >> L2), CUDA events'.format(
                iters, nsets, nsets * 3 * B * M * Fin * 4 // 1000000),
            'single_launch_after_rewriting_512MB_us': float(np.median(singles))}


def parity_check(L, p, world, rank, dev, tol=1e-4, with_bf16=True):
    """ONE training step of the benchmark's own configuration (cosmo config 2, global batch 16, M = 262,144) on the CUDA
    path and on the CPU oracle from identical weights and inputs: logits, loss and every gradient.  With N ranks the
    global batch of 16 is split 16/N per rank, so that the check also covers SyncBN, the Gram
    all-reduce of the fused input layer and the flat gradient all-reduce; the oracle always sees the 16 samples.
    Returns the dict of the JSON line's `parity` key (rank 0) -- the run fails above `tol`."""
    import torch
    import torch.distributed as dist
    from deepsphere_tf1_b200 import models, ops, optimizers
    Bg = 16
    Bl = Bg // world
    x, labels = synthetic_batch(Bg, 7)
    xs, ls = x[rank * Bl:(rank + 1) * Bl], labels[rank * Bl:(rank + 1) * Bl]

    def cuda_step(dtype):
        m = models.cgcnn(L=L, p=p, F=F, K=[KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
                         batch_size=Bl, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=False, dtype=dtype,
                         scheduler=lambda step: 2e-4, optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8),
                         dir_name='_bench_parity', device=dev)
        xd, ld = m._to_device(xs), m._labels_to_device(ls)
        ops.fill(m._P.g, 0.0)
        m._loss.zero_()
        m._arena.zero_()
        logits, _ = m._forward(xd, training=True, save=True)
        dlogits = m._loss_fwd_bwd(logits, ld, True)
        m._backward(dlogits)
        if world > 1:
            m._allreduce(m._P.g)
            ops.scale(m._loss, 1.0 / world)
            m._allreduce(m._loss)
            gathered = [torch.empty_like(logits) for _ in range(world)]
            dist.all_gather(gathered, logits)
            logits = torch.cat(gathered, dim=0)
        torch.cuda.synchronize()
        return m, logits

    m, logits = cuda_step(np.float32)
    mb = logitsb = None
    if with_bf16:
        mb, logitsb = cuda_step('bfloat16')              # same seed: identical initial weights
    if rank != 0:
        return None
    import warnings
    from oracle.model import OracleCgcnn
    warnings.filterwarnings('ignore')
    torch.set_num_threads(os.cpu_count())
    # the oracle evaluated in float64: at this size (4 M rows per batch-norm layer) a float32 evaluation on the CPU is itself
    # 1e-3..1e-2 away from the exact gradients of layers 1-3 (summation error), which would be the error measured
    o = OracleCgcnn(L, F=F, K=[KCHEB] * 6, p=p, batch_norm=[True] * 6, M=[], statistics='mean', dtype=torch.float64)
    for name, v in o.params.items():
        o.params[name] = m._P.views[name].detach().cpu().double().reshape(v.shape).clone()
    total, ologits, grads, _ = o.loss_and_grads(torch.from_numpy(x).double(), torch.from_numpy(labels), True)

    def rel(a, b, floor=0.0):
        a, b = a.detach().double().cpu().numpy().reshape(-1), b.detach().double().cpu().numpy().reshape(-1)
        return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30))
    gmax = max(float(g.abs().max()) for g in grads.values())
    # what float32 arithmetic can deliver at this size: the same oracle evaluated in float32 on the CPU against its float64
    # evaluation.  The gradients of layers 1-3 sum 4-17 M terms that cancel to ~1e-4 of their magnitude (everything behind a
    # batch-norm layer has zero mean), so ANY float32 evaluation -- TF1's included -- is 1e-3..1e-2 off on them.
    o32 = OracleCgcnn(L, F=F, K=[KCHEB] * 6, p=p, batch_norm=[True] * 6, M=[], statistics='mean')
    for name, v in o32.params.items():
        o32.params[name] = m._P.views[name].detach().cpu().reshape(v.shape).clone()
    _, _, grads32, _ = o32.loss_and_grads(torch.from_numpy(x), torch.from_numpy(labels), True)
    per_var, worst, worst_name = {}, 0.0, ''
    num = den = num32 = 0.0
    for name, g in grads.items():
        e = rel(m._P.gviews[name], g, floor=1e-3 * gmax)
        e32 = rel(grads32[name], g, floor=1e-3 * gmax)
        per_var[name] = {'cuda_vs_f64': e, 'cpu_f32_vs_f64': e32}
        gd = g.double().reshape(-1)
        num += float(((m._P.gviews[name].detach().cpu().double().reshape(-1) - gd) ** 2).sum())
        num32 += float(((grads32[name].double().reshape(-1) - gd) ** 2).sum())
        den += float((gd ** 2).sum())
        if e > worst:
            worst, worst_name = e, name
    grad_l2 = (num / den) ** 0.5
    grad_l2_32 = (num32 / den) ** 0.5
    grads_ok = grad_l2 <= max(tol, 4.0 * grad_l2_32)
    out = {'loss_rel': abs(float(m._loss) - float(total)) / max(abs(float(total)), 1e-30),
           'max_logit_rel': rel(logits, ologits), 'grad_rel_l2': grad_l2,
           'grad_rel_l2_of_cpu_float32_oracle': (num32 / den) ** 0.5, 'max_grad_rel': worst, 'worst_gradient': worst_name,
           'max_grad_rel_of_cpu_float32_oracle': max(v['cpu_f32_vs_f64'] for v in per_var.values()),
           'gradients': per_var,
           'loss': float(m._loss), 'oracle_loss': float(total), 'global_batch': Bg, 'ranks': world, 'tolerance': tol,
           'what': 'one training step of the benchmark configuration (M=262,144, batch 16 split over the ranks): CUDA path '
                   '(fp32) vs CPU oracle evaluated in float64 from identical weights; errors relative to the largest '
                   'expected magnitude of each tensor; grad_rel_l2 = ||g - g_ref|| / ||g_ref|| over the flat gradient of all '
                   'variables.  ok = loss_rel and max_logit_rel within the tolerance, grad_rel_l2 within max(tolerance, 4 x the '
                   'grad_rel_l2 of the float32 CPU evaluation of the same oracle).  Why the gradients cannot be held to 1e-4 at '
                   'this size: the loss is not differentiable where two candidates of a max-pool group or a ReLU input tie, and '
                   'among 17 M activations a few are within float32 rounding of a tie; which one wins differs between ANY two '
                   'float32 evaluations whose forward passes differ by 1e-6 (two kernels of this library, TF1 on CPU vs GPU), and '
                   'one flipped arg-max moves a weight gradient by ~1e-3 (tools/diag_tile_model.py: every weight-gradient kernel '
                   'equals the float64 product of its own inputs to 1e-6 x its condition number).  Per-tensor max-norm errors '
                   'are listed beside those of the float32 CPU oracle'}
    out['ok'] = bool(out['loss_rel'] <= tol and out['max_logit_rel'] <= tol and grads_ok)
    if mb is not None:
        # BF16 mode (dtype='bfloat16': layers 2-6 on bfloat16 activations) against the same float64 oracle.  Forward: 1e-2 (north
        # star).  Gradients: BF16 storage flips arg-max / ReLU decisions, so they are held to the deviation of the CPU oracle
        # evaluated with the same tensors rounded to bfloat16 (oracle.ops.STORE) -- the noise floor of BF16 storage itself.
        oe = OracleCgcnn(L, F=F, K=[KCHEB] * 6, p=p, batch_norm=[True] * 6, M=[], statistics='mean')
        oe.bf16_layers = {2, 3, 4, 5, 6}
        for name, v in oe.params.items():
            oe.params[name] = m._P.views[name].detach().cpu().reshape(v.shape).clone()
        tote, logitse, gradse, _ = oe.loss_and_grads(torch.from_numpy(x), torch.from_numpy(labels), True)
        nb = ne = 0.0
        for name, g in grads.items():
            gd = g.double().reshape(-1)
            nb += float(((mb._P.gviews[name].detach().cpu().double().reshape(-1) - gd) ** 2).sum())
            ne += float(((gradse[name].double().reshape(-1) - gd) ** 2).sum())
        b = {'loss_rel': abs(float(mb._loss) - float(total)) / max(abs(float(total)), 1e-30),
             'max_logit_rel': rel(logitsb, ologits), 'grad_rel_l2': (nb / den) ** 0.5,
             'grad_rel_l2_of_bf16_storage_oracle': (ne / den) ** 0.5,
             'max_logit_rel_of_bf16_storage_oracle': rel(logitse, ologits), 'tolerance': 1e-2,
             'what': 'the same step with dtype=bfloat16 vs the float64 oracle; ok = loss_rel and max_logit_rel within 1e-2, '
                     'grad_rel_l2 within max(1e-2, 3 x that of the CPU oracle with BF16 storage emulation)'}
        b['ok'] = bool(b['loss_rel'] <= 1e-2 and b['max_logit_rel'] <= 1e-2 and
                       b['grad_rel_l2'] <= max(1e-2, 3.0 * b['grad_rel_l2_of_bf16_storage_oracle']))
        out['bf16'] = b
    return out


def e2e_loop(model, x_pin, l_pin, nxt, steps):
    """The inner loop of fit() (models.py: train_step + loss_handle): every step copies a full batch host -> device (on the copy
    stream, overlapping the step) and reads its loss back to the host -- one step behind, i.e. the host waits for the loss of
    step i - 1 after it has enqueued step i, and for the last one before the timed region ends.  `steps` losses are read."""
    pending, last = None, float('nan')
    for _ in range(steps):
        model.train_step(x_pin, l_pin, next_batch=nxt)
        handle = model.loss_handle()
        if pending is not None:
            last = pending.value()
        pending = handle
    if pending is not None:
        last = pending.value()
    return last


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from deepsphere_tf1_b200 import _lib, models, optimizers, utils

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world == 1 and args.gpus > 1:
        raise SystemExit('launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node {0} '
                         '--master-addr 127.0.0.1 bench.py --gpus {0}'.format(args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    if args.parity_only:
        L, p = build_laplacians(args.recompute_lmax)
        par = parity_check(L, p, world, rank, torch.device('cuda', local_rank))
        if rank == 0:
            print(json.dumps({'parity': par}))
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        return
    B_local = 16 if args.scaling == 'weak' else max(1, 16 // world)
    B_global = B_local * world
    model = models.deepsphere(
        nsides=NSIDES, indexes=utils.nside2indexes(NSIDES, ORDER), new=False,
        lmax=None if args.recompute_lmax else LMAX,
        F=F, K=[KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1,
        batch_size=B_local, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=not args.no_graph,
        scheduler=lambda step: optimizers.exponential_decay(2e-4, step, decay_steps=1, decay_rate=0.999),
        optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_bench')
    L, p = model.L, model.p
    dev = model.device
    warm = max(args.warmup, 3)

    x_np, l_np = synthetic_batch(B_local, 100 + rank)
    x_pin = torch.from_numpy(x_np).pin_memory()
    l_pin = torch.from_numpy(l_np).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- e2e: public API, host buffers, loss read back every step
    last_loss = float('nan')
    # `next_batch`: what fit() does -- the host -> device copy of the following step's batch goes to a copy stream as
    # soon as this step's kernels are enqueued, so every timed step contains one full batch copy, overlapped
    nxt = (x_pin, l_pin)
    for _ in range(warm):
        _, loss = model.train_step(x_pin, l_pin, next_batch=nxt)
        float(loss)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last_loss = e2e_loop(model, x_pin, l_pin, nxt, args.steps)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) * 1e-3

    # ---------------- device-resident: the same step with the batch already in HBM
    xd, ld = model._to_device(x_pin), model._labels_to_device(l_pin)
    if model.cuda_graph:
        model._gx.copy_(xd)
        model._gl.copy_(ld)
        step = lambda: model._step_graph_resident()              # noqa: E731
    else:
        step = lambda: model._step_device(xd, ld)                # noqa: E731
    for _ in range(warm):
        model._advance_hyper()
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.2)
    t0 = time.time()
    barrier()
    e0.record()
    for _ in range(args.steps):
        model._advance_hyper()
        step()
    e1.record()
    barrier()
    t1 = time.time()
    t_dev = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop(t0, t1) if sampler else None

    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])

    parity = None if args.no_parity else parity_check(L, p, world, rank, dev, with_bf16=not args.no_bf16)

    # ---------------- BF16 mode: the same step with dtype='bfloat16' (a second line of evidence, never the headline)
    bf16 = None
    if not args.no_bf16:
        mb = models.deepsphere(
            nsides=NSIDES, indexes=utils.nside2indexes(NSIDES, ORDER), new=False, lmax=None if args.recompute_lmax else LMAX,
            F=F, K=[KCHEB] * 6, batch_norm=[True] * 6, M=[], statistics='mean', num_epochs=1, dtype='bfloat16',
            batch_size=B_local, eval_frequency=10 ** 9, seed=2, verbose=False, cuda_graph=not args.no_graph,
            scheduler=lambda step: optimizers.exponential_decay(2e-4, step, decay_steps=1, decay_rate=0.999),
            optimizer=lambda lr: optimizers.AdamOptimizer(lr, epsilon=1e-8), dir_name='_bench_bf16')
        for _ in range(warm):
            _, lossb = mb.train_step(x_pin, l_pin, next_batch=nxt)
            float(lossb)
        barrier()
        e0.record()
        lastb = e2e_loop(mb, x_pin, l_pin, nxt, args.steps)
        e1.record()
        barrier()
        tb_e2e = e0.elapsed_time(e1) * 1e-3
        if mb.cuda_graph:
            mb._gx.copy_(xd)
            mb._gl.copy_(ld)
            stepb = lambda: mb._step_graph_resident()              # noqa: E731
        else:
            stepb = lambda: mb._step_device(xd, ld)                # noqa: E731
        for _ in range(warm):
            mb._advance_hyper()
            stepb()
        barrier()
        e0.record()
        for _ in range(args.steps):
            mb._advance_hyper()
            stepb()
        e1.record()
        barrier()
        tb = torch.tensor([e0.elapsed_time(e1) * 1e-3, tb_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bf16 = {'dtype': 'bf16', 'value': B_global * args.steps / float(tb[0]), 'unit': 'samples/s',
                'ms_per_step': float(tb[0]) / args.steps * 1e3,
                'e2e': {'value': B_global * args.steps / float(tb[1]), 'unit': 'samples/s', 'ms_per_step': float(tb[1]) / args.steps * 1e3},
                'final_loss': lastb,
                'what': "the same training step with cgcnn(dtype='bfloat16'): layers 2-6 store activations, Chebyshev planes and "
                        'activation gradients as bfloat16 (fp32 arithmetic / accumulation, fp32 weights and statistics; tcgen05 '
                        'kind::f16 contractions); parity in parity.bf16'}
        if rank == 0:
            bf16["roofline"] = spmm_roofline(mb, B_local, nsets=12, bf16=True)
        del mb

    # kernels per step, counted by the library itself on one eager (non-graph) step
    c0 = _lib.kernel_launches()
    model._advance_hyper()
    model._step_device(xd, ld)
    torch.cuda.synchronize()
    per_step = _lib.kernel_launches() - c0

    exit_code = 0
    if rank == 0:
        roof = spmm_roofline(model, B_local)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_reference_steps(L, p, 16, 3, 1)
            cpu = {'value': 16 / float(np.mean(times)), 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                   'sample': '3 full training steps of batch 16 (after 1 warm-up) of the same workload, '
                             'torch-CPU oracle, {} threads'.format(cores)}
        out = {
            'metric': 'train samples/s', 'value': B_global * args.steps / t_dev, 'unit': 'samples/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': warm,
            'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': bench_config(B_global),
            'details': {'local_batch': B_local, 'cuda_graph': bool(model.cuda_graph),
                        'parallelism': 'dp{} (SyncBN + one flat NCCL gradient all-reduce)'.format(world)},
            'e2e': {'value': B_global * args.steps / t_e2e, 'unit': 'samples/s',
                    'h2d_bytes_per_step': int(x_pin.numel() * 4 + l_pin.numel() * 4 + 32), 'd2h_bytes_per_step': 4,
                    'ms_per_step': t_e2e / args.steps * 1e3,
                    'api': 'the loop of fit(): train_step(pinned host batch, next_batch=...) copies the following batch on a copy '
                           'stream while the step runs; the loss of EVERY step is read back to the host through '
                           'model.loss_handle(), one step behind (the host waits for the loss of step i - 1 after enqueuing step '
                           'i, and for the last one inside the timed region)'},
            'gpu_launches': int(per_step * args.steps),
            'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'parity': parity, 'final_loss': last_loss,
            'bf16_mode': bf16,
        }
        print(json.dumps(out))
        if parity is not None and not (parity['ok'] and parity.get('bf16', {'ok': True})['ok']):
            sys.stderr.write('PARITY FAILED: {}\n'.format(json.dumps(parity)))
            exit_code = 3
    if world == 1 and exit_code:
        sys.exit(exit_code)
    if world > 1:
        # Captured CUDA graphs hold NCCL work; tearing the process group down under them can block
        # (seen with NCCL 2.28 / torch 2.11).  Release the graph, meet at a barrier, flush and leave.
        model._graph_exec = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(exit_code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--no-graph', action='store_true', help='issue every launch from Python instead of one CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parity-only', action='store_true', help='only the full-size CUDA-vs-oracle check (no timing)')
    ap.add_argument('--no-parity', action='store_true', help='skip the full-size CUDA-vs-oracle check of one training step')
    ap.add_argument('--no-bf16', action='store_true', help="skip the second measurement with dtype='bfloat16' (key bf16_mode)")
    ap.add_argument('--recompute-lmax', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
