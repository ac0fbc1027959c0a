"""torchrun worker: the one-shot peer-memory all-reduce (csrc/peer_allreduce.cu) against NCCL on random fp64 vectors --
eager calls of many sizes, then thousands of replays of a captured CUDA graph holding several calls."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsphere_tf1_b200.peer import PeerAllReduce  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank))))
pa = PeerAllReduce()
assert pa.enabled, 'peer all-reduce could not be set up'
g = torch.Generator(device='cuda')
g.manual_seed(100 + rank)
for it in range(300):
    n = [1, 2, 30, 64, 112, 128, 256, 1000, 1024][it % 9]
    a = torch.randn(n, dtype=torch.float64, device='cuda', generator=g)
    b = a.clone()
    assert pa.allreduce(a)
    dist.all_reduce(b)
    assert torch.allclose(a, b, rtol=1e-14, atol=1e-14), (it, n, (a - b).abs().max().item())
assert not pa.allreduce(torch.zeros(2000, dtype=torch.float64, device='cuda'))          # too large: caller uses NCCL
assert not pa.allreduce(torch.zeros(8, dtype=torch.float32, device='cuda'))
# graph replay: three calls per graph, the sequence counter advances on the device
xs = [torch.zeros(n, dtype=torch.float64, device='cuda') for n in (64, 128, 30)]
src = [torch.randn(n, dtype=torch.float64, device='cuda', generator=g) for n in (64, 128, 30)]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for x, s in zip(xs, src):
        x.copy_(s)
        pa.allreduce(x)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for x, s in zip(xs, src):
        x.copy_(s)
        pa.allreduce(x)
ref = [s.clone() for s in src]
for r in ref:
    dist.all_reduce(r)
for it in range(3000):
    graph.replay()
torch.cuda.synchronize()
for x, r in zip(xs, ref):
    assert torch.allclose(x, r, rtol=1e-14, atol=1e-14)
# latency, eager
for name, fn in (('peer', lambda t: pa.allreduce(t)), ('nccl', lambda t: dist.all_reduce(t))):
    t = torch.ones(128, dtype=torch.float64, device='cuda')
    for _ in range(10):
        fn(t)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn(t)
    e1.record()
    e1.synchronize()
    if rank == 0:
        print('{}: {:.1f} us per all-reduce of 128 doubles (eager, {} ranks)'.format(name, e0.elapsed_time(e1) * 5, world))
dist.barrier()
if rank == 0:
    print('PEER ALLREDUCE OK on {} GPUs'.format(world))
sys.stdout.flush()
os._exit(0)
