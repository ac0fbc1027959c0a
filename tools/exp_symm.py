import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
import torch.distributed._symmetric_memory as symm
group = dist.group.WORLD
for dtype in (torch.float32, torch.float64):
    try:
        t = symm.empty(128, dtype=dtype, device='cuda')
        h = symm.rendezvous(t, group.group_name)
        t.fill_(rank + 1)
        torch.cuda.synchronize(); dist.barrier()
        out = torch.ops.symm_mem.one_shot_all_reduce(t, 'sum', group.group_name)
        torch.cuda.synchronize()
        if rank == 0: print(dtype, 'one_shot ok', out[:2].tolist())
        # timing vs nccl
        for name, fn in (('one_shot', lambda: torch.ops.symm_mem.one_shot_all_reduce(t, 'sum', group.group_name)),
                         ('nccl', lambda: dist.all_reduce(t))):
            for _ in range(5): fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): fn()
            e1.record(); e1.synchronize()
            if rank == 0: print('  ', name, '%.1f us per all-reduce of 128 elements' % (e0.elapsed_time(e1) * 10))
        # graph capture
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g):
                o2 = torch.ops.symm_mem.one_shot_all_reduce(t, 'sum', group.group_name)
        g.replay(); torch.cuda.synchronize()
        if rank == 0: print('   graph capture ok', o2[:2].tolist())
    except Exception as e:
        if rank == 0: print(dtype, 'FAILED:', type(e).__name__, str(e)[:300])
dist.barrier()
os._exit(0)
