import torch, time
x = torch.empty(16*262144, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1, 4, 16):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for c in range(n):
            k = x.numel() // n
            d[c*k:(c+1)*k].copy_(x[c*k:(c+1)*k], non_blocking=True)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('chunks', n, 'H2D 16.8MB: %.3f ms  %.1f GB/s' % (ms, x.numel()*4/ms/1e6))
