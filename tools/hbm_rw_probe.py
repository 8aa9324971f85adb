import torch, time
x = torch.empty(8 * 1024**3, dtype=torch.uint8, device="cuda")
y = torch.empty(4 * 1024**3, dtype=torch.uint8, device="cuda")
for name, fn, nbytes in (("write(fill)", lambda: x.zero_(), x.numel()), ("copy 4GB", lambda: x[:y.numel()].copy_(y), 2 * y.numel()), ("read(sum)", lambda: x.view(torch.int64).sum(), x.numel())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, round(nbytes / ms / 1e6, 1), "GB/s")
