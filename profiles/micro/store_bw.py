"""Development aid: aggregate store bandwidth (torch fill kernels) for L2-resident and DRAM-sized buffers."""
import torch
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for mb in (8, 16, 32, 64, 256, 2048):
    x = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    y = torch.empty_like(x)
    s = t(lambda: x.zero_())
    c = t(lambda: y.copy_(x))
    print(f"{mb:5d} MB: fill {mb / 1024 / s:7.2f} GB/ms-> {mb / 1e3 * 1.048576 / s / 1e3:6.2f} TB/s write | copy {2 * mb * 1.048576e6 / c / 1e12:6.2f} TB/s (r+w)")
