import torch
x = torch.empty(354*1024*1024//4, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
tf = t(lambda: x.fill_(1.0)); tc = t(lambda: y.copy_(x))
nb = x.numel() * 4
print(f"fill 354MiB: {tf:.1f} us = {nb/1e9/(tf/1e6):.0f} GB/s written; copy: {tc:.1f} us = {2*nb/1e9/(tc/1e6):.0f} GB/s read+write")
