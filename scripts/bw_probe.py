"""Reference points for the HBM roofline of write-only / read-only kernels (torch library kernels, CUDA events)."""
import torch
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
bufs = [torch.empty(24 * 17 * 512 * 512, device="cuda") for _ in range(3)]   # 428 MB each: rotate to defeat L2
i = [0]
def fill():
    i[0] = (i[0] + 1) % 3; bufs[i[0]].fill_(1.0)
def read():
    i[0] = (i[0] + 1) % 3; return bufs[i[0]].sum()
def copy():
    i[0] = (i[0] + 1) % 3; bufs[i[0]].copy_(bufs[(i[0] + 1) % 3])
nb = bufs[0].numel() * 4
for name, fn, by in (("fill (write-only)", fill, nb), ("sum (read-only)", read, nb), ("copy (read+write)", copy, 2 * nb)):
    ms = timeit(fn); print("%-20s %.3f ms  %.0f GB/s" % (name, ms, by / ms / 1e6))
