"""Host<->device copy rates for the e2e leg's buffers (pinned memory, one stream)."""
import torch
def ev(): return torch.cuda.Event(enable_timing=True)
shapes = dict(f_n=(24, 256, 32, 32), f_o=(24, 256, 32, 32), l_po=(24, 16, 32, 32), logits_lr=(24, 17, 32, 32))
host = {k: torch.randn(*s).pin_memory() for k, s in shapes.items()}
host["labels"] = torch.zeros(24, 512, 512, dtype=torch.uint8).pin_memory()
nbytes = sum(v.numel() * v.element_size() for v in host.values())
dev = {k: torch.empty_like(v, device="cuda") for k, v in host.items()}
for rep in range(3):
    a, b = ev(), ev(); a.record()
    for _ in range(10):
        for k in host: dev[k].copy_(host[k], non_blocking=True)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("H2D %.1f MB per step in %.3f ms = %.1f GB/s" % (nbytes / 1e6, ms, nbytes / ms / 1e6))
g = torch.randn(24, 256, 32, 32, device="cuda"); gh = torch.empty_like(g, device="cpu").pin_memory()
for rep in range(3):
    a, b = ev(), ev(); a.record()
    for _ in range(10): gh.copy_(g, non_blocking=True)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("D2H %.1f MB in %.3f ms = %.1f GB/s" % (g.numel() * 4 / 1e6, ms, g.numel() * 4 / ms / 1e6))
