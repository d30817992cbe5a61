"""Device times of the streaming kernels (upsample / UNCE / UNKD, fwd+bwd) at a given shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ucd_b200 as U
B, C, C_old, H, W = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (24, 17, 16, 512, 512)
h, w = H // 16, W // 16
lr = (torch.randn(B, C, h, w, device="cuda") * 3).requires_grad_(True)
lpo = torch.randn(B, C_old, h, w, device="cuda") * 3
lab = torch.randint(0, C, (B, H, W), device="cuda")
unce = U.UnbiasedCrossEntropy(old_cl=C_old, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss()
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
npx = B * H * W
out = U.interpolate_bilinear(lr, (H, W)); old = U.interpolate_bilinear(lpo, (H, W)).detach()
g = torch.randn_like(out)
res = {}
res["up_fwd"] = (timeit(lambda: U.interpolate_bilinear(lr, (H, W))), npx * 4 * C)
res["up_bwd"] = (timeit(lambda: torch.autograd.grad(U.interpolate_bilinear(lr, (H, W)), lr, g)) - res["up_fwd"][0], npx * 4 * C)
x = out.detach().requires_grad_(True)
def nograd(fn):
    def run():
        with torch.no_grad():   # forward only: no autograd graph (and no gradient chain on the persistent input)
            return fn()
    return run
res["unce_fwd"] = (timeit(nograd(lambda: unce(x, lab))), npx * (4 * C + 12))
res["unce_fwd+bwd"] = (timeit(lambda: torch.autograd.grad(unce(x, lab).mean(), x)), npx * (12 * C + 36))
res["unkd_fwd"] = (timeit(nograd(lambda: unkd(x, old))), npx * (4 * C + 4 * C_old))
res["unkd_fwd+bwd"] = (timeit(lambda: torch.autograd.grad(unkd(x, old), x)), npx * (12 * C + 8 * C_old + 24))
print("shape B=%d C=%d C_old=%d %dx%d" % (B, C, C_old, H, W))
for k, (ms, by) in res.items():
    print("  %-14s %8.3f ms  %7.1f GB/s (%.2f of 6454)" % (k, ms, by / ms / 1e6, by / ms / 1e6 / 6454))
# ---- N1 fused path: same losses from the low-res logits ----
fused = U.FusedUnbiasedLosses(old_cl=C_old)
def unfused():
    o = U.interpolate_bilinear(lr, (H, W))
    with torch.no_grad():
        oo = U.interpolate_bilinear(lpo, (H, W))
    l = unce(o, lab).mean() + 10 * unkd(o, oo)
    return torch.autograd.grad(l, lr)
def fused_run():
    ce, kd = fused(lr, lpo, lab)
    return torch.autograd.grad(ce + 10 * kd, lr)
print("  unfused upsample+CE+KD fwd+bwd: %.3f ms   fused (N1): %.3f ms" % (timeit(unfused), timeit(fused_run)))
