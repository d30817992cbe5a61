"""Micro-benchmark of the contrastive kernels with per-kernel device times (torch.profiler / CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import ucd_b200 as U
from ucd_b200 import _lib
_lib.debug_lib(as_product=True)   # the traced build (per-role cycle counters) serves the modules of this process
from bench import make_inputs, WORKLOAD

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
inp = {k: v.cuda() for k, v in make_inputs(0, B, WORKLOAD).items()}
if os.environ.get("UCD_SKEW"):   # one dominant new class (real label maps are far more skewed than the bench blobs)
    H = inp["labels"].shape[-1]
    inp["labels"][:, H // 25:, :] = WORKLOAD["C_old"]
    inp["labels"][:, 9 * H // 10:, : H // 4] = WORKLOAD["C"] - 1
con = U.PixelConLossV2(temperature=0.07)

def run(grad):
    f_n = inp["f_n"].clone().requires_grad_(grad)
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    loss = con(*tup)
    if grad:
        loss.backward()
    return loss

for grad in (True, False):
    for _ in range(3):
        run(grad)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run(grad)
        torch.cuda.synchronize()
    print("== requires_grad =", grad)
    rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if "ucd::" in e.key]
    for k, t, c in sorted(rows, key=lambda r: -r[1]):
        print("%10.1f us  x%-3d %s" % (t, c, k[:90]))

# ---- per-role cycle counters (ucd_con_debug_trace) ----
L = _lib.debug_lib()
n_px = B * 32 * 32
rt, ct = (n_px + 127) // 128, L.ucd_con_max_tiles(n_px)
splits = L.ucd_con_debug_splits(rt, ct)
for grad in (True, False):
    f_n = inp["f_n"].clone().requires_grad_(grad)
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    rt_act = (tup[0].shape[0] + 127) // 128
    splits = L.ucd_con_debug_splits(rt_act, ct)
    trace = torch.zeros(2 * rt_act * splits, 16, dtype=torch.int64, device="cuda")
    L.ucd_con_debug_trace(_lib.ptr(trace))
    con(*tup)
    torch.cuda.synchronize()
    L.ucd_con_debug_trace(None)
    tr = trace.cpu().double()
    tr = (tr[:rt_act * splits], tr[rt_act * splits:])
    for sw in (0, 1):
        x = tr[sw][tr[sw][:, 8] > 0]
        if x.numel() == 0:
            continue
        nt = x[:, 11].clamp_min(1)
        print("grad=%s sweep %d: CTAs %d tiles/CTA %.1f | per tile cycles: epilogue total %.0f (wait S %.0f, wait E %.0f) | "
              "mma total %.0f (idle %.0f) | producer total %.0f (wait stage %.0f, wait P %.0f)" % (
                  grad, sw + 1, x.shape[0], nt.mean(), (x[:, 8] / nt).mean(), (x[:, 9] / nt).mean(), (x[:, 10] / nt).mean(),
                  (x[:, 4] / nt).mean(), (x[:, 5] / nt).mean(), (x[:, 0] / nt).mean(), (x[:, 1] / nt).mean(), (x[:, 2] / nt).mean()))
        print("      per CTA cycles: set-up %.0f | tile loop %.0f | tail (wait last MMAs, write partials) %.0f" % (
            x[:, 7].mean(), x[:, 8].mean(), x[:, 15].mean()))
        print("      V/U(+P) issuer per tile: total %.0f (waiting %.0f)" % ((x[:, 12] / nt).mean(), (x[:, 13] / nt).mean()))
        tot = x[:, 7] + x[:, 8] + x[:, 15]
        tiles = x[:, 11]
        print("      balance: tiles total %.0f, per CTA max %.0f / mean %.1f | CTA cycles max %.0f mean %.0f | sum/148 SMs "
              "%.0f cycles | loop cycles per tile (pooled) %.0f" % (tiles.sum(), tiles.max(), tiles.mean(), tot.max(),
              tot.mean(), tot.sum() / 148, x[:, 8].sum() / tiles.sum().clamp_min(1)))
