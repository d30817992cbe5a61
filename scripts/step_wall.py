"""Per-step wall / device time of the drop-in step of a bench workload, plus the allocator's device-malloc count:
shows whether a slow bench line is uniformly slow or hit by periodic stalls (cudaMalloc / cudaFree, host syncs).
usage: python scripts/step_wall.py <voc|ade|city>[:batch] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ucd_b200 as U

spec = (sys.argv[1] if len(sys.argv) > 1 else "voc").split(":")
wl = bench.get_workload(spec[0], int(spec[1]) if len(spec) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, H, W, C_old = wl["B"], wl["H"], wl["W"], wl["C_old"]
inp = {k: v.to(dev) for k, v in bench.make_inputs(0, B, wl).items()}
conloss = U.PixelConLossV2(temperature=0.07)
unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)


def step():
    f_n = inp["f_n"].detach().requires_grad_(True)
    lr = inp["logits_lr"].detach().requires_grad_(True)
    with torch.no_grad():   # the old model runs first (train.py:100-102), then the new one (:105-108)
        outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
    outputs = U.interpolate_bilinear(lr, (H, W))
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"], max_label=wl["max_label"])
    ce = unce(outputs, inp["labels"]).mean()
    con = conloss(*tup)
    kd = unkd(outputs, outputs_old)
    (ce + con / 100 + 10 * kd).backward()
    return f_n.grad, lr.grad


def stat(k):
    return torch.cuda.memory_stats().get(k, 0)


rows = []
for i in range(n):
    m0, s0 = stat("num_device_alloc"), stat("num_device_free")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    keep = step() if (len(sys.argv) > 3 and sys.argv[3] == "keep") else (step(), None)[1]
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t0), stat("num_device_alloc") - m0, stat("num_device_free") - s0))
print("%s: step  host_ms  total_ms  cudaMalloc  cudaFree" % wl["name"])
for i, r in enumerate(rows):
    print("%4d  %7.3f  %7.3f  %3d %3d" % ((i,) + r))
print("reserved MB %.0f allocated MB %.0f" % (torch.cuda.memory_reserved() / 1e6, torch.cuda.memory_allocated() / 1e6))
# back-to-back loop like bench.py (no sync between steps)
for k in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    m0 = stat("num_device_alloc")
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    print("back-to-back x10: %.3f ms/step, cudaMalloc %d" % (a.elapsed_time(b) / 10, stat("num_device_alloc") - m0))
