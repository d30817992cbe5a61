"""GPU timeline of one drop-in step (kernel start, duration, idle gap before it) from torch.profiler (CUPTI).
Dev tool: shows where the step's time goes beyond the kernels themselves (launch gaps, host syncs).  Under torchrun
(N ranks) every rank runs the exchanged step and rank 0 prints its timeline (kernels of all its streams, NCCL included)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import ucd_b200 as U

fusedmode = len(sys.argv) > 1 and sys.argv[1] == "fused"
# UCD_WL=voc|ade|city[:batch] picks the workload (default: the bench default)
_wl = os.environ.get("UCD_WL", "voc").split(":")
wl = bench.get_workload(_wl[0], int(_wl[1]) if len(_wl) > 1 else 0)
B, H, W, C_old = wl["B"], wl["H"], wl["W"], wl["C_old"]
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
inp = {k: v.to(dev) for k, v in bench.make_inputs(rank, B, wl).items()}
conloss = U.PixelConLossV2(temperature=0.07, gather_negatives=world > 1)
unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
fused = U.FusedUnbiasedLosses(old_cl=C_old, ignore_index=255, alpha=1.0)


def step():
    f_n = inp["f_n"].detach().requires_grad_(True)
    lr = inp["logits_lr"].detach().requires_grad_(True)
    if fusedmode:
        tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"], max_label=wl["max_label"])
        con = conloss(*tup)
        ce, kd = fused(lr, inp["l_po"], inp["labels"])
    else:
        with torch.no_grad():   # the old model runs first (train.py:100-102), then the new one (:105-108)
            outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
        outputs = U.interpolate_bilinear(lr, (H, W))
        tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"], max_label=wl["max_label"])
        ce = unce(outputs, inp["labels"]).mean()
        con = conloss(*tup)
        kd = unkd(outputs, outputs_old)
    (ce + con / 100 + 10 * kd).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
NSTEP = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
per = len(ev) // NSTEP
# the third step (steady state)
sel = ev[2 * per:3 * per]
t0 = sel[0].time_range.start
prev_end = ev[2 * per - 1].time_range.end
busy = gaps = 0.0
print("%8s %8s %8s  %s" % ("start", "dur", "gap", "kernel"))
for e in sel:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = s - prev_end
    print("%8.1f %8.1f %8.1f  %s" % (s - t0, d, gap, e.name[:90]))
    busy += d
    gaps += max(gap, 0.0)
    prev_end = max(prev_end, e.time_range.end)
print("kernels per step %d | busy %.1f us | idle gaps %.1f us | span %.1f us" % (per, busy, gaps, sel[-1].time_range.end - t0))
if os.environ.get("UCD_HOST_OPS"):
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
