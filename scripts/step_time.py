"""Plain timing of the drop-in step (no per-call events, no sampler thread): device ms per step over N steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import ucd_b200 as U
wl = dict(bench.WORKLOAD); B = int(sys.argv[1]) if len(sys.argv) > 1 else wl["B"]
H, W, C_old = wl["H"], wl["W"], wl["C_old"]
inp = {k: v.cuda() for k, v in bench.make_inputs(0, B, wl).items()}
conloss = U.PixelConLossV2(temperature=0.07)
unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
def step():
    f_n = inp["f_n"].detach().requires_grad_(True)
    lr = inp["logits_lr"].detach().requires_grad_(True)
    outputs = U.interpolate_bilinear(lr, (H, W))
    with torch.no_grad():
        outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    ce = unce(outputs, inp["labels"]).mean()
    con = conloss(*tup)
    kd = unkd(outputs, outputs_old)
    (ce + con / 100 + 10 * kd).backward()
for _ in range(5): step()
torch.cuda.synchronize()
for n in (20, 20, 50, 50, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    print("steps %d: device %.3f ms/step, wall %.3f ms/step" % (n, e0.elapsed_time(e1) / n, 1e3 * (time.perf_counter() - t0) / n))
# host time of one step when the GPU is idle at the sync point: launch cost only
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("host time of one isolated step (launch to return): %.3f ms" % (1e3 * (t1 - t0)))
