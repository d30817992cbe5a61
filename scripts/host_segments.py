"""Host time of each segment of the drop-in step at batch 1 (GPU work negligible: what remains is Python / launch
overhead), to see what sits on the critical path after pre_contrastive_pixel's host sync."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, ucd_b200 as U
wl = dict(bench.WORKLOAD, B=1)
H, W, C_old = wl["H"], wl["W"], wl["C_old"]
inp = {k: v.cuda() for k, v in bench.make_inputs(0, 1, wl).items()}
conloss = U.PixelConLossV2(temperature=0.07)
unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
acc = {}
def seg(name, t0):
    t1 = time.perf_counter_ns()
    acc[name] = acc.get(name, 0) + (t1 - t0)
    return t1
def step():
    t = time.perf_counter_ns()
    f_n = inp["f_n"].detach().requires_grad_(True)
    lr = inp["logits_lr"].detach().requires_grad_(True)
    t = seg("inputs", t)
    with torch.no_grad():   # the old model runs first (train.py:100-102), then the new one (:105-108)
        outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
    outputs = U.interpolate_bilinear(lr, (H, W))
    t = seg("interpolate x2", t)
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    t = seg("pre_contrastive_pixel (incl. sync)", t)
    ce = unce(outputs, inp["labels"])
    t = seg("unce()", t)
    ce = ce.mean()
    t = seg(".mean()", t)
    con = conloss(*tup)
    t = seg("conloss()", t)
    kd = unkd(outputs, outputs_old)
    t = seg("unkd()", t)
    loss = ce + con / 100 + 10 * kd
    t = seg("combine", t)
    loss.backward()
    t = seg("backward", t)
for _ in range(30): step()
torch.cuda.synchronize(); acc.clear()
N = 300
for _ in range(N): step()
torch.cuda.synchronize()
tot = 0
for k, v in acc.items():
    print("%-38s %7.1f us" % (k, v / N / 1e3)); tot += v
print("%-38s %7.1f us" % ("total host per step", tot / N / 1e3))
