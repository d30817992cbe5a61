"""A few iterations of the N1 fused module at the bench shape (target for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, ucd_b200 as U
wl = dict(bench.WORKLOAD)
inp = {k: v.cuda() for k, v in bench.make_inputs(0, wl["B"], wl).items()}
fused = U.FusedUnbiasedLosses(old_cl=wl["C_old"], ignore_index=255, alpha=1.0)
for _ in range(3):
    lr = inp["logits_lr"].detach().requires_grad_(True)
    ce, kd = fused(lr, inp["l_po"], inp["labels"])
    (ce + 10 * kd).backward()
torch.cuda.synchronize()
print("ok", float(ce), float(kd))
