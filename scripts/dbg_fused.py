import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ucd_b200 as U
from oracle import ucd_oracle as O
B, C, c_old, h, w, scale = 1, 6, 6, 4, 4, 16
H, W = h * scale, w * scale
g = torch.Generator().manual_seed(1)
lr = torch.randn(B, C, h, w, generator=g) * 3; lo = torch.randn(B, c_old, h, w, generator=g) * 3
lab = torch.randint(0, C, (B, H, W), generator=g)
for term in ("ce", "kd"):
    lr_ref = lr.double().requires_grad_(True)
    out = O.upsample_bilinear(lr_ref, H, W); old = O.upsample_bilinear(lo.double(), H, W)
    ce_ref = O.unbiased_ce(out, lab.clone(), c_old, 255, "none").mean(); kd_ref = O.unbiased_kd(out, old, 1.0)
    (ce_ref if term == "ce" else kd_ref).backward()
    lr_c = lr.cuda().requires_grad_(True)
    ce, kd = U.FusedUnbiasedLosses(old_cl=c_old)(lr_c, lo.cuda(), lab.cuda())
    (ce if term == "ce" else kd).backward()
    d = (lr_c.grad.cpu().double() - lr_ref.grad)
    print(term, "max ref", float(lr_ref.grad.abs().max()), "max err", float(d.abs().max()))
    print((d.abs()[0].amax(0) / lr_ref.grad.abs().max()).numpy().round(4))
