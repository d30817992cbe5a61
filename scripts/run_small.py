"""One small pass over every kernel of the library (for compute-sanitizer runs): drop-in step, fused N1, sync-free N4,
compat path with dense P, sibling losses, pixel-to-pixel branches, self-contrast losses, bf16 feature hand-off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ucd_b200 as U
from oracle import ucd_oracle as O

B, h, w, H, W, C, C_old = 2, 16, 16, 256, 256, 8, 6
case = {k: v.cuda() for k, v in O.synthetic_case(B, h, w, H, W, C, C_old, correlated=True).items()}
f_n = case["f_n"].clone().requires_grad_(True)
lr = case["logits_lr"].clone().requires_grad_(True)
out = U.interpolate_bilinear(lr, (H, W))
with torch.no_grad():
    old = U.interpolate_bilinear(case["l_po"], (H, W))
tup = U.pre_contrastive_pixel(f_n, case["labels"], l_po=case["l_po"], f_o=case["f_o"])
con = U.PixelConLossV2(temperature=0.07)
loss = (U.UnbiasedCrossEntropy(old_cl=C_old, reduction="none")(out, case["labels"].clone()).mean() + con(*tup) / 100
        + 10 * U.UnbiasedKnowledgeDistillationLoss()(out, old))
loss.backward()
ce, kd = U.FusedUnbiasedLosses(old_cl=C_old)(lr, case["l_po"], case["labels"].clone())
sf = U.PixelContrastiveDistillation()(f_n, case["labels"], case["l_po"], case["f_o"])
(ce + kd + sf).backward()
P = tup[4].dense()
a = tup[0].detach().clone().requires_grad_(True)
con(a, tup[1], tup[2], tup[3], P).backward()
con(a, tup[1], tup[2], tup[3], None).backward()
U.KnowledgeDistillationLoss()(out, old).backward(retain_graph=True)
U.MaskKnowledgeDistillationLoss()(out, old, mask=(case["labels"] > 0).float()).backward(retain_graph=True)
U.MaskCrossEntropy(old_cl=C_old)(out, case["labels"].clone(), outputs_old=old).backward()
o2, l2 = U.pre_contrastive_pixel(f_n, case["labels"], f_o=case["f_o"])
o2.sum().backward()
# self-contrast siblings (sweep 3) on the pixel-to-pixel rows, SupCon with two views
o1, l1 = U.pre_contrastive_pixel(f_n, case["labels"])
U.PixelConLoss(temperature=0.5)(o1, l1).backward()
xs = torch.nn.functional.normalize(torch.randn(150, 2, 64, device="cuda"), dim=2).requires_grad_(True)
U.SupConLoss(contrast_mode="one")(xs, torch.randint(0, 5, (150,), device="cuda")).backward()
# bf16 feature hand-off
fb = case["f_n"].to(torch.bfloat16).requires_grad_(True)
con(*U.pre_contrastive_pixel(fb, case["labels"], l_po=case["l_po"], f_o=case["f_o"].to(torch.bfloat16))).backward()
torch.cuda.synchronize()
print("ok", float(loss), float(sf))
