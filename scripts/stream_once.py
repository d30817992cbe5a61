"""One pass of each streaming kernel at the bench shape (for ncu captures of the HBM-bound kernels)."""
import os, sys
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
for _ in range(2):
    out = U.interpolate_bilinear(lr, (H, W))
    with torch.no_grad():
        old = U.interpolate_bilinear(lpo, (H, W))
    (unce(out, lab).mean() + 10 * unkd(out, old)).backward()
torch.cuda.synchronize()
