"""One forward+backward of the contrastive path at the bench workload (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ucd_b200 as U
from bench import make_inputs, WORKLOAD
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
inp = {k: v.cuda() for k, v in make_inputs(0, B, WORKLOAD).items()}
con = U.PixelConLossV2(temperature=0.07)
for _ in range(reps):
    f_n = inp["f_n"].clone().requires_grad_(True)
    loss = con(*U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"]))
    loss.backward()
torch.cuda.synchronize()
print(float(loss))
