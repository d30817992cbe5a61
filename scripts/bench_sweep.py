"""Config-5 style scaling sweep of the contrastive path alone: N_px per GPU in {8k..256k}, fwd+bwd."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ucd_b200 as U
from bench import make_inputs, WORKLOAD

con = U.PixelConLossV2(temperature=0.07)
for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64, 128, 256]:
    wl = dict(WORKLOAD, H=128, W=128)           # labels at 128x128 keep the host generator light; h=w=32 as in VOC
    inp = {k: v.cuda() for k, v in make_inputs(0, B, wl).items()}
    def run():
        f_n = inp["f_n"].clone().requires_grad_(True)
        tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
        loss = con(*tup)
        loss.backward()
        return tup[0].shape[0], tup[1].shape[0], loss
    for _ in range(2):
        na, nc, loss = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5 if B <= 64 else 2
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    pairs = na * nc
    print(json.dumps(dict(n_px=B * 1024, n_a=na, n_c=nc, pairs=pairs, ms=round(ms, 3), gpairs_s=round(pairs / ms / 1e6, 1),
                          tflops_alg=round(pairs * 1056 / ms / 1e9, 1), loss=round(float(loss), 5),
                          mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))))
