"""cProfile (tottime) of the drop-in step at batch 1: which Python-level calls cost the host time."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, ucd_b200 as U
wl = dict(bench.WORKLOAD, B=1)
H, W, C_old = wl["H"], wl["W"], wl["C_old"]
inp = {k: v.cuda() for k, v in bench.make_inputs(0, 1, wl).items()}
def step():
    f_n = inp["f_n"].detach().requires_grad_(True)
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    return tup
for _ in range(50): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(500): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(26)
