"""cProfile of the drop-in step at batch 1 (host-bound regime): where does the Python time go?"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, ucd_b200 as U
wl = dict(bench.WORKLOAD, B=1)
H, W, C_old = wl["H"], wl["W"], wl["C_old"]
inp = {k: v.cuda() for k, v in bench.make_inputs(0, 1, wl).items()}
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
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print("wall per step: %.3f ms" % (1e3 * (time.perf_counter() - t0) / 200))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
