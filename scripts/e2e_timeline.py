"""Timeline of the e2e leg (host buffers, side-stream copies): where do the extra ~0.3 ms per step go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
import ucd_b200 as U

wl = dict(bench.WORKLOAD)
B, H, W, C_old = wl["B"], wl["H"], wl["W"], wl["C_old"]
dev = torch.device("cuda", 0)
host = bench.make_inputs(0, B, wl)
pinned = {k: (v.to(torch.uint8) if k == "labels" else v).pin_memory() for k, v in host.items()}
conloss = U.PixelConLossV2(temperature=0.07)
unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
state = {}

def step(inp):
    f_n = inp["f_n"].detach().requires_grad_(True)
    lr = inp["logits_lr"].detach().requires_grad_(True)
    with torch.no_grad():   # the old model runs first (train.py:100-102), then the new one (:105-108)
        outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
    outputs = U.interpolate_bilinear(lr, (H, W))
    tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
    ce = unce(outputs, inp["labels"]).mean()
    con = conloss(*tup)
    kd = unkd(outputs, outputs_old)
    (ce + con / 100 + 10 * kd).backward()
    state.update(con=con.detach(), ce=ce.detach(), kd=kd.detach(), g_fn=f_n.grad, g_lr=lr.grad)

out_host = dict(losses=torch.empty(3).pin_memory(), g_fn=torch.empty_like(host["f_n"]).pin_memory(),
                g_lr=torch.empty_like(host["logits_lr"]).pin_memory())
copy_stream = torch.cuda.Stream()
d2h_stream = torch.cuda.Stream()     # one stream per direction, like bench.py
main = torch.cuda.current_stream()
slots = [dict() for _ in range(2)]

def prefetch(slot):
    with torch.cuda.stream(copy_stream):
        for k, v in pinned.items():
            slot[k] = v.to(dev, non_blocking=True)
        slot["ready"] = torch.cuda.Event()
        slot["ready"].record(copy_stream)

def e2e_run(n_steps, mode):
    prefetch(slots[0])
    for i in range(n_steps):
        cur = slots[i % 2]
        main.wait_event(cur["ready"])
        if i + 1 < n_steps and mode != "nocopy":
            prefetch(slots[(i + 1) % 2])
        elif i + 1 < n_steps:
            slots[(i + 1) % 2] = cur
        inp = {k: cur[k] for k in pinned}
        inp["labels"] = inp["labels"].to(torch.long)
        for v in inp.values():
            v.record_stream(main)
        step(inp)
        if mode == "nocopy":
            continue
        losses = torch.stack([state["con"], state["ce"], state["kd"]])
        done = torch.cuda.Event(); done.record(main)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            out_host["losses"].copy_(losses, non_blocking=True)
            if mode != "lossonly":
                out_host["g_fn"].copy_(state["g_fn"], non_blocking=True)
                out_host["g_lr"].copy_(state["g_lr"], non_blocking=True)
            for t_ in (losses, state["g_fn"], state["g_lr"]):
                t_.record_stream(d2h_stream)
    main.wait_stream(d2h_stream)

for mode in ("full", "lossonly", "nocopy"):
    e2e_run(5, mode); torch.cuda.synchronize()
    t0 = time.perf_counter(); e2e_run(20, mode); torch.cuda.synchronize()
    print("%s: %.3f ms per step (wall)" % (mode, 1e3 * (time.perf_counter() - t0) / 20))

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e2e_run(6, "full"); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
names = [e.name for e in ev]
starts = [i for i, n in enumerate(names) if "upsample_fwd" in n][::2]
a, b = starts[3], starts[4]
t0 = ev[a].time_range.start
print("one steady step: %.1f us between the first kernels of consecutive steps" % (ev[b].time_range.start - t0))
prev_end = None
for e in ev[a:b]:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    is_copy = e.name.startswith("Memcpy")
    if prev_end is not None and not is_copy and s - prev_end > 8:
        print("   gap %.1f us before %s at %.1f" % (s - prev_end, e.name[:60], s))
    if is_copy and d > 20:
        print("   copy %-28s start %.1f dur %.1f" % (e.name, s, d))
    if not is_copy:
        prev_end = max(prev_end or 0, s + d)
