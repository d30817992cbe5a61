#!/usr/bin/env python
"""Benchmark of the UCD distillation-loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one synthetic batch, forward + backward:
  bilinear logit upsample (new: grad, old: no grad) -> pre_contrastive_pixel -> PixelConLossV2
  -> UnbiasedCrossEntropy(...).mean() + con/100 + 10 * UnbiasedKnowledgeDistillationLoss -> backward
through the reference-shaped modules of ucd_b200 (train.py:115-116,133 wiring).

Workload at every N: BASELINE configs[1] per GPU - VOC 15-5s step 1 (17 classes, 16 old), batch 24 at
512x512 (32x32 embeddings, D=256).  Weak scaling: each rank holds its own 24 images; with N>1 the
contrast columns are all-gathered over NCCL so negatives span the global batch.
metric = pixel pairs (N_a local x N_c global, summed over ranks) per second, in Mpairs/s.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ucd_loss_fwd_bwd_pixel_pairs_per_s"
UNIT = "Mpixel-pairs/s"
WORKLOAD = dict(name="VOC 15-5s step 1, batch 24 @512x512 per GPU", B=24, h=32, w=32, H=512, W=512, C=17, C_old=16)
CPU_SAMPLE_B = 3   # images in the bounded CPU sample of the same workload (N^2 fp32 temporaries limit the CPU)

# kernels launched by each C-ABI entry point (for gpu_launches; memsets are not kernels)
KERNELS_PER_CALL = {"ucd_upsample_bilinear_fwd": 1, "ucd_upsample_bilinear_bwd": 1, "ucd_unce_fwd": 1,
                    "ucd_unce_bwd": 1, "ucd_kd_fwd": 2, "ucd_kd_bwd": 1, "ucd_con_prep_labels": 2,
                    "ucd_con_prep_pack": 4, "ucd_con_prep_bwd": 1, "ucd_con_fwd": 5, "ucd_con_bwd": 1,
                    "ucd_con_pack_rows": 1}


def gen(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def make_inputs(rank, B, wl):
    """Synthetic inputs of SURVEY.md 8d: seeds 1-4 (+1000*rank), class-correlated features, blob labels."""
    h, w, H, W, C, C_old = wl["h"], wl["w"], wl["H"], wl["W"], wl["C"], wl["C_old"]
    o = 1000 * rank
    f_n, f_o = gen(1 + o, B, 256, h, w), gen(2 + o, B, 256, h, w)
    l_po = gen(3 + o, B, C_old, h, w, scale=3.0)
    lr = gen(4 + o, B, C, h, w, scale=3.0)
    proto = gen(9, C, 256)
    add = proto[l_po.argmax(1)].permute(0, 3, 1, 2)
    f_n, f_o = 0.3 * f_n + 0.7 * add, 0.3 * f_o + 0.7 * add
    lab = torch.zeros(B, H, W, dtype=torch.int64)
    lab[:, H // 5:3 * H // 5, W // 5:3 * W // 5] = C_old
    lab[:, 3 * H // 5:4 * H // 5, W // 10:2 * W // 5] = C - 1
    lab[:, :H // 25, :] = 255
    return dict(f_n=f_n.contiguous(), f_o=f_o.contiguous(), l_po=l_po, logits_lr=lr, labels=lab)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 5 ms in a thread; the
    nvidia-smi query of the profiling recipe is the fallback)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.thr, self.max_mhz = gpu_index, [], False, None, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((mhz, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.thr = threading.Thread(target=self._loop, daemon=True)
            self.thr.start()

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self.stop_flag = True
        self.thr.join(timeout=1)
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for _, rs in self.rows:
            bits |= rs
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=self.max_mhz,
                    reasons=[n for n, b in self.REASONS if bits & b], samples=len(sm), source="nvml, 5 ms period")

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return dict(sm_mhz=float(out[0]), sm_max_mhz=float(out[1]),
                        reasons=[n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")],
                        samples=1, source="nvidia-smi after the timed region (NVML unavailable)")
        except Exception:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock query unavailable"], samples=0)


class CallTimer:
    """Proxy around the ctypes library: records a CUDA-event pair on the current stream around every C-ABI call,
    so per-kernel-group device times come from the same timed region as the headline number."""

    def __init__(self, handle):
        self._h, self.events, self.counts, self.enabled = handle, {}, {}, False

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.startswith("ucd_") or name in ("ucd_last_error", "ucd_version", "ucd_device_ok"):
            return fn

        def wrapped(*args):
            if not self.enabled:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.events.setdefault(name, []).append((e0, e1))
            self.counts[name] = self.counts.get(name, 0) + 1
            return rc
        return wrapped

    def totals_ms(self):
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.events.items()}


def cpu_reference_sample(steps, warmup, wl):
    """The reference's CPU implementation of the path (oracle port, torch CPU, all host threads) on a bounded
    sample of the workload: CPU_SAMPLE_B images of the same shapes."""
    from oracle import ucd_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    case = make_inputs(0, CPU_SAMPLE_B, wl)
    times, pairs = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        f_n = case["f_n"].clone().requires_grad_(True)
        lr = case["logits_lr"].clone().requires_grad_(True)
        res = O.hot_path(f_n, case["f_o"], case["l_po"], lr, case["labels"].clone(), old_cl=wl["C_old"])
        res.total.backward()
        dt = time.perf_counter() - t0
        pairs = res.n_anchor * res.n_contrast
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=pairs / (ms * 1e-3) / 1e6, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample="%d of the %d images of one step (N_a x N_c = %d pairs), fwd+bwd of the whole path, mean of %d steps"
                       % (CPU_SAMPLE_B, wl["B"], pairs, len(times))), ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOAD
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    cb, ms = cpu_reference_sample(steps, warmup, wl)
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=wl["name"], sample=cb["sample"], classes=[wl["C"], wl["C_old"]]),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ucd_b200")
    ap.add_argument("--batch", type=int, default=WORKLOAD["B"], help="images per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ucd_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ucd_b200 as U
    from ucd_b200 import _lib
    timer = CallTimer(_lib.lib())
    _lib._lib = timer

    wl = dict(WORKLOAD, B=args.batch)
    B, H, W, C_old = wl["B"], wl["H"], wl["W"], wl["C_old"]
    host = make_inputs(rank, B, wl)
    # the dataloader hands labels over as uint8 and the trainer casts them on the device (train.py:97-98,
    # dataset/transform.py:350): the end-to-end leg copies uint8 labels and casts after the copy, like the reference
    pinned = {k: (v.to(torch.uint8) if k == "labels" else v).pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    conloss = U.PixelConLossV2(temperature=0.07, gather_negatives=world > 1)
    unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
    unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
    state = {}

    def step(inp):
        f_n = inp["f_n"].detach().requires_grad_(True)
        lr = inp["logits_lr"].detach().requires_grad_(True)
        outputs = U.interpolate_bilinear(lr, (H, W))
        with torch.no_grad():
            outputs_old = U.interpolate_bilinear(inp["l_po"], (H, W))
        # evaluation order of train.py:115-116,133: prep, criterion, contrastive loss, distillation
        tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
        ce = unce(outputs, inp["labels"]).mean()
        con = conloss(*tup)
        kd = unkd(outputs, outputs_old)
        loss = ce + con / 100 + 10 * kd
        loss.backward()
        state.update(n_a=tup[0].shape[0], n_c=tup[1].shape[0], con=con.detach(), ce=ce.detach(), kd=kd.detach(),
                     g_fn=f_n.grad, g_lr=lr.grad)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        step(devin)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timer.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(devin)
    e1.record()
    barrier()
    timer.enabled = False
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = e0.elapsed_time(e1) / args.steps
    call_ms = {k: v / args.steps for k, v in timer.totals_ms().items()}
    launches = sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in timer.counts.items())

    # ---- the same step with the opt-in N1 fusion (FusedUnbiasedLosses: no full-res logits), for information ----
    fused = U.FusedUnbiasedLosses(old_cl=C_old, ignore_index=255, alpha=1.0)

    def step_fused(inp):
        f_n = inp["f_n"].detach().requires_grad_(True)
        lr = inp["logits_lr"].detach().requires_grad_(True)
        tup = U.pre_contrastive_pixel(f_n, inp["labels"], l_po=inp["l_po"], f_o=inp["f_o"])
        con = conloss(*tup)
        ce, kd = fused(lr, inp["l_po"], inp["labels"])
        (ce + con / 100 + 10 * kd).backward()

    for _ in range(3):
        step_fused(devin)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_fused(devin)
    f1.record()
    barrier()
    ms_fused = f0.elapsed_time(f1) / args.steps

    # ---- opt-in sync-free step (N1 fused CE/KD + N4 contrastive without the 5-tuple), for information: eager, and
    #      (single GPU) forward + backward replayed from one CUDA graph ----
    con_static = U.PixelContrastiveDistillation(temperature=0.07, gather_negatives=world > 1)
    gs = {k: devin[k].clone() for k in ("f_n", "f_o", "l_po", "logits_lr", "labels")}
    gs["f_n"].requires_grad_(True), gs["logits_lr"].requires_grad_(True)

    def step_static():
        gs["f_n"].grad = gs["logits_lr"].grad = None
        ce, kd = fused(gs["logits_lr"], gs["l_po"], gs["labels"])
        (ce + con_static(gs["f_n"], gs["labels"], gs["l_po"], gs["f_o"]) / 100 + 10 * kd).backward()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            fn()
        a1.record()
        barrier()
        return a0.elapsed_time(a1) / args.steps

    ms_static = timed(step_static)
    ms_graph = None
    if world == 1:
        try:
            if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_static()
            torch.cuda.current_stream().wait_stream(side)
            gs["f_n"].grad = gs["logits_lr"].grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                ce, kd = fused(gs["logits_lr"], gs["l_po"], gs["labels"])
                (ce + con_static(gs["f_n"], gs["labels"], gs["l_po"], gs["f_o"]) / 100 + 10 * kd).backward()
            ms_graph = timed(graph.replay)
        except Exception as exc:  # informational leg only
            ms_graph = "capture failed: %s" % (str(exc).splitlines()[0][:120],)

    # ---- end-to-end through the public API with host buffers (e2e) ----
    # Every step copies its inputs from pinned host memory and copies losses + both gradients back.  Like a
    # DataLoader with pin_memory / non_blocking prefetch, the copies of step i+1 / i-1 run on a side stream while
    # step i computes; all of them are inside the timed region.
    out_host = dict(losses=torch.empty(3).pin_memory(), g_fn=torch.empty_like(host["f_n"]).pin_memory(),
                    g_lr=torch.empty_like(host["logits_lr"]).pin_memory())
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    slots = [dict() for _ in range(2)]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            for k, v in pinned.items():
                slot[k] = v.to(dev, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(copy_stream)

    def e2e_run(n_steps):
        prefetch(slots[0])
        for i in range(n_steps):
            cur = slots[i % 2]
            main.wait_event(cur["ready"])
            if i + 1 < n_steps:
                prefetch(slots[(i + 1) % 2])
            inp = {k: cur[k] for k in pinned}
            inp["labels"] = inp["labels"].to(torch.long)          # train.py:98
            for v in inp.values():
                v.record_stream(main)
            step(inp)
            losses = torch.stack([state["con"], state["ce"], state["kd"]])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                out_host["losses"].copy_(losses, non_blocking=True)
                out_host["g_fn"].copy_(state["g_fn"], non_blocking=True)
                out_host["g_lr"].copy_(state["g_lr"], non_blocking=True)
                for t_ in (losses, state["g_fn"], state["g_lr"]):
                    t_.record_stream(copy_stream)
        main.wait_stream(copy_stream)

    e2e_run(max(3, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2a, e2b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2a.record()
    e2e_run(args.steps)
    e2b.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    ms_e2e = max(e2a.elapsed_time(e2b) / args.steps, wall_ms)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())

    # ---- aggregate over ranks: max time, sum of pairs ----
    n_a = state["n_a"]
    stats = torch.tensor([ms_dev, ms_e2e, float(n_a), float(state["n_c"]), ms_fused], device=dev, dtype=torch.float64)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    allst = torch.stack(allst).cpu()
    ms_dev_max, ms_e2e_max, ms_fused_max = float(allst[:, 0].max()), float(allst[:, 1].max()), float(allst[:, 4].max())
    n_c_global = float(allst[:, 3].sum())
    pairs_total = float((allst[:, 2] * n_c_global).sum())
    pairs_local = float(n_a) * n_c_global

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)      # kernels are timed inside a long step
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained bf16 / copy)" if peaks else "fallback"
        f_pair = 4 * 256 + 2 * C_old                                 # algorithmic flop per pixel pair
        con_ms = call_ms.get("ucd_con_fwd", float("nan"))
        con_tflops = pairs_local * f_pair / (con_ms * 1e-3) / 1e12
        npx = B * H * W
        C = wl["C"]
        hbm = {}
        for name, nbytes in (("ucd_unce_fwd", npx * (4 * C + 8 + 4 + 8)), ("ucd_unce_bwd", npx * (8 * C + 8 + 8)),
                             ("ucd_kd_fwd", npx * (4 * C + 4 * C_old + 12)), ("ucd_kd_bwd", npx * (8 * C + 4 * C_old + 12)),
                             ("ucd_upsample_bilinear_fwd", npx * 4 * (C + C_old)), ("ucd_upsample_bilinear_bwd", npx * 4 * C)):
            if name in call_ms:
                gbs = nbytes / (call_ms[name] * 1e-3) / 1e9
                hbm[name] = dict(ms=round(call_ms[name], 4), achieved_gbs=round(gbs, 1), frac=round(gbs / hbm_peak, 4))
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline, _ = cpu_reference_sample(3, 1, wl)
        line = dict(
            metric=METRIC, value=pairs_total / (ms_dev_max * 1e-3) / 1e6, unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms_dev_max, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="bf16 operands / f32 accumulate (contrastive); f32 (CE, KD, upsample)", data="synthetic",
            config=dict(workload=wl["name"] if B == WORKLOAD["B"] else wl["name"].replace("batch 24", "batch %d" % B),
                        classes=[C, C_old], pixels_per_gpu=B * wl["h"] * wl["w"], n_anchor_rank0=n_a,
                        n_contrast_global=int(n_c_global), pairs_per_step=int(pairs_total), temperature=0.07,
                        l2="no flush: full-res logits (%.0f MB) exceed the 126 MB L2; the %.0f MB of bf16 column tiles are "
                           "meant to be L2-resident" % (npx * C * 4 / 1e6, n_c_global * 512 / 1e6),
                        parallelism="dp%d, all-gathered contrast columns" % world),
            roofline=dict(bound="tensor", kernel="ucd_con_fwd (sweep 1 + combine + sweep 2 + finalize)",
                          achieved=con_tflops, peak=bf16_peak, unit="TFLOP/s", frac=con_tflops / bf16_peak,
                          traffic=(9.18e7 if (world == 1 and B == WORKLOAD["B"]) else None),
                          traffic_note="dram read+write of the two sweep kernels per step, ncu --set full, profiles/r01f_ncu_full.md (sweep 1: 21.4 MB read + 45.0 MB written, sweep 2: 22.9 + 2.6)",
                          flop_per_pair=f_pair, ms=con_ms, peak_source=peak_src),
            roofline_hbm=hbm,
            call_ms={k: round(v, 4) for k, v in sorted(call_ms.items())},
            cpu_baseline=cpu_baseline,
            e2e=dict(value=pairs_total / (ms_e2e_max * 1e-3) / 1e6, unit=UNIT, ms_per_step=ms_e2e_max,
                     h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
            n1_fused=dict(note="same step with the opt-in FusedUnbiasedLosses (upsample+CE+KD from low-res logits)",
                          ms_per_step=ms_fused_max, value=pairs_total / (ms_fused_max * 1e-3) / 1e6, unit=UNIT),
            n4_sync_free=dict(note="rank 0: N1 fused CE/KD + PixelContrastiveDistillation (no 5-tuple, no host sync); "
                                   "graph = forward+backward replayed from one CUDA graph (single GPU only)",
                              ms_per_step_eager=ms_static, ms_per_step_graph=ms_graph),
            gpu_launches=int(launches), clocks=clocks,
            losses=dict(con=float(state["con"]), ce=float(state["ce"]), kd=float(state["kd"])),
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
