#!/usr/bin/env python
"""BASELINE configs[1] end to end: VOC 15-5s overlapped step 1, batch 24 at 512x512 = 3 images per GPU on 8 GPUs,
DeepLabv3-ResNet-101 (random init) + the UCD loss, one process per GPU under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/config2_step.py

What is the product here and what is not: the backbone (torchvision ResNet-101, output stride 16, DeepLab ASPP head ->
256-d `pre_logits` -> 1x1 classifier; stock cuDNN kernels, torch DDP for the parameter gradients) stands in for the
reference's `IncrementalSegmentationModule` (segmentation_module.py:63-143; its inplace_abn / apex dependencies are not
installable here) and is NOT what this repo builds.  The step wiring is the reference trainer's (train.py:95-151):
old model forward under no_grad, new model forward, the hot path (this repo), backward, SGD step.  Three loss
variants are timed on the same models and inputs:
  none      a stand-in loss (means of the outputs) - the step without the hot path
  dropin    the reference-shaped modules: interpolate + pre_contrastive_pixel + PixelConLossV2(gather_negatives) +
            UnbiasedCrossEntropy + UnbiasedKnowledgeDistillationLoss (train.py:115-116,133)
  fused     opt-in rows N1 + N4: FusedUnbiasedLosses on the low-res logits + PixelContrastiveDistillation (no 5-tuple,
            no host sync)
and the loss section of `dropin` / `fused` is also timed on its own with CUDA events (forward part inside the step).
Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class Segmenter(nn.Module):
    """ResNet-101 (output stride 16) + DeepLabv3 head; forward returns (low-res logits, {"pre_logits", "sem"}) - the
    tensors `Trainer.train` hands to the loss (train.py:100-116; full-res logits are produced by the loss side here)."""

    def __init__(self, n_classes):
        super().__init__()
        from torchvision.models import resnet101
        from torchvision.models.segmentation.deeplabv3 import DeepLabHead
        r = resnet101(weights=None, replace_stride_with_dilation=[False, False, True])
        self.body = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4)
        head = DeepLabHead(2048, n_classes)
        self.head = nn.Sequential(*list(head.children())[:-1])   # ASPP + 3x3 conv + BN + ReLU -> 256 channels
        self.cls = list(head.children())[-1]                    # 1x1 classifier (segmentation_module.py:46-49)

    def forward(self, x):
        pre = self.head(self.body(x))
        sem = self.cls(pre)
        return sem, {"pre_logits": pre, "sem": sem}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--classes", type=int, nargs=2, default=[17, 16], help="C C_old (VOC 15-5s step 1)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ucd_b200 as U
    C, C_old = args.classes
    B, H, W = args.batch, 512, 512
    torch.manual_seed(1234)                      # same initial weights on every rank (DDP broadcasts anyway)
    model, model_old = Segmenter(C).to(dev), Segmenter(C_old).to(dev).eval()
    for p in model_old.parameters():
        p.requires_grad_(False)
    model.train()
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)   # run.py:186-190
    g = torch.Generator().manual_seed(100 + rank)
    images = torch.randn(B, 3, H, W, generator=g).to(dev)
    labels0 = torch.zeros(B, H, W, dtype=torch.int64)
    labels0[:, H // 5:3 * H // 5, W // 5:3 * W // 5] = C_old
    labels0[:, 3 * H // 5:4 * H // 5, W // 10:2 * W // 5] = C - 1
    labels0[:, :H // 25] = 255
    labels0 = labels0.to(dev)

    unce = U.UnbiasedCrossEntropy(old_cl=C_old, ignore_index=255, reduction="none")
    unkd = U.UnbiasedKnowledgeDistillationLoss(alpha=1.0)
    con = U.PixelConLossV2(temperature=0.07, gather_negatives=world > 1)
    fused = U.FusedUnbiasedLosses(old_cl=C_old, ignore_index=255, alpha=1.0)
    con_sf = U.PixelContrastiveDistillation(temperature=0.07, gather_negatives=world > 1)
    ev = {}

    def loss_none(sem, feats, sem_old, feats_old, labels):
        return sem.mean() + feats["pre_logits"].mean()

    def loss_dropin(sem, feats, sem_old, feats_old, labels):
        outputs = U.interpolate_bilinear(sem, (H, W))                         # segmentation_module.py:133
        with torch.no_grad():
            outputs_old = U.interpolate_bilinear(sem_old, (H, W))
        tup = U.pre_contrastive_pixel(feats["pre_logits"], labels, l_po=feats_old["sem"], f_o=feats_old["pre_logits"])
        loss = unce(outputs, labels).mean() + con(*tup) / 100                 # train.py:115-116
        return loss + 10 * unkd(outputs, outputs_old)                         # train.py:133

    def loss_fused(sem, feats, sem_old, feats_old, labels):
        ce, kd = fused(sem, sem_old, labels)
        return ce + con_sf(feats["pre_logits"], labels, feats_old["sem"], feats_old["pre_logits"]) / 100 + 10 * kd

    def step(loss_fn, mark=False):
        labels = labels0.clone()                                               # the CE remaps labels in place
        with torch.no_grad():
            sem_old, feats_old = model_old(images)                              # train.py:100-102
        opt.zero_grad(set_to_none=True)
        sem, feats = ddp(images)                                                # train.py:108
        if mark:
            ev["a"].record()
        loss = loss_fn(sem, feats, sem_old, feats_old, labels)
        if mark:
            ev["b"].record()
        loss.backward()                                                        # train.py:135-138 (O0: plain backward)
        opt.step()                                                             # train.py:149
        return loss

    def timed(loss_fn):
        for _ in range(args.warmup):
            step(loss_fn)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fwd_ms = 0.0
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            ev["a"], ev["b"] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            last = step(loss_fn, mark=True)
            ev.setdefault("pairs", []).append((ev["a"], ev["b"]))
        e1.record()
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t0) / args.steps
        for a, b in ev.pop("pairs"):
            fwd_ms += a.elapsed_time(b)
        t = torch.tensor([e0.elapsed_time(e1) / args.steps, wall, fwd_ms / args.steps, float(last)], device=dev,
                         dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(ms_per_step=float(t[0]), wall_ms_per_step=float(t[1]), loss_forward_section_ms=float(t[2]),
                    loss=float(t[3]))

    # the first variant also pays cuDNN autotuning / NCCL set-up: warm the whole step up first, and measure the stand-in
    # loss twice (before and after the two real variants); the later, steadier figure is the base of the shares
    for _ in range(8):
        step(loss_none)
    res = {name: timed(fn) for name, fn in (("none_first", loss_none), ("dropin", loss_dropin), ("fused", loss_fused),
                                            ("none", loss_none))}
    if rank == 0:
        base = res["none"]["ms_per_step"]
        out = dict(config="VOC 15-5s overlapped step 1, DeepLabv3-R101 (random init, torchvision, output stride 16), "
                          "batch %d per GPU x %d GPUs @512x512, fp32, torch DDP + SGD" % (B, world),
                   n_gpus=world, images_per_s={k: B * world / (v["ms_per_step"] * 1e-3) for k, v in res.items()},
                   steps=args.steps, warmup=args.warmup, variants=res,
                   loss_share_of_step={k: (res[k]["ms_per_step"] - base) / res[k]["ms_per_step"] for k in ("dropin", "fused")},
                   note="share = (step with the hot path - step with the stand-in loss) / step with the hot path; "
                        "loss_forward_section_ms = CUDA events around the loss forward inside the step (host-bound at "
                        "this batch: launch gaps included)")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
