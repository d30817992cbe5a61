"""2-GPU NCCL test of the data-parallel contrastive path (skipped on a single-GPU box): every rank packs its
own images, the contrast columns are all-gathered, and loss / gradients must match the rank-sharded oracle
(== the reference on the rank-concatenated batch)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ucd_oracle as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ucd_b200 as U
        cases = [O.synthetic_case(2, 16, 16, 256, 256, 8, 6, rank=r, correlated=True) for r in range(world)]
        c = cases[rank]
        f_n = c["f_n"].cuda().requires_grad_(True)
        con = U.PixelConLossV2(temperature=0.07, gather_negatives=True, ddp_grad_scale=False)
        loss = con(*U.pre_contrastive_pixel(f_n, c["labels"].cuda(), l_po=c["l_po"].cuda(), f_o=c["f_o"].cuda()))
        loss.backward()
        # oracle on all ranks' inputs
        f_refs = [x["f_n"].double().requires_grad_(True) for x in cases]
        per_rank, Cg, lcg, _ = O.pre_contrastive_pixel_global(
            f_refs, [x["labels"] for x in cases], [x["l_po"].double() for x in cases], [x["f_o"].double() for x in cases])
        ref = O.pixel_con_loss_global(per_rank, Cg, lcg)
        ref.backward()
        g_ref = f_refs[rank].grad
        a, b = f_n.grad.double().cpu().reshape(-1), g_ref.reshape(-1)
        cos = float(a @ b / (a.norm() * b.norm()))
        # the sync-free module (no 5-tuple) takes the same exchange step
        f_s = c["f_n"].cuda().requires_grad_(True)
        loss_s = U.PixelContrastiveDistillation(temperature=0.07, gather_negatives=True, ddp_grad_scale=False)(
            f_s, c["labels"].cuda(), c["l_po"].cuda(), c["f_o"].cuda())
        loss_s.backward()
        s_rel = abs(loss_s.item() - loss.item()) / abs(loss.item())
        s_cos = float(torch.nn.functional.cosine_similarity(f_s.grad.reshape(1, -1).double(),
                                                            f_n.grad.reshape(1, -1).double()))
        ret[rank] = (abs(loss.item() - ref.item()) / abs(ref.item()), cos, float(a.norm() / b.norm()), s_rel, s_cos)
    finally:
        dist.destroy_process_group()


def test_global_negatives_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29600 + (os.getpid() % 1000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        for r in range(world):
            rel, cos, ratio, s_rel, s_cos = ret[r]
            assert s_rel <= 1e-5 and s_cos >= 1 - 1e-6, (r, s_rel, s_cos)
            assert rel <= 1e-3, (r, rel)
            assert cos >= 0.999, (r, cos)
            assert abs(ratio - 1) < 2e-2, (r, ratio)
