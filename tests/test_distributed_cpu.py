"""world_size-2 gloo test (CPU) of the one exchange step of the data-parallel path: the single all-gather
of every rank's payload (ucd_b200.losses.gather_contrast_columns / payload_layout).  Payloads are filled here on the
CPU from oracle outputs with the documented layout; the gathered chunks must reproduce the column set of
the rank-sharded oracle (oracle.pre_contrastive_pixel_global), synchronously and with the asynchronous
(overlappable) form of the collective."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ucd_oracle as O

TILE, D = 128, 256


def pack_tiles(rows, n_tiles, width):
    """rows [n, width] -> [n_tiles, width/8, 128, 8] bf16 (include/ucd_b200.h tile layout)."""
    out = torch.zeros(n_tiles * TILE, width)
    out[:rows.shape[0]] = rows
    return out.reshape(n_tiles, TILE, width // 8, 8).permute(0, 2, 1, 3).contiguous().to(torch.bfloat16)


def unpack_tiles(t):
    nt, ch, _, _ = t.shape
    return t.float().permute(0, 2, 1, 3).reshape(nt * TILE, ch * 8)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ucd_b200.losses import gather_contrast_columns, payload_layout, payload_views
        cases = [O.synthetic_case(2, 8, 8, 128, 128, 6, 4, rank=r) for r in range(world)]
        c = cases[rank]
        A, Cst, la, lc, P, prep = O.pre_contrastive_pixel(c["f_n"], c["labels"], c["l_po"], c["f_o"])
        n_px = 2 * 8 * 8
        T = (2 * n_px + 127) // 128 + 1
        p = torch.softmax(c["l_po"].permute(0, 2, 3, 1).reshape(n_px, 4), 1)
        pc = torch.cat([p[torch.from_numpy(prep.anchor)], p[torch.from_numpy(prep.pseudo_mask)]])
        feat = pack_tiles(Cst, T, D)
        prob = pack_tiles(torch.nn.functional.pad(pc, (0, 12)), T, 16)
        lab = torch.full((T * TILE,), -1, dtype=torch.int32)
        lab[:lc.numel()] = lc.to(torch.int32)
        counts = torch.tensor([A.shape[0], Cst.shape[0] - A.shape[0], prep.min_new, n_px], dtype=torch.int32)
        rng = torch.stack([lab.view(T, TILE).clamp_min(0).amin(1), lab.view(T, TILE).amax(1)], 1).to(torch.int32)
        payload = torch.zeros(payload_layout(T, 16)["nbytes"], dtype=torch.uint8)
        pv = payload_views(payload, T, 16)
        pv["feat"].copy_(feat), pv["prob"].copy_(prob), pv["lab"].copy_(lab.view(T, TILE))
        pv["range"].copy_(rng), pv["counts"].copy_(counts)
        g = gather_contrast_columns(payload, T, 16, dist.group.WORLD)
        g2 = gather_contrast_columns(payload, T, 16, dist.group.WORLD, async_op=True)   # the overlappable form
        g2["work"].wait()
        assert torch.equal(g2["buf"], g["buf"]) and g["chunk_stride"] == payload.numel() and g["rank"] == rank
        assert g["range"].shape == (world, T, 2) and torch.equal(g["range"][rank], rng)
        # reference: the rank-sharded oracle on all ranks' inputs
        per_rank, Cg, lcg, min_new = O.pre_contrastive_pixel_global(
            [x["f_n"] for x in cases], [x["labels"] for x in cases], [x["l_po"] for x in cases], [x["f_o"] for x in cases])
        assert g["n_chunks"] == world and g["chunk_tiles"] == T and g["self_tile0"] == rank * T
        assert int(g["counts"][:, 2].min()) == min_new   # the sweeps take the global threshold from the headers
        cols, labs, off = [], [], 0
        for r in range(world):
            n_c = int(g["counts"][r, :2].sum())
            cols.append(unpack_tiles(g["feat"][r])[:n_c])
            labs.append(g["lab"][r].reshape(-1)[:n_c])
            assert bool((g["lab"][r].reshape(-1)[n_c:] == -1).all())
            # self columns of rank r's anchors are the first N_a columns of chunk r
            assert torch.equal(per_rank[r][3], off + torch.arange(int(g["counts"][r][0])))
            off += n_c
        cols, labs = torch.cat(cols), torch.cat(labs)
        assert cols.shape == Cg.shape and float((cols - Cg).abs().max()) < 2 ** -8
        assert torch.equal(labs.long(), lcg)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gather_contrast_columns_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}
