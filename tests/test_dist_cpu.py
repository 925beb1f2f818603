"""CPU, world_size 2, gloo: the host-side multi-GPU logic (sharding arithmetic, frame<->pixel re-shard,
clip-wide GroupNorm statistics from per-rank partial sums, CFG branch exchange) without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from pandora_b200.dist import CFGParallel, SequenceParallel, make_groups, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_ranks(fn, world=2):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 9, 100):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _reshard(rank, world):
    T, HW, C = 16, 24, 8
    g = torch.Generator().manual_seed(0)
    full = torch.randn(T, HW, C, generator=g)                     # (t, p, c) of the whole clip
    sp = SequenceParallel()
    assert sp.transport == "nccl"                                 # gloo group: the torch.distributed transport
    sp.check(T, HW)
    t0, t1 = sp.frame_range(T)
    mine = full[t0:t1].reshape(-1, C).contiguous()                # frame-sharded rows (t_local, p)
    pix = sp.frames_to_pixels(mine, 1, t1 - t0, HW)               # -> (t, p_local)
    hwl = HW // world
    want = full[:, rank * hwl:(rank + 1) * hwl].reshape(-1, C)
    ok1 = torch.equal(pix, want)
    back = sp.pixels_to_frames(pix, 1, t1 - t0, HW)
    ok2 = torch.equal(back, mine)
    # two samples in one exchange (cond | uncond batched): rows (b, t_local, p) <-> (b, t, p_local)
    full2 = torch.stack([full, full.flip(0) * 2.0])               # (b, t, p, c)
    mine2 = full2[:, t0:t1].reshape(-1, C).contiguous()
    pix2 = sp.frames_to_pixels(mine2, 2, t1 - t0, HW)
    ok1 = ok1 and torch.equal(pix2, full2[:, :, rank * hwl:(rank + 1) * hwl].reshape(-1, C))
    ok2 = ok2 and torch.equal(sp.pixels_to_frames(pix2, 2, t1 - t0, HW), mine2)
    # clip-wide GroupNorm statistics from per-rank partial sums == statistics of the whole tensor
    G = 4
    x = pix.double().reshape(-1, G, C // G)
    sums = torch.stack([x.sum(dim=(0, 2)), (x * x).sum(dim=(0, 2))], dim=1).reshape(-1)
    sp.all_reduce_sum(sums)
    cnt = T * HW * (C // G)
    mean = sums.view(G, 2)[:, 0] / cnt
    var = sums.view(G, 2)[:, 1] / cnt - mean ** 2
    ref = F.group_norm(full.permute(2, 0, 1)[None].double(), G)   # (1, C, T, HW)
    mine_n = (pix.double().reshape(-1, G, C // G) - mean[None, :, None]) / torch.sqrt(var + 1e-5)[None, :, None]
    want_n = ref[0].permute(1, 2, 0)[:, rank * hwl:(rank + 1) * hwl].reshape(-1, G, C // G)
    ok3 = bool((mine_n - want_n).abs().max() < 1e-9)
    v = sp.gather_frames(full[t0:t1].permute(2, 0, 1)[None, :, :, :, None].contiguous())
    ok4 = torch.equal(v[0, :, :, :, 0], full.permute(2, 0, 1))
    return ok1 and ok2 and ok3 and ok4


def test_sequence_parallel_reshard_and_groupnorm_stats():
    assert all(run_ranks(_reshard, 2))


def _cfg(rank, world):
    cfg = CFGParallel()
    cond, uncond = torch.full((1, 4, 2), 3.0), torch.full((1, 4, 2), -1.0)
    mine = cfg.my_branch(cond, uncond)
    v_local = mine * 2 + 1                                         # stand-in denoiser
    v_c, v_u = cfg.exchange(v_local)
    v = v_u + 7.5 * (v_c - v_u)
    return bool(torch.equal(v_c, cond * 2 + 1) and torch.equal(v_u, uncond * 2 + 1) and float(v[0, 0, 0]) == -1 + 7.5 * 8)


def test_cfg_branch_exchange():
    assert all(run_ranks(_cfg, 2))


def _groups(rank, world):
    dp, cfg_g, sp_g, rep_g = make_groups(world, cfg_ways=2, sp_ways=2)
    assert dist.get_world_size(rep_g) == 4
    return (dp, dist.get_world_size(cfg_g), dist.get_rank(cfg_g), dist.get_world_size(sp_g), dist.get_rank(sp_g))


def test_group_partition_cfg_x_sp():
    out = run_ranks(_groups, 4)
    assert out == [(0, 2, 0, 2, 0), (0, 2, 0, 2, 1), (0, 2, 1, 2, 0), (0, 2, 1, 2, 1)]
