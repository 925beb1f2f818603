"""Multi-GPU plumbing for the denoising hot path (one process per GPU, torch.distributed / NCCL).

The reference only ever slices a prompt list across ranks and never exchanges data
(`DynamiCrafter/scripts/evaluation/inference.py:276-282`, `ddp_wrapper.py:8-12`).  The path shards on three
independent axes:

  prompt batch   `shard_range` — same arithmetic as the reference's launcher; zero collectives.
  CFG branches   `CFGParallel`  — the cond / uncond forwards of a step (ddim.py:233-234) run on two ranks;
                 one all-gather of v per step, then every rank performs the (deterministic) update.
  sequence       `SequenceParallel` — one clip across P ranks: frames are sharded for the per-frame layers
                 (conv2d, per-frame GroupNorm, spatial attention), pixels are sharded for the per-pixel layers
                 (temporal conv, temporal attention); an all-to-all re-shards around every temporal block and
                 the clip-wide GroupNorm statistics of those blocks are all-reduced (2 x 32 fp64 per sample).

Everything here is host-side glue over `torch.distributed`; it runs unchanged on gloo/CPU tensors, which is how
the logic is tested without GPUs (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Items [lo, hi) handled by `rank` (reference: `indices = range(split*rank, split*(rank+1))`, plus the
    remainder spread over the first ranks instead of being dropped)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SequenceParallel:
    """Frame <-> pixel re-sharding of a single clip (batch 1) over `group`.

    Row layouts (C = channels, contiguous):
      frame-sharded  [T/P frames, H*W pixels, C]   -> rows (t_local, p)
      pixel-sharded  [T frames,  H*W/P pixels, C]  -> rows (t, p_local)
    Rank r owns frames [r*T/P, (r+1)*T/P) and pixels [r*HW/P, (r+1)*HW/P)."""

    def __init__(self, group=None):
        self.group = group
        self.P = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bytes_moved = 0

    def check(self, T, HW):
        if T % self.P or HW % self.P:
            raise RuntimeError(f"sequence parallel degree {self.P} must divide frames ({T}) and pixels ({HW})")

    def frame_range(self, T):
        tl = T // self.P
        return self.rank * tl, (self.rank + 1) * tl

    def frames_to_pixels(self, x, t_local, hw):
        """x: [t_local*hw, C] (frame-sharded rows) -> [P*t_local * hw/P, C] (pixel-sharded rows)."""
        P, C = self.P, x.shape[1]
        hwl = hw // P
        send = x.reshape(t_local, P, hwl, C).permute(1, 0, 2, 3).contiguous()      # [dst rank, t_local, hwl, C]
        recv = torch.empty_like(send)                                              # [src rank = frame block, ...]
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.bytes_moved += send.numel() * send.element_size()
        return recv.view(P * t_local * hwl, C)

    def pixels_to_frames(self, y, t_local, hw):
        """y: [T * hw/P, C] (pixel-sharded rows) -> [t_local*hw, C] (frame-sharded rows)."""
        P, C = self.P, y.shape[1]
        hwl = hw // P
        send = y.contiguous()                                                      # already [dst rank, t_local, hwl, C]
        recv = torch.empty_like(send)                                              # [src rank = pixel block, ...]
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.bytes_moved += send.numel() * send.element_size()
        return recv.view(P, t_local, hwl, C).permute(1, 0, 2, 3).reshape(t_local * hw, C).contiguous()

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_frames(self, v_local):
        """v_local: (b, c, T/P, h, w) -> (b, c, T, h, w) on every rank."""
        parts = [torch.empty_like(v_local) for _ in range(self.P)]
        dist.all_gather(parts, v_local.contiguous(), group=self.group)
        return torch.cat(parts, dim=2)


class CFGParallel:
    """Two-way split of the classifier-free-guidance branches over `group` (size 2):
    rank 0 of the group evaluates the conditional branch, rank 1 the unconditional one."""

    def __init__(self, group=None):
        self.group = group
        if dist.get_world_size(group) != 2:
            raise RuntimeError("CFGParallel needs a process group of exactly 2 ranks")
        self.rank = dist.get_rank(group)

    def my_branch(self, cond, uncond):
        return cond if self.rank == 0 else uncond

    def exchange(self, v_local):
        """v of this rank's branch -> (v_cond, v_uncond) on both ranks."""
        parts = [torch.empty_like(v_local), torch.empty_like(v_local)]
        dist.all_gather(parts, v_local.contiguous(), group=self.group)
        return parts[0], parts[1]


def make_groups(world, cfg_ways=1, sp_ways=1):
    """Partition `world` ranks as dp x cfg x sp (sp fastest).  Returns (dp_index, cfg_group, sp_group,
    replica_group) for the calling rank — replica_group = all ranks working on the same clip (noise is broadcast
    inside it); every rank must call this with the same arguments (new_group is collective)."""
    rank = dist.get_rank()
    if world % (cfg_ways * sp_ways):
        raise RuntimeError(f"world {world} is not divisible by cfg_ways*sp_ways = {cfg_ways * sp_ways}")
    dp = world // (cfg_ways * sp_ways)
    my_sp = my_cfg = my_rep = None
    for d in range(dp):
        ranks = list(range(d * cfg_ways * sp_ways, (d + 1) * cfg_ways * sp_ways))
        g = dist.new_group(ranks) if len(ranks) > 1 else None
        if rank in ranks:
            my_rep = g
        for c in range(cfg_ways):
            ranks = [(d * cfg_ways + c) * sp_ways + s for s in range(sp_ways)]
            g = dist.new_group(ranks) if sp_ways > 1 else None
            if rank in ranks:
                my_sp = g
        for s in range(sp_ways):
            ranks = [(d * cfg_ways + c) * sp_ways + s for c in range(cfg_ways)]
            g = dist.new_group(ranks) if cfg_ways > 1 else None
            if rank in ranks:
                my_cfg = g
    return rank // (cfg_ways * sp_ways), my_cfg, my_sp, my_rep
