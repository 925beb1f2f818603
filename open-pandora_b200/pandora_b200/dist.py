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
                 the clip-wide GroupNorm statistics of those blocks are summed over the ranks.

Sequence-parallel transports
  "p2p"   the data path is two KERNELS of libpandora_b200 over peer memory (symmetric memory mapped through NVLink):
          `pb_sp_reshard` pulls every rank's share straight out of the peers' send slots into the consumer's row
          order (no pack / unpack pass, no NCCL call; the producing GEMM writes into the slot), and
          `pb_sp_groupnorm_affine` adds the per-rank GroupNorm partial sums; both carry their own flag barrier, so
          they are CUDA-graph capturable.
  "nccl"  `all_to_all_single` + torch permutes + `all_reduce` (the first cut; runs on gloo/CPU tensors too, which is
          how the host logic is tested without GPUs, tests/test_dist_cpu.py).
`VirtualSequenceParallel` runs P ranks as THREADS of one process on ONE device with the same two kernels (host
barriers instead of device flags): the frame<->pixel addressing, ragged shards and clip-wide statistics are
exercised on a single-GPU box (tests/test_sp_virtual_gpu.py).
"""
import threading

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Items [lo, hi) handled by `rank` (reference: `indices = range(split*rank, split*(rank+1))`, plus the
    remainder spread over the first ranks instead of being dropped)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _SPCommon:
    """Frame <-> pixel re-sharding of a clip over P ranks.

    Row layouts (C = channels, contiguous), b = samples (1, or 2 when the CFG branches are batched):
      frame-sharded  [b, T/P frames, H*W pixels, C]   -> rows (b, t_local, p)
      pixel-sharded  [b, T frames,  H*W/P pixels, C]  -> rows (b, t, p_local)
    Rank r owns frames [r*T/P, (r+1)*T/P) and pixels [r*HW/P, (r+1)*HW/P)."""
    P = 1
    rank = 0
    batch_ok = True            # the re-shard handles several samples (cond|uncond batched in one forward)

    def check(self, T, HW):
        if T % self.P or HW % self.P:
            raise RuntimeError(f"sequence parallel degree {self.P} must divide frames ({T}) and pixels ({HW})")

    def frame_range(self, T):
        tl = T // self.P
        return self.rank * tl, (self.rank + 1) * tl

    def prepare(self, max_rows, max_cols, samples, device):
        """Called (by every rank, same arguments) at the start of a forward: size the communication buffers."""

    def send_buffer(self, rows, cols, device):
        """Where the producer of the NEXT exchange should write its fp32 rows (None = anywhere)."""
        return None


# ------------------------------------------------------------------------------------------- NCCL / gloo
class _NcclTransport(_SPCommon):
    def _init_group(self, group):
        self.group = group
        self.P = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bytes_moved = 0

    def frames_to_pixels(self, x, samples, t_local, hw, out=None):
        """x: [b*t_local*hw, C] (frame-sharded rows) -> [b * P*t_local * hw/P, C] (pixel-sharded rows)."""
        P, C = self.P, x.shape[1]
        hwl = hw // P
        send = x.reshape(samples, t_local, P, hwl, C).permute(2, 0, 1, 3, 4).contiguous()   # [dst, b, tl, hwl, C]
        recv = torch.empty_like(send)                                                       # [src = frame block, ...]
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.bytes_moved += send.numel() * send.element_size()
        res = recv.permute(1, 0, 2, 3, 4).reshape(samples * P * t_local * hwl, C)           # (b, t = (src, tl), pl)
        if out is not None:
            out.copy_(res)
            return out
        return res.contiguous()

    def pixels_to_frames(self, y, samples, t_local, hw, out=None):
        """y: [b * T * hw/P, C] (pixel-sharded rows) -> [b*t_local*hw, C] (frame-sharded rows)."""
        P, C = self.P, y.shape[1]
        hwl = hw // P
        send = y.reshape(samples, P, t_local, hwl, C).permute(1, 0, 2, 3, 4).contiguous()   # [dst, b, tl, hwl, C]
        recv = torch.empty_like(send)                                                       # [src = pixel block, ...]
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.bytes_moved += send.numel() * send.element_size()
        res = recv.permute(1, 2, 0, 3, 4).reshape(samples * t_local * hw, C)                # (b, tl, p = (src, pl))
        if out is not None:
            out.copy_(res)
            return out
        return res.contiguous()

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def groupnorm(self, abi, x, y, n_inst, rows_local, gamma, beta, eps, silu):
        abi.groupnorm_sp(x, y, n_inst, rows_local, gamma, beta, eps, silu, self.all_reduce_sum, rows_local * self.P)

    def gather_frames(self, v_local):
        """v_local: (b, c, T/P, h, w) -> (b, c, T, h, w) on every rank."""
        parts = [torch.empty_like(v_local) for _ in range(self.P)]
        dist.all_gather(parts, v_local.contiguous(), group=self.group)
        return torch.cat(parts, dim=2)


# ------------------------------------------------------------------------------------------- pull kernels
class _PullTransport(_SPCommon):
    """Shared logic of the kernel transports: two alternating send slots + two GroupNorm-sum slots per rank, all
    addressable by every rank; `pb_sp_reshard` / `pb_sp_groupnorm_affine` do the exchanges."""
    N_SLOTS = 2
    MAX_GN_INST = 4

    def _init_pull(self):
        self._k = 0              # send slot of the next re-shard
        self._g = 0              # sums slot of the next clip-wide GroupNorm
        self._slot_bytes = 0
        self._slots = None       # [N_SLOTS] flat uint8 tensors (local)
        self._slot_ptrs = None   # [N_SLOTS][P] device addresses
        self._sums = None
        self._sums_ptrs = None
        self._flag_ptrs = (None, None)      # (reshard channel, groupnorm channel): lists of P addresses or None
        self._epoch_ptrs = (None, None)
        self.bytes_moved = 0

    # -- hooks
    def _alloc(self, slot_bytes, device):
        raise NotImplementedError

    def _order_ranks(self):
        """Host-side ordering point between 'every rank has enqueued its producer' and the pull (virtual ranks)."""

    def prepare(self, max_rows, max_cols, samples, device):
        need = max_rows * max_cols * 4
        if self._slots is None or need > self._slot_bytes:
            self._alloc(need, device)
            self._slot_bytes = need
            self._k = self._g = 0

    def send_buffer(self, rows, cols, device):
        assert rows * cols * 4 <= self._slot_bytes, "sequence-parallel send slot too small (prepare() not called?)"
        return self._slots[self._k][:rows * cols * 4].view(torch.float32).view(rows, cols)

    def _staged(self, x, abi):
        """x as rows inside the current send slot (copied there unless its producer already wrote it there)."""
        slot = self._slots[self._k]
        if x.data_ptr() == slot.data_ptr() and x.is_contiguous():
            return x
        buf = self.send_buffer(x.shape[0], x.shape[1], x.device)
        abi.copy_rows(x, buf)
        return buf

    def _reshard(self, x, direction, samples, t_local, hw, out):
        from . import abi
        assert x.dtype == torch.float32
        P, C = self.P, x.shape[1]
        hwl = hw // P
        src = self._staged(x, abi)
        rows = samples * t_local * hw if direction == 1 else samples * P * t_local * hwl
        dst = out if out is not None else torch.empty((rows, C), dtype=x.dtype, device=x.device)
        self._order_ranks()
        abi.sp_reshard(self._slot_ptrs[self._k], src.stride(0), dst, direction, self.rank, samples, t_local, hwl, C, 4,
                       self._flag_ptrs[0], self._epoch_ptrs[0])
        self._k = (self._k + 1) % self.N_SLOTS
        self.bytes_moved += rows * C * 4 * (P - 1) // P
        return dst

    def frames_to_pixels(self, x, samples, t_local, hw, out=None):
        return self._reshard(x, 0, samples, t_local, hw, out)

    def pixels_to_frames(self, y, samples, t_local, hw, out=None):
        return self._reshard(y, 1, samples, t_local, hw, out)

    def groupnorm(self, abi, x, y, n_inst, rows_local, gamma, beta, eps, silu):
        assert n_inst <= self.MAX_GN_INST
        g = self._g
        abi.groupnorm_sp_peers(x, y, n_inst, rows_local, gamma, beta, eps, silu, self._sums[g], self._sums_ptrs[g],
                               self.rank, rows_local * self.P, self._flag_ptrs[1], self._epoch_ptrs[1],
                               before_reduce=self._order_ranks)
        self._g = (g + 1) % self.N_SLOTS


class SequenceParallel(_NcclTransport, _PullTransport):
    """One clip over the ranks of `group` (real ranks, one process per GPU).

    transport: "p2p" (kernels over NVLink peer memory), "nccl", or "auto" = p2p when the symmetric-memory rendezvous
    works on this box, else nccl."""

    def __init__(self, group=None, transport="auto"):
        self._init_group(group)
        self._init_pull()
        self._hdls = []
        if transport == "auto":
            transport = "p2p" if self._p2p_available() else "nccl"
        self.transport = transport

    def _p2p_available(self):
        if dist.get_backend(self.group) != "nccl" or not torch.cuda.is_available():
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        except Exception:
            return False
        return True

    # ---- symmetric allocation: [flags ch0 | flags ch1 | sums x2 | slot x2] in ONE rendezvous
    def _alloc(self, slot_bytes, device):
        import torch.distributed._symmetric_memory as symm_mem
        P = self.P
        slot_bytes = (slot_bytes + 255) // 256 * 256
        sums_bytes = self.MAX_GN_INST * 32 * 2 * 8
        hdr = 1024
        total = hdr + self.N_SLOTS * sums_bytes + self.N_SLOTS * slot_bytes
        group = self.group if self.group is not None else dist.group.WORLD
        name = group.group_name if hasattr(group, "group_name") else group
        buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, name)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)           # nobody signals into a flag array that is not zeroed yet
        self._hdls.append((buf, hdl))            # keep the mapping alive
        bases = [int(p) for p in hdl.buffer_ptrs]
        self._flag_ptrs = ([b for b in bases], [b + 256 for b in bases])
        epochs = torch.zeros(4, dtype=torch.int32, device=device)
        self._epochs = epochs
        self._epoch_ptrs = (epochs.data_ptr(), epochs.data_ptr() + 8)
        off = hdr
        self._sums, self._sums_ptrs = [], []
        for _ in range(self.N_SLOTS):
            self._sums.append(buf[off:off + sums_bytes].view(torch.float64))
            self._sums_ptrs.append([b + off for b in bases])
            off += sums_bytes
        self._slots, self._slot_ptrs = [], []
        for _ in range(self.N_SLOTS):
            self._slots.append(buf[off:off + slot_bytes])
            self._slot_ptrs.append([b + off for b in bases])
            off += slot_bytes

    def prepare(self, max_rows, max_cols, samples, device):
        if self.transport == "p2p":
            _PullTransport.prepare(self, max_rows, max_cols, samples, device)

    def send_buffer(self, rows, cols, device):
        if self.transport == "p2p":
            return _PullTransport.send_buffer(self, rows, cols, device)
        return None

    def frames_to_pixels(self, x, samples, t_local, hw, out=None):
        if self.transport == "p2p":
            return _PullTransport.frames_to_pixels(self, x, samples, t_local, hw, out)
        return _NcclTransport.frames_to_pixels(self, x, samples, t_local, hw, out)

    def pixels_to_frames(self, y, samples, t_local, hw, out=None):
        if self.transport == "p2p":
            return _PullTransport.pixels_to_frames(self, y, samples, t_local, hw, out)
        return _NcclTransport.pixels_to_frames(self, y, samples, t_local, hw, out)

    def groupnorm(self, abi, x, y, n_inst, rows_local, gamma, beta, eps, silu):
        if self.transport == "p2p":
            return _PullTransport.groupnorm(self, abi, x, y, n_inst, rows_local, gamma, beta, eps, silu)
        return _NcclTransport.groupnorm(self, abi, x, y, n_inst, rows_local, gamma, beta, eps, silu)


# ------------------------------------------------------------------------------------------- virtual ranks
class VirtualSequenceParallel:
    """P sequence-parallel ranks as threads of ONE process on ONE device (single-GPU test harness).

    `hub.run(fn)` calls `fn(rank_view)` on P threads; inside, `model.sequence_parallel = hub` resolves to the calling
    thread's rank view (`current()`).  The exchanges use the same pull kernels as the p2p transport; ranks are ordered
    by host barriers (every thread enqueues on the same CUDA stream, so 'all producers enqueued before any pull is
    enqueued' is enough) instead of device flags."""

    batch_ok = True

    def __init__(self, P):
        self.P = P
        self._barrier = threading.Barrier(P)
        self._tls = threading.local()
        self._views = [_VirtualRank(self, r) for r in range(P)]
        self._gather = [None] * P
        self._alloc_lock = threading.Lock()
        self._arena = None       # (slot_bytes, slots[rank][k], sums[rank][k])

    def current(self):
        v = getattr(self._tls, "view", None)
        if v is None:
            raise RuntimeError("VirtualSequenceParallel used outside hub.run()")
        return v

    def barrier(self):
        self._barrier.wait()

    def run(self, fn):
        """fn(rank) on P threads, lock-step at every exchange; returns the list of results (exceptions re-raised)."""
        out, err = [None] * self.P, [None] * self.P
        stream = torch.cuda.current_stream()
        dev = torch.cuda.current_device()

        def work(r):
            try:
                torch.cuda.set_device(dev)
                self._tls.view = self._views[r]
                with torch.cuda.stream(stream):
                    out[r] = fn(r)
            except BaseException as e:  # noqa: BLE001
                err[r] = e
                self._barrier.abort()

        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.P)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        self._barrier.reset()
        real = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)]
        if real or any(err):
            raise (real[0] if real else next(e for e in err if e is not None))
        return out


class _VirtualRank(_PullTransport):
    def __init__(self, hub, rank):
        self.hub, self.P, self.rank = hub, hub.P, rank
        self._init_pull()

    def _order_ranks(self):
        self.hub.barrier()

    def _alloc(self, slot_bytes, device):
        hub = self.hub
        hub.barrier()
        if self.rank == 0:
            sums_n = self.MAX_GN_INST * 32 * 2
            hub._arena = (
                [[torch.zeros(slot_bytes, dtype=torch.uint8, device=device) for _ in range(self.N_SLOTS)]
                 for _ in range(self.P)],
                [[torch.zeros(sums_n, dtype=torch.float64, device=device) for _ in range(self.N_SLOTS)]
                 for _ in range(self.P)])
        hub.barrier()
        slots, sums = hub._arena
        self._slots = slots[self.rank]
        self._slot_ptrs = [[slots[r][k].data_ptr() for r in range(self.P)] for k in range(self.N_SLOTS)]
        self._sums = sums[self.rank]
        self._sums_ptrs = [[sums[r][k].data_ptr() for r in range(self.P)] for k in range(self.N_SLOTS)]

    def gather_frames(self, v_local):
        hub = self.hub
        hub._gather[self.rank] = v_local.contiguous()
        hub.barrier()
        full = torch.cat(list(hub._gather), dim=2)
        hub.barrier()
        return full


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
