"""All multi-GPU parity modes of tests/test_sp_gpu.py in ONE process group (the pytest file spawns a fresh group per
case, ~100 s each on a fresh box): run under torchrun on 2 GPUs,

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py

Every rank computes the single-GPU answer itself and compares; rank 0 prints one line per mode (with the seconds it
took).  The two transports are constructed ONCE and reused: the symmetric-memory rendezvous of a p2p transport is the
slow part of a fresh group."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200"), os.path.join(ROOT, "tests")]
from helpers import rel_l2  # noqa: E402
from pandora_b200 import DDIMSampler, DDIMSampler_multicond, LatentDiffusionLite, UNetModel, config  # noqa: E402
from pandora_b200.dist import CFGParallel, SequenceParallel  # noqa: E402
from pandora_b200.synthetic import randomize_module_, synth_inputs  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))


def build(dtype):
    p = config.small_unet()
    m = randomize_module_(UNetModel(**p).eval(), 21).to(device="cuda", dtype=dtype)
    inp = {k: v.cuda() for k, v in synth_inputs(1, (4, 16, 16, 16), p["context_dim"], seed=5).items()}
    return p, m, inp


def same_on_all_ranks(x):
    xs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(xs, x.contiguous())
    return all(torch.equal(xs[0], y) for y in xs)


T0 = [time.time()]


def report(mode, err, tol, same):
    ok = same and err < tol
    if rank == 0:
        print(f"{'PASS' if ok else 'FAIL'}  {mode:28s} rel-L2 vs single GPU = {err:.3e} (tol {tol:g}); ranks identical: {same}"
              f"  [{time.time() - T0[0]:.1f} s]", flush=True)
    T0[0] = time.time()
    return ok


SP = {"p2p": SequenceParallel(transport="p2p"), "nccl": SequenceParallel(transport="nccl")}
if rank == 0:
    print(f"transports ready [{time.time() - T0[0]:.1f} s]", flush=True)
ok = True
t = torch.tensor([499], device="cuda")
for mode, tol in (("unet-fp32-p2p", 1e-4), ("unet-bf16-p2p", 2e-2), ("unet-fp32-nccl", 1e-4), ("unet-bf16-nccl", 2e-2)):
    dtype = torch.bfloat16 if "bf16" in mode else torch.float32
    p, m, inp = build(dtype)
    ref = m(inp["x"], t, context=inp["context"], x_concat=[inp["c_concat"]]).float()
    m.sequence_parallel = SP[mode.split("-")[-1]]
    y = m(inp["x"], t, context=inp["context"], x_concat=[inp["c_concat"]])
    y2 = m(inp["x"], t, context=inp["context"], x_concat=[inp["c_concat"]])          # slots / epochs reused
    ok &= report(mode, rel_l2(y.float(), ref), tol, same_on_all_ranks(y.float()) and torch.equal(y, y2))

for mode in ("cfg-eager", "cfg-graph", "sp-eager-p2p", "sp-graph-p2p", "sp-eager-nccl", "sp-graph-nccl", "sp-graph-p2p-noxT",
             "cfg-graph-noxT", "sp-graph-p2p-multicond"):
    parts = mode.split("-")
    p, m, inp = build(torch.float32)
    model = LatentDiffusionLite(m, **config.DIFFUSION_512).cuda()
    cls = DDIMSampler_multicond if "multicond" in parts else DDIMSampler
    cond = dict(c_crossattn=[inp["context"]], c_concat=[inp["c_concat"]])
    uc = dict(c_crossattn=[inp["uc_context"]], c_concat=[inp["c_concat"]])
    extra = {}
    if "multicond" in parts:
        extra = dict(cfg_img=2.0, unconditional_conditioning_img_nonetext=dict(
            c_crossattn=[torch.cat([inp["uc_context"][:, :77], inp["context"][:, 77:]], 1)], c_concat=[inp["c_concat"]]))

    def sample(s, seed):
        torch.manual_seed(seed)
        x, _ = s.sample(S=4, batch_size=1, shape=(4, 16, 16, 16), conditioning=cond, unconditional_conditioning=uc,
                        unconditional_guidance_scale=7.5, eta=1.0, x_T=None if "noxT" in parts else inp["x"],
                        verbose=False, timestep_spacing="uniform_trailing", guidance_rescale=0.7, **extra)
        return x.float()

    ref = sample(cls(model), 100)                 # rank 0's seed: its noise is what gets broadcast
    s = cls(model)
    if parts[0] == "cfg":
        s.cfg_parallel = CFGParallel()
    else:
        m.sequence_parallel = SP["nccl" if "nccl" in parts else "p2p"]
    s.replica_group = dist.group.WORLD
    s.use_cuda_graph = "graph" in parts
    x = sample(s, 100 + rank)                     # different seeds: the broadcast must make the noise identical
    ok &= report(mode, rel_l2(x, ref), 1e-4, same_on_all_ranks(x))
    m.sequence_parallel = None

# Teardown: captured graphs that contain NCCL kernels keep the communicator busy and destroy_process_group() can then
# block for minutes; drop them first, and do not let a stuck teardown hold the GPUs.
import gc
import threading

dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
threading.Timer(20.0, lambda: os._exit(0 if ok else 1)).start()
gc.collect()
dist.destroy_process_group()
os._exit(0 if ok else 1)
