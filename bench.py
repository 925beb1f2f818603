#!/usr/bin/env python
"""bench.py — headline benchmark of the video-denoising hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is ONE classifier-free-guidance DDIM step of the DynamiCrafter U-Net at 16 x 320 x 512
(latent 4x16x40x64, context 333x1024): two U-Net forwards (cond / uncond) + the fused CFG / DDIM
update, bf16 compute, synthetic inputs and random-init weights of the real architecture
(1.44 B parameters).  One process per GPU; for N > 1 launch with torchrun.  Rank 0 prints ONE JSON line.

  value          steps/s, all ranks (weak scaling: every rank denoises its own clip, no data-path collective —
                 SURVEY.md 8e axis (2)), inputs resident in HBM, the step replayed as a CUDA graph
  e2e            the same metric through the PUBLIC API: wall clock of `DDIMSampler.sample(S=50, ...)` from pinned
                 HOST conditioning tensors to a HOST latent, pred_x0 read back to the host every step (img_callback);
                 `sec_per_clip` is measured, not derived; the first call (graph capture) is reported separately
  roofline       the dominant kernel family (tc_gemm: every Linear / conv3x3 / temporal conv), algorithmic FLOPs /
                 CUDA-event time from an instrumented eager step, against MEASURED_PEAKS.json (sustained bf16)
  single_clip_1024   BASELINE config 3: ONE 16 x 576 x 1024 clip on all N GPUs (CFG branches on two halves x
                 sequence parallel N/2: frames sharded for the spatial layers, pixels for the temporal ones,
                 re-sharded by the pb_sp_reshard kernels over NVLink peer memory): ms/step, measured s/clip,
                 per-GPU fraction of the bf16 peak, NVLink bytes per step
  torch_gpu_baseline the reference's op sequence as stock torch ops on this GPU (einsum and SDPA attention arms)
  cpu_baseline   the oracle (CPU restatement of the reference, oracle/) on the host cores
  --impl reference   times the reference algorithm's CPU path (oracle port) on the host cores: real CFG steps
                 (2 full-latent fp32 forwards each), no extrapolation
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]

METRIC = "unet_cfg_denoise_steps_per_s_16x320x512"
S_DDIM, CFG_SCALE, RESCALE, ETA = 50, 7.5, 0.7, 1.0
FLOP_PER_FORWARD = {"512": 12602.83e9, "1024": 52336.20e9}
REFERENCE_BUDGET_S = 330.0                                       # wall-clock bound of the --impl reference run


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), bf16_burst=d.get("bf16_tflops"),
                    hbm=d.get("hbm_gbs"), src="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # clocks under load: upper half of the samples (idle samples before/after the region drag the median)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return dict(sm_mhz=(load[len(load) // 2] if load else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=sorted(reasons), samples=len(sm))


def build_native(device):
    from pandora_b200 import UNetModel, config
    from pandora_b200.synthetic import randomize_module_
    with torch.device(device):
        unet = UNetModel(**config.UNET_512).eval()       # UNET_1024 differs only in default_fs (passed explicitly)
    randomize_module_(unet, 1234)
    return unet.to(torch.bfloat16)                       # Open-Pandora: model.to(bf16) (gradio_app.py:128-129)


def sampler_for(unet, res, device):
    from pandora_b200 import DDIMSampler, LatentDiffusionLite, config
    dcfg = config.DIFFUSION_512 if res == "512" else config.DIFFUSION_1024
    model = LatentDiffusionLite(unet, **dcfg).to(device)
    sampler = DDIMSampler(model)
    sampler.make_schedule(S_DDIM, "uniform_trailing", ETA, verbose=False)
    return model, sampler


def oracle_inputs(res):
    from oracle import unet_oracle
    from pandora_b200 import config
    from pandora_b200.synthetic import synth_inputs, synth_state_dict
    ucfg = config.UNET_512 if res == "512" else config.UNET_1024
    latent = config.LATENT_512 if res == "512" else config.LATENT_1024
    sd = synth_state_dict(unet_oracle.unet_param_shapes(ucfg), 1234)
    inp = synth_inputs(1, latent, ucfg["context_dim"], 0)
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    fs = torch.tensor([ucfg["default_fs"]])
    return sd, ucfg, x, inp, fs


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path on all host threads.  The reference is pure Python over
    torch CPU ops, which is exactly what the oracle restates (pinned to it by tests/golden).  One step = one real CFG
    step at the FULL latent: two sequential fp32 U-Net forwards (cond, uncond; ddim.py:233-234) — no crop, no
    extrapolation.  The run is bounded by wall clock: after one warm-up forward it times `steps` steps or as many as
    fit in REFERENCE_BUDGET_S, and reports the number it actually ran."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import unet_oracle
    torch.set_num_threads(os.cpu_count())
    sd, ucfg, x, inp, fs = oracle_inputs("512")
    t = torch.tensor([999])
    t_start = time.perf_counter()
    unet_oracle.unet_forward(sd, ucfg, x, t, inp["context"], fs)          # warm-up (thread pools, allocator)
    ts = []
    while len(ts) < args.steps:
        t0 = time.perf_counter()
        unet_oracle.unet_forward(sd, ucfg, x, t, inp["context"], fs)
        unet_oracle.unet_forward(sd, ucfg, x, t, inp["uc_context"], fs)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + ts[-1] > REFERENCE_BUDGET_S and len(ts) >= 3:
            break
    per_step = sum(ts) / len(ts)
    v = 1.0 / per_step
    line = dict(impl="reference", metric=METRIC, value=v, unit="steps/s", n_gpus=args.gpus, steps=len(ts),
                warmup=1, ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="DynamiCrafter 320x512 16-frame CFG DDIM step (2 U-Net forwards), CPU fp32",
                            latent=[4, 16, 40, 64], context=[333, 1024]),
                cpu_baseline=dict(value=v, unit="steps/s", cores=torch.get_num_threads(), kind="port",
                                  sample=f"{len(ts)} real CFG steps (2 full-latent fp32 oracle forwards each, "
                                         f"{per_step:.1f} s per step) after 1 warm-up forward; requested {args.steps} "
                                         f"steps, bounded to {REFERENCE_BUDGET_S:.0f} s of wall clock"),
                e2e=dict(value=v, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def torch_gpu_baseline(dev, d, fs):
    """What the unmodified reference does on this box: its op sequence through stock PyTorch kernels (cuDNN conv,
    cuBLAS GEMM; attention once as the materialised einsum the reference falls back to without xformers and once as
    torch SDPA, the stand-in for xformers), bf16, eager."""
    from oracle import unet_oracle
    from pandora_b200 import config as _cfg
    from pandora_b200.synthetic import synth_state_dict
    ocfg = _cfg.UNET_512
    sd = {k: v.to(dev, torch.bfloat16) for k, v in synth_state_dict(unet_oracle.unet_param_shapes(ocfg), 1234).items()}
    xo = torch.cat([d["x"], d["c_concat"]], 1).to(torch.bfloat16)
    to = torch.tensor([999], device=dev)
    ctx = d["context"].to(torch.bfloat16)
    out = {}
    try:
        for impl in ("einsum", "sdpa"):
            unet_oracle.ATTN_IMPL = impl
            ts_ms = []
            with torch.no_grad():
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    unet_oracle.unet_forward(sd, ocfg, xo, to, ctx, fs)
                    e1.record()
                    e1.synchronize()
                    ts_ms.append(e0.elapsed_time(e1))
            fwd = min(ts_ms[1:])
            out[impl] = dict(ms_per_forward=round(fwd, 2), steps_per_s=round(1e3 / (2 * fwd), 3))
    finally:
        unet_oracle.ATTN_IMPL = "einsum"
    out["note"] = ("oracle = the reference's op sequence as stock torch ops (cuDNN conv, cuBLAS GEMM), bf16, eager; "
                   "one forward = half a CFG step; informational, not the reference arm")
    del sd
    torch.cuda.empty_cache()
    return out


def vae_decode_leg(dev, latent):
    """per-frame first-stage decode of one clip's latents (what follows the sampler every round,
    ddpm3d.py:630-655) through the native AutoencoderKL.decode, bf16; device time of `decode_first_stage`."""
    from pandora_b200 import AutoencoderKL, config
    from pandora_b200.synthetic import randomize_module_, synth_inputs
    dd = config.VAE_DDCONFIG
    with torch.device(dev):
        vae = AutoencoderKL(ddconfig=dd, embed_dim=4).eval()
    randomize_module_(vae, 1234)
    vae = vae.to(torch.bfloat16)
    z = synth_inputs(1, latent, 8, 5)["x"].to(dev)
    c, t, h, w = latent
    # algorithmic FLOPs of Decoder.forward per frame: 2*M*N*K over the convolutions + 4*HW*HW*C for the attention
    flops, ch, hw = 0.0, dd["ch"] * dd["ch_mult"][-1], h * w
    conv = lambda cin, cout, px, k=9: 2.0 * px * cout * cin * k
    flops += conv(dd["z_channels"], ch, hw) + 4 * conv(ch, ch, hw) + 4 * conv(ch, ch, hw, 1) + 4.0 * hw * hw * ch
    cin = ch
    for lv in reversed(range(len(dd["ch_mult"]))):
        cout = dd["ch"] * dd["ch_mult"][lv]
        for _ in range(dd["num_res_blocks"] + 1):
            flops += conv(cin, cout, hw) + conv(cout, cout, hw) + (conv(cin, cout, hw, 1) if cin != cout else 0)
            cin = cout
        if lv:
            hw *= 4
            flops += conv(cin, cin, hw)
    flops += conv(cin, dd["out_ch"], hw)
    flops *= t
    y = vae.decode_first_stage(z, config.SCALE_FACTOR)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = vae.decode_first_stage(z, config.SCALE_FACTOR)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    out = dict(workload=f"AutoencoderKL.decode of one clip's latents {list(latent)} -> {list(y.shape[1:])}, bf16, eager, "
                        "frame by frame semantics (8 frames per pass)", ms_per_clip=round(ms, 2),
               tflops=round(flops / ms / 1e9, 1), gflop_per_clip=round(flops / 1e9, 1))
    del vae, y
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--res", default="512", choices=["512", "1024"], help="resolution of the headline (weak-scaling) leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-single-clip", action="store_true", help="skip the 16x576x1024 single-clip leg")
    ap.add_argument("--no-vae", action="store_true", help="skip the first-stage decode leg")
    ap.add_argument("--sp", type=int, default=1, help="headline leg: sequence-parallel ways (ONE clip across this many ranks)")
    ap.add_argument("--cfg-par", type=int, default=1, choices=[1, 2],
                    help="headline leg: 2 = the cond / uncond branches of a step on two ranks")
    ap.add_argument("--sp-transport", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get("PB_BENCH_WATCHDOG"):       # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PB_BENCH_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pandora_b200 import abi, config
    from pandora_b200.sampler import GraphedCFGStep
    from pandora_b200.synthetic import synth_inputs
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi.load()
    unet = build_native(dev)
    groups = {}

    def make_clip_groups(cfg_par, sp):
        """(dp_index, cfg_group, sp_group, replica_group) for a dp x cfg x sp partition; cached (new_group is collective)."""
        key = (cfg_par, sp)
        if key not in groups:
            from pandora_b200.dist import make_groups
            groups[key] = make_groups(world, cfg_par, sp) if cfg_par * sp > 1 else (rank, None, None, None)
        return groups[key]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return v

    def leg(res, cfg_par, sp, n_steps, n_warm, clip=True, instrument=False):
        """One benchmark leg: the graph-replayed CFG step (device-resident inputs) + optionally one measured 50-step
        clip through DDIMSampler.sample with host conditioning.  Returns a dict of measurements."""
        from pandora_b200.dist import CFGParallel, SequenceParallel
        latent = config.LATENT_512 if res == "512" else config.LATENT_1024
        ucfg = config.UNET_512 if res == "512" else config.UNET_1024
        clip_ways = cfg_par * sp
        n_clips = world // clip_ways
        dp_index, cfg_group, sp_group, rep_group = make_clip_groups(cfg_par, sp)
        model, sampler = sampler_for(unet, res, dev)
        unet.sequence_parallel = SequenceParallel(sp_group, transport=args.sp_transport) if sp > 1 else None
        sampler.cfg_parallel = CFGParallel(cfg_group) if cfg_par > 1 else None
        sampler.replica_group = rep_group
        inp = synth_inputs(1, latent, ucfg["context_dim"], seed=dp_index)      # the same clip on all its ranks
        shape = (1, *latent)
        d = {k: v.to(dev) for k, v in inp.items()}
        cond = dict(c_crossattn=[d["context"]], c_concat=[d["c_concat"]])
        uc = dict(c_crossattn=[d["uc_context"]], c_concat=[d["c_concat"]])
        fs = torch.tensor([ucfg["default_fs"]], dtype=torch.long, device=dev)
        table, _ = sampler._step_table(False)
        g = GraphedCFGStep(sampler, unet, shape, cond, uc, fs, CFG_SCALE, RESCALE, 1.0, table)
        steps_ddpm = [int(s) for s in reversed(sampler.ddim_timesteps)]
        sp_obj = unet.sequence_parallel
        sp_bytes0 = sp_obj.bytes_moved if sp_obj is not None else 0
        if args.no_graph:
            def one_step(i):
                g.ts.fill_(steps_ddpm[i % S_DDIM])
                g.noise.normal_()
                sampler._sync_noise(g.noise)
                g._body()
            g._body(); g.step_index.zero_()
            sp_passes = 1
        else:
            g.capture()
            sp_passes = 2                              # eager warm-up + the captured pass

            def one_step(i):
                g.step(steps_ddpm[i % S_DDIM])
        # bytes this rank pulled from its peers per step (counted by the binding when the exchanges are enqueued)
        nvlink_bytes = (sp_obj.bytes_moved - sp_bytes0) // sp_passes if sp_obj is not None else 0

        def timed(fn, n_warm_, n):
            g.reset(d["x"])
            for i in range(n_warm_):
                fn(i)
            barrier()
            l0 = abi.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                if (n_warm_ + i) % S_DDIM == 0:
                    g.reset(d["x"])                        # a new clip: x_T, step index 0
                fn(n_warm_ + i)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), abi.launches - l0

        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms_total, launches = timed(one_step, n_warm, n_steps)
        clk = clocks.stop() if rank == 0 else None
        ms_step = ms_total / n_steps
        out = dict(res=res, latent=list(latent), clip_ways=clip_ways, n_clips=n_clips, ms_step=ms_step,
                   launches=launches, clocks=clk, graph=not args.no_graph,
                   parallelism=f"dp{n_clips}" + (f"xcfg{cfg_par}" if cfg_par > 1 else "") + (f"xsp{sp}" if sp > 1 else ""),
                   transport=getattr(unet.sequence_parallel, "transport", None), d=d, fs=fs)
        out["nvlink_bytes_per_step"] = nvlink_bytes
        # ---- instrumented eager step: per-launch CUDA events -> breakdown + roofline of the GEMM family
        if instrument:
            g.step_index.zero_()
            g.ts.fill_(steps_ddpm[0])
            if rank == 0:
                abi.profile_begin()
            g._body()                                  # (every rank of a clip runs it: the step may hold collectives)
            if rank == 0:
                out["profile"] = abi.profile_end()
            torch.cuda.synchronize()
        del g
        # ---- the public API end to end: host conditioning -> DDIMSampler.sample(S=50) -> host latent
        if clip:
            host = {k: inp[k].pin_memory() for k in ("x", "c_concat", "context", "uc_context")}
            host_fs = torch.tensor([ucfg["default_fs"]], dtype=torch.long).pin_memory()
            host_x0 = torch.empty(shape).pin_memory()
            host_out = torch.empty(shape).pin_memory()

            def img_callback(pred_x0, i):                # the step's result read back to the host, every step
                host_x0.copy_(pred_x0, non_blocking=True)

            def one_clip():
                barrier()
                t0 = time.perf_counter()
                cnd = dict(c_crossattn=[host["context"].to(dev, non_blocking=True)],
                           c_concat=[host["c_concat"].to(dev, non_blocking=True)])
                ucd = dict(c_crossattn=[host["uc_context"].to(dev, non_blocking=True)], c_concat=cnd["c_concat"])
                s = type(sampler)(model)                 # Open-Pandora constructs a sampler per call (model.py:705)
                s.cfg_parallel, s.replica_group = sampler.cfg_parallel, sampler.replica_group
                s.use_cuda_graph = not args.no_graph
                x, _ = s.sample(S=S_DDIM, batch_size=1, shape=latent, conditioning=cnd, unconditional_conditioning=ucd,
                                unconditional_guidance_scale=CFG_SCALE, eta=ETA, x_T=host["x"].to(dev, non_blocking=True),
                                fs=host_fs.to(dev, non_blocking=True), verbose=False, timestep_spacing="uniform_trailing",
                                guidance_rescale=RESCALE, img_callback=img_callback, precision=torch.float32)
                host_out.copy_(x, non_blocking=True)
                torch.cuda.synchronize()
                return max_over_ranks(time.perf_counter() - t0)

            first = one_clip()                           # includes the CUDA-graph capture of this shape
            n_clip = 2 if res == "512" else 1
            secs = [one_clip() for _ in range(n_clip)]
            h2d = sum(host[k].numel() * host[k].element_size() for k in host) + 8
            out.update(first_clip_s=first, sec_per_clip=sum(secs) / len(secs), clips_timed=n_clip,
                       h2d_bytes_per_step=h2d / S_DDIM, d2h_bytes_per_step=host_x0.numel() * 4 + host_out.numel() * 4 / S_DDIM)
        unet._graph_cache.clear()
        unet._ctx_cache.clear()
        unet.sequence_parallel = None
        gc.collect()
        torch.cuda.empty_cache()
        return out

    # =============================================================== headline leg (weak scaling, BASELINE config 2)
    A = leg(args.res, args.cfg_par, args.sp, args.steps, args.warmup, clip=True, instrument=True)
    ms_step = A["ms_step"]
    value = A["n_clips"] * 1e3 / ms_step
    pk = peaks()
    breakdown = roofline = None
    if rank == 0:
        agg = {}
        for it in A["profile"]:
            a = agg.setdefault(it["name"], dict(n=0, ms=0.0, flops=0, bytes=0))
            a["n"] += 1; a["ms"] += it["ms"]; a["flops"] += it["flops"]; a["bytes"] += it["bytes"]
        tot = sum(a["ms"] for a in agg.values())
        breakdown = {k: dict(n=a["n"], ms=round(a["ms"], 3), share=round(a["ms"] / tot, 4),
                             tflops=round(a["flops"] / a["ms"] / 1e9, 1), gbs=round(a["bytes"] / a["ms"] / 1e6, 1))
                     for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        fam = [a for k, a in agg.items() if k.startswith("tc_gemm") or k.startswith("tc_conv")]
        # DRAM traffic of the family's most expensive launch shape, from the committed `ncu --set full` capture
        traffic = traffic_note = None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if args.res == "512" and os.path.exists(tpath):
            t0 = json.load(open(tpath))["launches"][0]
            traffic = t0["dram_bytes"]
            traffic_note = dict(launch=t0["launch"], dram_bytes=t0["dram_bytes"], algorithmic_bytes=t0["algorithmic_bytes"],
                                source="profiles/r2_ncu_full_key_metrics.md row 1 (ncu --set full, per launch)")
        if fam:
            fl, ms, n = sum(a["flops"] for a in fam), sum(a["ms"] for a in fam), sum(a["n"] for a in fam)
            ach = fl / ms / 1e9
            roofline = dict(bound="tensor", kernel="tc_gemm_kernel (Linear + conv3x3 + temporal conv, tcgen05)",
                            achieved=round(ach, 1), peak=pk["bf16"], unit="TFLOP/s", frac=round(ach / pk["bf16"], 4),
                            traffic=traffic, traffic_of=traffic_note, launches_per_step=n, avg_launch_us=round(ms / n * 1e3, 2),
                            flops_per_launch=fl / n, share_of_step=round(ms / tot, 4), peak_source=pk["src"] + " sustained",
                            frac_of_burst_peak=round(ach / pk["bf16_burst"], 4),
                            how="per-launch CUDA events on the launching stream, instrumented eager step after the timed region")
    whole = 2 * FLOP_PER_FORWARD[args.res] / (ms_step * 1e-3) / 1e12 / A["clip_ways"]     # per GPU

    # =============================================================== BASELINE config 3: one 576x1024 clip on N GPUs
    single = None
    if not args.no_single_clip and not (args.res == "1024" and args.cfg_par * args.sp == world):
        cfg_par = 2 if world >= 2 else 1
        sp = world // cfg_par
        B = leg("1024", cfg_par, sp, max(3, min(args.steps, 10)), 3, clip=True)
        per_gpu = 2 * FLOP_PER_FORWARD["1024"] / (B["ms_step"] * 1e-3) / 1e12 / world
        single = dict(workload="ONE DynamiCrafter 576x1024 16-frame clip (latent 4x16x72x128), 50-step DDIM, CFG 7.5, "
                               "rescale 0.7, bf16, on all GPUs", parallelism=B["parallelism"], n_gpus=world,
                      sp_transport=B["transport"], ms_per_step=round(B["ms_step"], 3),
                      steps_per_s=round(1e3 / B["ms_step"], 3), sec_per_clip=round(B["sec_per_clip"], 3),
                      sec_per_clip_how="wall clock of DDIMSampler.sample(S=50) from host conditioning to a host latent, "
                                       "max over ranks (graph cached from the first call)",
                      first_clip_s=round(B["first_clip_s"], 3), per_gpu_tflops=round(per_gpu, 1),
                      per_gpu_frac_of_peak=round(per_gpu / pk["bf16"], 4), scaling="strong",
                      nvlink_bytes_per_step=B.get("nvlink_bytes_per_step", 0), clocks=B["clocks"],
                      gpu_launches=B["launches"])

    cpu_baseline = torch_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import unet_oracle
        sd, ocfg, x, oinp, fs_c = oracle_inputs(args.res)
        t0 = time.perf_counter()
        unet_oracle.unet_forward(sd, ocfg, x, torch.tensor([999]), oinp["context"], fs_c)
        sec = time.perf_counter() - t0
        del sd
        cpu_baseline = dict(value=1.0 / (2 * sec), unit="steps/s", cores=torch.get_num_threads(), kind="port",
                            sample=f"1 fp32 oracle U-Net forward at the full latent ({sec:.1f} s, no warm-up); "
                                   "a CFG step = 2 forwards")
    if rank == 0 and world == 1 and not args.no_torch_baseline and args.res == "512":
        torch_gpu = torch_gpu_baseline(dev, A["d"], A["fs"])
    vae_leg = None
    if rank == 0 and world == 1 and not args.no_vae:
        vae_leg = vae_decode_leg(dev, tuple(A["latent"]))
    if rank == 0:
        sec_clip = A["sec_per_clip"]
        line = dict(metric=METRIC if args.res == "512" else METRIC.replace("320x512", "576x1024"), value=value,
                    unit="steps/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                    higher_is_better=True, scaling="weak" if A["clip_ways"] == 1 else "strong", vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=dict(workload="DynamiCrafter 320x512 16-frame 50-step DDIM, CFG 7.5, guidance_rescale 0.7, "
                                         "eta 1.0, uniform_trailing; step = 2 U-Net forwards + fused CFG/DDIM update"
                                if args.res == "512" else "DynamiCrafter 576x1024 16-frame CFG DDIM step",
                                latent=A["latent"], context=[333, 1024], clips_in_flight=A["n_clips"],
                                parallelism=A["parallelism"], params=1438854980, cuda_graph=A["graph"],
                                l2="working set per step (2.9 GB bf16 weights + activations) exceeds the 126 MB L2"),
                    e2e=dict(value=A["n_clips"] * S_DDIM / sec_clip, unit="steps/s",
                             h2d_bytes_per_step=A["h2d_bytes_per_step"], d2h_bytes_per_step=A["d2h_bytes_per_step"],
                             ms_per_step=1e3 * sec_clip / S_DDIM, sec_per_clip=round(sec_clip, 4),
                             first_clip_s=round(A["first_clip_s"], 3), clips_timed=A["clips_timed"],
                             how="wall clock of DDIMSampler.sample(S=50, cfg 7.5, rescale 0.7, eta 1) per clip: pinned host "
                                 "conditioning uploaded inside the call, pred_x0 copied to the host every step "
                                 "(img_callback), final latent to the host; the first call (CUDA-graph capture) is "
                                 "first_clip_s and is not in the average"),
                    gpu_launches=A["launches"], clocks=A["clocks"], roofline=roofline, cpu_baseline=cpu_baseline,
                    whole_step_tflops=round(whole, 1), whole_step_frac_of_peak=round(whole / pk["bf16"], 4),
                    sec_per_50_step_clip=round(sec_clip, 4), single_clip_1024=single, torch_gpu_baseline=torch_gpu,
                    vae_decode=vae_leg,
                    breakdown=breakdown)
        print(json.dumps(line), flush=True)
    if world > 1:
        # every captured graph (they hold NCCL kernels of the clip's communicators) has been released by leg();
        # a watchdog only guards against a teardown that never returns
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        wd = threading.Timer(60.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
