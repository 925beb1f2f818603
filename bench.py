#!/usr/bin/env python
"""bench.py — headline benchmark of the video-denoising hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is ONE classifier-free-guidance DDIM step of the DynamiCrafter U-Net at 16 x 320 x 512
(latent 4x16x40x64, context 333x1024): two U-Net forwards (cond / uncond) + the fused CFG / DDIM
update, bf16 compute, synthetic inputs and random-init weights of the real architecture
(1.44 B parameters).  One process per GPU; for N > 1 launch with torchrun (weak scaling: every rank
denoises its own clip, no data-path collective — SURVEY.md section 8e axis (2)).  Rank 0 prints ONE JSON line.

  value         steps/s, all ranks, inputs already resident in HBM, step replayed as a CUDA graph
  e2e           same metric through the public sampler objects with HOST (pinned) buffers: every step
                uploads x_t, noise, c_concat, both contexts and the timestep, refreshes the context
                K/V projections, runs the step and reads x_{t-1} back
  roofline      the dominant kernel family (tc_gemm_kernel: every Linear / conv3x3 / temporal conv),
                algorithmic FLOPs / CUDA-event time from an instrumented eager step run right after
                the timed region, against MEASURED_PEAKS.json (sustained bf16 figure)
  cpu_baseline  the oracle (CPU restatement of the reference, oracle/) on the host cores
  --impl reference  times the reference algorithm's CPU path (oracle port) on the host cores instead
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
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]

METRIC = "unet_cfg_denoise_steps_per_s_16x320x512"
S_DDIM, CFG_SCALE, RESCALE, ETA = 50, 7.5, 0.7, 1.0
FLOP_PER_FORWARD = {"512": 12602.83e9, "1024": 52336.20e9}


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


def build_native(res, device):
    from pandora_b200 import DDIMSampler, LatentDiffusionLite, UNetModel, config
    from pandora_b200.synthetic import randomize_module_, synth_inputs
    latent = config.LATENT_512 if res == "512" else config.LATENT_1024
    ucfg = config.UNET_512 if res == "512" else config.UNET_1024
    dcfg = config.DIFFUSION_512 if res == "512" else config.DIFFUSION_1024
    with torch.device(device):
        unet = UNetModel(**ucfg).eval()
    randomize_module_(unet, 1234)
    unet = unet.to(torch.bfloat16)                      # Open-Pandora: model.to(bf16) (gradio_app.py:128-129)
    model = LatentDiffusionLite(unet, **dcfg).to(device)
    sampler = DDIMSampler(model)
    sampler.make_schedule(S_DDIM, "uniform_trailing", ETA, verbose=False)
    inp = synth_inputs(1, latent, ucfg["context_dim"], seed=int(os.environ.get("RANK", 0)))
    return unet, model, sampler, inp, latent, ucfg


def cpu_oracle_forward_seconds(res, crop=None, threads=None):
    """One fp32 forward of the oracle on the host cores; crop=(h, w) restricts the latent."""
    from oracle import unet_oracle
    from pandora_b200 import config
    from pandora_b200.synthetic import synth_inputs, synth_state_dict
    ucfg = config.UNET_512
    latent = config.LATENT_512 if res == "512" else config.LATENT_1024
    if crop:
        latent = (latent[0], latent[1], crop[0], crop[1])
    torch.set_num_threads(threads or os.cpu_count())
    sd = synth_state_dict(unet_oracle.unet_param_shapes(ucfg), 1234)
    inp = synth_inputs(1, latent, ucfg["context_dim"], 0)
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([999])
    fs = torch.tensor([ucfg["default_fs"]])
    return sd, ucfg, x, t, inp["context"], fs, latent


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port; the reference is pure Python
    on torch CPU ops, which is exactly what the oracle restates) on all host threads.  Each step is a
    bounded sample: one fp32 U-Net forward of the full-width model on a 16 x (24 x 32) crop of the
    latent (3/10 of the 40 x 64 pixels); per-pixel cost is extrapolated linearly to the full latent
    (this favours the CPU: spatial attention grows quadratically) and doubled for the two CFG branches."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import unet_oracle
    sd, ucfg, x, t, ctx, fs, latent = cpu_oracle_forward_seconds("512", crop=(24, 32))
    frac = (24 * 32) / (40 * 64)
    for _ in range(max(1, min(args.warmup, 2))):
        unet_oracle.unet_forward(sd, ucfg, x, t, ctx, fs)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        unet_oracle.unet_forward(sd, ucfg, x, t, ctx, fs)
        ts.append(time.perf_counter() - t0)
    per_step = 2.0 * (sum(ts) / len(ts)) / frac
    v = 1.0 / per_step
    line = dict(impl="reference", metric=METRIC, value=v, unit="steps/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="DynamiCrafter 320x512 16-frame CFG DDIM step (2 U-Net forwards), CPU fp32",
                            latent=[4, 16, 40, 64], context=[333, 1024]),
                cpu_baseline=dict(value=v, unit="steps/s", cores=torch.get_num_threads(), kind="port",
                                  sample="per step: 1 fp32 oracle forward, full-width U-Net, latent crop 16x24x32 "
                                         "(3/10 of the pixels), scaled x10/3 x2 branches"),
                e2e=dict(value=v, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--res", default="512", choices=["512", "1024"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sp", type=int, default=1, help="sequence-parallel ways: ONE clip across this many ranks "
                    "(frames sharded for spatial layers, NCCL all-to-all to pixel shards for temporal layers)")
    ap.add_argument("--cfg-par", type=int, default=1, choices=[1, 2],
                    help="2 = the cond / uncond branches of a step on two ranks (one all-gather of v per step)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--torch-baseline", action="store_true",
                    help="also time the reference's op sequence (the oracle: stock torch ops -> cuDNN / cuBLAS / einsum "
                         "attention, bf16) on this GPU and report it as torch_gpu_baseline (informational)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get("PB_BENCH_WATCHDOG"):       # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PB_BENCH_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pandora_b200 import abi
    from pandora_b200.sampler import GraphedCFGStep
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi.load()

    unet, model, sampler, inp, latent, ucfg = build_native(args.res, dev)
    clip_ways = args.sp * args.cfg_par            # ranks that cooperate on one clip
    if world % clip_ways:
        raise SystemExit(f"world size {world} is not divisible by sp*cfg_par = {clip_ways}")
    n_clips = world // clip_ways                  # independent clips in flight (data-parallel replicas)
    if clip_ways > 1:
        from pandora_b200.dist import CFGParallel, SequenceParallel, make_groups
        from pandora_b200.synthetic import synth_inputs
        dp_index, cfg_group, sp_group, rep_group = make_groups(world, args.cfg_par, args.sp)
        inp = synth_inputs(1, latent, ucfg["context_dim"], seed=dp_index)      # the same clip on all its ranks
        if args.sp > 1:
            unet.sequence_parallel = SequenceParallel(sp_group)
        if args.cfg_par > 1:
            sampler.cfg_parallel = CFGParallel(cfg_group)
        sampler.replica_group = rep_group
    b = 1
    shape = (b, *latent)
    d = {k: v.to(dev) for k, v in inp.items()}
    cond = dict(c_crossattn=[d["context"]], c_concat=[d["c_concat"]])
    uc = dict(c_crossattn=[d["uc_context"]], c_concat=[d["c_concat"]])
    fs = torch.tensor([ucfg["default_fs"]], dtype=torch.long, device=dev)
    table, _ = sampler._step_table(False)
    g = GraphedCFGStep(sampler, unet, shape, cond, uc, fs, CFG_SCALE, RESCALE, 1.0, table)
    steps_ddpm = [int(s) for s in reversed(sampler.ddim_timesteps)]
    if args.no_graph:
        def one_step(i):
            g.ts.fill_(steps_ddpm[i % S_DDIM])
            g.noise.normal_()
            g._body()
        g._body(); g.step_index.zero_()
        launches_per_step = None
    else:
        g.capture()
        launches_per_step = g.launches_per_step + 3

        def one_step(i):
            g.step(steps_ddpm[i % S_DDIM])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n_warm, n):
        g.reset(d["x"])
        for i in range(n_warm):
            fn(i)
        barrier()
        l0 = abi.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            if (n_warm + i) % S_DDIM == 0:
                g.reset(d["x"])                        # a new clip: x_T, step index 0
            fn(n_warm + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, abi.launches - l0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, launches = timed(one_step, args.warmup, args.steps)
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = n_clips * 1e3 / ms_step

    # ---- e2e: public sampler objects, host (pinned) buffers, H2D/D2H inside the timed region
    host = {k: inp[k].pin_memory() for k in ("x", "c_concat", "context", "uc_context")}
    host["noise"] = torch.randn(shape).pin_memory()
    host["ts"] = torch.zeros(g.ts.shape[0], dtype=torch.long).pin_memory()
    host_out = torch.empty(shape).pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = host_out.numel() * host_out.element_size()

    def e2e_step(i):
        host["ts"].fill_(steps_ddpm[i % S_DDIM])
        g.x.copy_(host["x"], non_blocking=True)
        g.noise.copy_(host["noise"], non_blocking=True)
        g.ts.copy_(host["ts"], non_blocking=True)
        d["c_concat"].copy_(host["c_concat"], non_blocking=True)
        d["context"].copy_(host["context"], non_blocking=True)
        d["uc_context"].copy_(host["uc_context"], non_blocking=True)
        g.refresh_conditioning()                       # K/V projections of the freshly uploaded contexts
        if args.no_graph:
            g._body()
        else:
            g.graph.replay()
            abi.launches += g.launches_per_step
        host_out.copy_(g.x, non_blocking=True)
        if (i + 1) % S_DDIM == 0:
            g.step_index.zero_()

    g.step_index.zero_()
    e2e_ms, _ = timed(e2e_step, 2, max(3, min(args.steps, 10)))
    e2e_ms_step = e2e_ms / max(3, min(args.steps, 10))

    # ---- instrumented eager step: per-launch CUDA events -> breakdown + roofline of the GEMM family
    breakdown, roofline = None, None
    if rank != 0 and clip_ways > 1:               # the step contains collectives: every rank of the clip runs it
        g.step_index.zero_()
        g.ts.fill_(steps_ddpm[0])
        g._body()
        torch.cuda.synchronize()
    if rank == 0:
        g.step_index.zero_()
        g.ts.fill_(steps_ddpm[0])
        abi.profile_begin()
        g._body()
        agg = {}
        for it in abi.profile_end():
            a = agg.setdefault(it["name"], dict(n=0, ms=0.0, flops=0, bytes=0))
            a["n"] += 1; a["ms"] += it["ms"]; a["flops"] += it["flops"]; a["bytes"] += it["bytes"]
        tot = sum(a["ms"] for a in agg.values())
        breakdown = {k: dict(n=a["n"], ms=round(a["ms"], 3), share=round(a["ms"] / tot, 4),
                             tflops=round(a["flops"] / a["ms"] / 1e9, 1), gbs=round(a["bytes"] / a["ms"] / 1e6, 1))
                     for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        fam = [a for k, a in agg.items() if k.startswith("tc_gemm") or k.startswith("tc_conv")]
        pk = peaks()
        # DRAM traffic of the family's most expensive launch shape, from the committed `ncu --set full` capture
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if args.res == "512" and os.path.exists(tpath):
            t0 = json.load(open(tpath))["launches"][0]
            traffic = t0["dram_bytes"]
            traffic_note = dict(launch=t0["launch"], dram_bytes=t0["dram_bytes"], algorithmic_bytes=t0["algorithmic_bytes"],
                                source="profiles/r1_ncu_full_key_metrics.md row 1 (ncu --set full, per launch)")
        if fam:
            fl, ms, n = sum(a["flops"] for a in fam), sum(a["ms"] for a in fam), sum(a["n"] for a in fam)
            ach = fl / ms / 1e9
            roofline = dict(bound="tensor", kernel="tc_gemm_kernel (Linear + conv3x3 + temporal conv, tcgen05)",
                            achieved=round(ach, 1), peak=pk["bf16"], unit="TFLOP/s", frac=round(ach / pk["bf16"], 4),
                            traffic=traffic, traffic_of=traffic_note, launches_per_step=n, avg_launch_us=round(ms / n * 1e3, 2),
                            flops_per_launch=fl / n, share_of_step=round(ms / tot, 4), peak_source=pk["src"] + " sustained",
                            how="per-launch CUDA events on the launching stream, instrumented eager step after the timed region")
        whole = 2 * FLOP_PER_FORWARD[args.res] / (ms_step * 1e-3) / 1e12 / clip_ways     # per GPU
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import unet_oracle
        sd, ocfg, x, t, ctx, fs_c, _ = cpu_oracle_forward_seconds(args.res)
        t0 = time.perf_counter()
        unet_oracle.unet_forward(sd, ocfg, x, t, ctx, fs_c)
        sec = time.perf_counter() - t0
        cpu_baseline = dict(value=1.0 / (2 * sec), unit="steps/s", cores=torch.get_num_threads(), kind="port",
                            sample=f"1 fp32 oracle U-Net forward at the full latent ({sec:.1f} s, no warm-up); "
                                   "a CFG step = 2 forwards")
    torch_gpu = None
    if rank == 0 and world == 1 and args.torch_baseline:
        # what the unmodified reference does on this box: the same op sequence through stock PyTorch kernels
        from oracle import unet_oracle
        from pandora_b200 import config as _cfg
        from pandora_b200.synthetic import synth_state_dict
        del g
        torch.cuda.empty_cache()
        ocfg = _cfg.UNET_512
        sd = {k: v.to(dev, torch.bfloat16) for k, v in synth_state_dict(unet_oracle.unet_param_shapes(ocfg), 1234).items()}
        xo = torch.cat([d["x"], d["c_concat"]], 1).to(torch.bfloat16)
        to = torch.tensor([999], device=dev)
        ts_ms = []
        with torch.no_grad():
            for i in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                unet_oracle.unet_forward(sd, ocfg, xo, to, d["context"].to(torch.bfloat16), fs)
                e1.record()
                e1.synchronize()
                ts_ms.append(e0.elapsed_time(e1))
        fwd = min(ts_ms[1:])
        torch_gpu = dict(ms_per_forward=round(fwd, 2), steps_per_s=round(1e3 / (2 * fwd), 3),
                         note="oracle (the reference's op sequence as stock torch ops: cuDNN conv, cuBLAS GEMM, einsum "
                              "attention with a materialised score matrix), bf16, eager, one forward = half a CFG step")
        del sd
    if rank == 0:
        line = dict(metric=METRIC if args.res == "512" else METRIC.replace("320x512", "576x1024"), value=value, unit="steps/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak" if clip_ways == 1 else "strong",
                    vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload="DynamiCrafter 320x512 16-frame 50-step DDIM, CFG 7.5, guidance_rescale 0.7, "
                                         "eta 1.0, uniform_trailing; step = 2 U-Net forwards + fused CFG/DDIM update"
                                if args.res == "512" else "DynamiCrafter 576x1024 16-frame CFG DDIM step",
                                latent=list(latent), context=[333, 1024], clips_in_flight=n_clips,
                                parallelism=f"dp{n_clips}" + (f"xcfg{args.cfg_par}" if args.cfg_par > 1 else "")
                                + (f"xsp{args.sp}" if args.sp > 1 else ""),
                                params=1438854980, cuda_graph=not args.no_graph,
                                l2="working set per step (2.9 GB bf16 weights + activations) exceeds the 126 MB L2"),
                    e2e=dict(value=n_clips * 1e3 / e2e_ms_step, unit="steps/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=d2h, ms_per_step=e2e_ms_step),
                    gpu_launches=launches, clocks=clk, roofline=roofline, cpu_baseline=cpu_baseline,
                    whole_step_tflops=round(whole, 1), whole_step_frac_of_peak=round(whole / peaks()["bf16"], 4),
                    sec_per_50_step_clip=round(S_DDIM * ms_step * 1e-3, 3), breakdown=breakdown)
        if torch_gpu is not None:
            line["torch_gpu_baseline"] = torch_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        if clip_ways > 1:
            # the captured CUDA graph holds NCCL kernels of the clip's communicators: tearing the process group
            # down underneath it blocks forever; release the graph first, and do not let teardown hang the run
            g.graph = None
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
