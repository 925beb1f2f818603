"""Time one full-size UNet forward (and optionally compare modes) on the GPU; prints JSON lines."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]
from pandora_b200 import UNetModel, abi, config  # noqa: E402
from pandora_b200.synthetic import randomize_module_, synth_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="512")
    ap.add_argument("--modes", default="bf16-tc,bf16-simt,fp32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--compare", action="store_true")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    latent = config.LATENT_512 if a.res == "512" else config.LATENT_1024
    t0 = time.time()
    with torch.device("cuda"):
        m = UNetModel(**config.UNET_512).eval()
    randomize_module_(m, 1234)
    print(json.dumps(dict(event="model_ready", s=round(time.time() - t0, 1))), flush=True)
    inp = {k: v.cuda() for k, v in synth_inputs(1, latent, 1024, 0).items()}
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([999], device="cuda")
    outs = {}
    for mode in a.modes.split(","):
        m.compute_mode = "fp32" if mode == "fp32" else "bf16"
        m.kernel_impl = "simt" if mode.endswith("simt") else "auto"
        m.invalidate_packed_weights()
        torch.cuda.synchronize()
        l0 = abi.launches
        y = m(x, t, context=inp["context"])
        torch.cuda.synchronize()
        nl = abi.launches - l0
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = m(x, t, context=inp["context"])
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        outs[mode] = y.float()
        if a.profile:
            abi.profile_begin()
            m(x, t, context=inp["context"])
            agg = {}
            for it in abi.profile_end():
                g = agg.setdefault(it["name"], dict(n=0, ms=0.0, flops=0, bytes=0))
                g["n"] += 1; g["ms"] += it["ms"]; g["flops"] += it["flops"]; g["bytes"] += it["bytes"]
            tot = sum(g["ms"] for g in agg.values())
            for k, g in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
                print(json.dumps(dict(mode=mode, kernel=k, n=g["n"], ms=round(g["ms"], 3), share=round(g["ms"] / tot, 3),
                                      tflops=round(g["flops"] / g["ms"] / 1e9, 1), gbs=round(g["bytes"] / g["ms"] / 1e6, 1))))
            print(json.dumps(dict(mode=mode, kernel="TOTAL(sum of launches)", ms=round(tot, 3))), flush=True)
        print(json.dumps(dict(mode=mode, res=a.res, ms=[round(v, 2) for v in ts], launches=nl,
                              finite=bool(torch.isfinite(y).all()),
                              mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))), flush=True)
    if a.compare and "fp32" in outs:
        ref = outs["fp32"].double()
        for k, v in outs.items():
            if k != "fp32":
                print(json.dumps(dict(compare=k, rel_l2_vs_fp32=float((v.double() - ref).norm() / ref.norm()))))


if __name__ == "__main__":
    main()
