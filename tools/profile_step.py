"""Per-shape breakdown of ONE U-Net forward as the sampler runs it (batch 2 = cond|uncond branches of a CFG
step): per-launch CUDA events through the C-ABI binding, aggregated by (kernel family, shape).

    python tools/profile_step.py [--res 512|1024] [--batch 2] [--top 60]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]
from pandora_b200 import UNetModel, abi, config  # noqa: E402
from pandora_b200.synthetic import randomize_module_, synth_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="512")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--top", type=int, default=80)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ncu", action="store_true", help="one warm-up forward + one forward, nothing else (for ncu)")
    a = ap.parse_args()
    latent = config.LATENT_512 if a.res == "512" else config.LATENT_1024
    with torch.device("cuda"):
        m = UNetModel(**config.UNET_512).eval()
    randomize_module_(m, 1234)
    m = m.to(torch.bfloat16)
    inp = {k: v.cuda() for k, v in synth_inputs(a.batch, latent, 1024, 0).items()}
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.full((a.batch,), 499, device="cuda")
    for _ in range(1 if a.ncu else 2):
        m(x, t, context=inp["context"])
    torch.cuda.synchronize()
    if a.ncu:                      # ncu --profile-from-start off: only the kernels of this forward are captured
        l0 = abi.launches
        torch.cuda.profiler.start()
        m(x, t, context=inp["context"])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(event="forward_done", launches=abi.launches - l0)))
        return
    abi.profile_begin()
    m(x, t, context=inp["context"])
    agg = {}
    for it in abi.profile_end():
        g = agg.setdefault((it["name"], it["detail"]), dict(n=0, ms=0.0, flops=0, bytes=0))
        g["n"] += 1; g["ms"] += it["ms"]; g["flops"] += it["flops"]; g["bytes"] += it["bytes"]
    tot = sum(g["ms"] for g in agg.values())
    rows = []
    for (k, d), g in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        rows.append(dict(kernel=k, shape=d, n=g["n"], ms=round(g["ms"], 3), share=round(g["ms"] / tot, 4),
                         us_per=round(1e3 * g["ms"] / g["n"], 1), tflops=round(g["flops"] / g["ms"] / 1e9, 1),
                         gbs=round(g["bytes"] / g["ms"] / 1e6, 1)))
    for r in rows[:a.top]:
        print(json.dumps(r))
    print(json.dumps(dict(kernel="TOTAL", ms=round(tot, 3), launches=sum(g["n"] for g in agg.values()))))
    if a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
