"""Kernel sweep (BASELINE.json config 5): time the C-ABI kernels on the exact shapes of one U-Net
forward with CUDA events; prints one line per shape with TFLOP/s or GB/s.

    python tools/bench_kernels.py [--res 512|1024] [--what gemm,conv,convt,attn,norm] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]
from pandora_b200 import abi  # noqa: E402

BF = torch.bfloat16


def timeit(fn, reps, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="512")
    ap.add_argument("--what", default="gemm,geglu,conv,convt,attn,norm")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--simt", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    H, W = (40, 64) if a.res == "512" else (72, 128)
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    tc = not a.simt
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF)
    what = a.what.split(",")
    lv = [(16 * H * W, 320), (16 * H * W // 4, 640), (16 * H * W // 16, 1280), (16 * H * W // 64, 1280)]
    if "gemm" in what:
        for (M, C) in lv[:3]:
            for (N, K, role) in [(C, C, "proj"), (3 * C, C, "qkv"), (C, 4 * C, "ff_out")]:
                x, w, c = rn(M, K), rn(N, K), torch.empty(M, N, device=dev)
                t = timeit(lambda: abi.gemm(x, w, c, tc=tc), a.reps, flush)
                rows.append(dict(kind="gemm", role=role, M=M, N=N, K=K, ms=t * 1e3, tflops=2 * M * N * K / t / 1e12))
    if "geglu" in what:
        for (M, C) in lv[:3]:
            x, w, c, b = rn(M, C), rn(8 * C, C), torch.empty(M, 4 * C, device=dev, dtype=BF), torch.zeros(8 * C, device=dev)
            t = timeit(lambda: abi.gemm(x, w, c, bias=b, geglu=True, tc=tc), a.reps, flush)
            rows.append(dict(kind="geglu", M=M, N=8 * C, K=C, ms=t * 1e3, tflops=2 * M * 8 * C * C / t / 1e12))
    if "conv" in what:
        for i, (cin, cout) in enumerate([(320, 320), (640, 640), (1280, 1280), (1280, 1280)]):
            h, w_ = H >> i, W >> i
            for ci in (cin, 2 * cin):
                M = 16 * h * w_
                x, w, c = rn(M, ci), rn(cout, 9 * ci), torch.empty(M, cout, device=dev)
                t = timeit(lambda: abi.gemm(x, w, c, mode=abi.A_CONV3X3, conv=(ci, h, w_, h, w_, 1, 0), tc=tc), a.reps, flush)
                rows.append(dict(kind="conv3x3", cin=ci, cout=cout, H=h, W=w_, ms=t * 1e3, tflops=2 * M * cout * 9 * ci / t / 1e12))
    if "convt" in what:
        for i, C in enumerate([320, 640, 1280, 1280]):
            hw = (H >> i) * (W >> i)
            M = 16 * hw
            x, w, c = rn(M, C), rn(C, 3 * C), torch.empty(M, C, device=dev)
            t = timeit(lambda: abi.gemm(x, w, c, mode=abi.A_CONVT3, convt=(C, 16, hw), tc=tc), a.reps, flush)
            rows.append(dict(kind="convt3", C=C, HW=hw, ms=t * 1e3, tflops=2 * M * C * 3 * C / t / 1e12))
    if "attn" in what:
        for i, C in enumerate([320, 640, 1280]):
            n, heads = (H >> i) * (W >> i), C // 64
            qkv = rn(16 * n, 3 * C)
            out = torch.empty(16 * n, C, device=dev, dtype=BF)
            ld = 3 * C
            op = lambda off: abi.operand(qkv, 1, n * ld, 0, ld, off)
            t = timeit(lambda: abi.attention(op(0), op(C), op(2 * C), out, (1, n * C, 0, C), 16, heads, n, n, 0.125,
                                             abi.BF16, tc=tc), a.reps, flush)
            rows.append(dict(kind="attn_self", n=n, heads=heads, ms=t * 1e3, tflops=4 * 16 * heads * n * n * 64 / t / 1e12))
            kv = rn(77, 2 * C)
            t = timeit(lambda: abi.attention(op(0), abi.operand(kv, 16, 77 * 2 * C, 0, 2 * C, 0),
                                             abi.operand(kv, 16, 77 * 2 * C, 0, 2 * C, C), out, (1, n * C, 0, C), 16, heads,
                                             n, 77, 0.125, abi.BF16, tc=tc), a.reps, flush)
            rows.append(dict(kind="attn_cross77", n=n, heads=heads, ms=t * 1e3, tflops=4 * 16 * heads * n * 77 * 64 / t / 1e12))
    if "norm" in what:
        for i, C in enumerate([320, 640, 1280, 1280]):
            hw = (H >> i) * (W >> i)
            x = torch.randn(16 * hw, C, device=dev)
            y = torch.empty(16 * hw, C, device=dev, dtype=BF)
            gmm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            for n_inst, tag in ((16, "frame"), (1, "clip")):
                t = timeit(lambda: abi.groupnorm(x, y, n_inst, 16 * hw // n_inst, gmm, bt, 1e-5, True), a.reps, flush)
                rows.append(dict(kind="groupnorm_" + tag, C=C, rows=16 * hw, ms=t * 1e3, gbs=(2 * x.numel() * 4 + y.numel() * 2) / t / 1e9))
            t = timeit(lambda: abi.layernorm(x, y, gmm, bt), a.reps, flush)
            rows.append(dict(kind="layernorm", C=C, rows=16 * hw, ms=t * 1e3, gbs=(x.numel() * 4 + y.numel() * 2) / t / 1e9))
    for r in rows:
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}))
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
