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
    ap.add_argument("--frames", type=int, default=16,
                    help="frames in the GEMM-family rows (16 = one clip; 32 = the batched cond|uncond forward)")
    ap.add_argument("--simt", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--torch", action="store_true",
                    help="also time the kernels the reference's torch ops dispatch to on this box (cuDNN conv2d / conv3d "
                         "channels-last bf16, F.scaled_dot_product_attention as the xformers stand-in, einsum attention, "
                         "F.group_norm + F.silu) on the same shapes: the `torch_*` rows")
    a = ap.parse_args()
    H, W = (40, 64) if a.res == "512" else (72, 128)
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    tc = not a.simt
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF)
    what = a.what.split(",")
    FR = a.frames
    lv = [(FR * H * W, 320), (FR * H * W // 4, 640), (FR * H * W // 16, 1280), (FR * H * W // 64, 1280)]
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
                M = FR * h * w_
                x, w, c = rn(M, ci), rn(cout, 9 * ci), torch.empty(M, cout, device=dev)
                t = timeit(lambda: abi.gemm(x, w, c, mode=abi.A_CONV3X3, conv=(ci, h, w_, h, w_, 1, 0), tc=tc), a.reps, flush)
                rows.append(dict(kind="conv3x3", cin=ci, cout=cout, H=h, W=w_, ms=t * 1e3, tflops=2 * M * cout * 9 * ci / t / 1e12))
    if "convt" in what:
        for i, C in enumerate([320, 640, 1280, 1280]):
            hw = (H >> i) * (W >> i)
            M = FR * hw
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
    if a.torch:
        import torch.nn.functional as F
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            if "conv" in what:
                for i, (cin, cout) in enumerate([(320, 320), (640, 640), (1280, 1280), (1280, 1280)]):
                    h, w_ = H >> i, W >> i
                    x = rn(16, cin, h, w_).contiguous(memory_format=torch.channels_last)
                    w = rn(cout, cin, 3, 3).contiguous(memory_format=torch.channels_last)
                    bias = rn(cout)
                    t = timeit(lambda: F.conv2d(x, w, bias, padding=1), a.reps, flush)
                    rows.append(dict(kind="torch_conv3x3_cudnn", cin=cin, cout=cout, H=h, W=w_, ms=t * 1e3,
                                     tflops=2 * 16 * h * w_ * cout * 9 * cin / t / 1e12))
            if "convt" in what:
                for i, C in enumerate([320, 640, 1280, 1280]):
                    h, w_ = H >> i, W >> i
                    x = rn(1, C, 16, h, w_).contiguous(memory_format=torch.channels_last_3d)
                    w = rn(C, C, 3, 1, 1).contiguous(memory_format=torch.channels_last_3d)
                    bias = rn(C)
                    t = timeit(lambda: F.conv3d(x, w, bias, padding=(1, 0, 0)), a.reps, flush)
                    rows.append(dict(kind="torch_convt3_cudnn", C=C, HW=h * w_, ms=t * 1e3,
                                     tflops=2 * 16 * h * w_ * C * 3 * C / t / 1e12))
            if "attn" in what:
                for i, C in enumerate([320, 640, 1280]):
                    n, heads = (H >> i) * (W >> i), C // 64
                    q, k, v = (rn(16, heads, n, 64) for _ in range(3))
                    t = timeit(lambda: F.scaled_dot_product_attention(q, k, v), a.reps, flush)
                    rows.append(dict(kind="torch_attn_self_sdpa", n=n, heads=heads, ms=t * 1e3,
                                     tflops=4 * 16 * heads * n * n * 64 / t / 1e12))

                    def einsum_attn():      # the path the reference takes without xformers (attention.py:101-125)
                        sim = torch.einsum("bhid,bhjd->bhij", q, k) * 0.125
                        return torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
                    t = timeit(einsum_attn, a.reps, flush)
                    rows.append(dict(kind="torch_attn_self_einsum", n=n, heads=heads, ms=t * 1e3,
                                     tflops=4 * 16 * heads * n * n * 64 / t / 1e12))
                for C, hw in ((320, H * W), (640, H * W // 4), (1280, H * W // 16)):
                    heads = C // 64
                    q, k, v = (rn(hw, heads, 16, 64) for _ in range(3))       # temporal: batch = pixels, seq 16
                    t = timeit(lambda: F.scaled_dot_product_attention(q, k, v), a.reps, flush)
                    rows.append(dict(kind="torch_attn_temporal_sdpa", pixels=hw, heads=heads, ms=t * 1e3,
                                     gbs=4 * hw * 16 * C * 2 / t / 1e9))
            if "norm" in what:
                for i, C in enumerate([320, 640, 1280, 1280]):
                    h, w_ = H >> i, W >> i
                    x = rn(16, C, h, w_).contiguous(memory_format=torch.channels_last)
                    gmm, bt = torch.ones(C, device=dev, dtype=BF), torch.zeros(C, device=dev, dtype=BF)
                    t = timeit(lambda: F.silu(F.group_norm(x, 32, gmm, bt, 1e-5)), a.reps, flush)
                    rows.append(dict(kind="torch_groupnorm_silu", C=C, rows=16 * h * w_, ms=t * 1e3,
                                     gbs=(4 * x.numel() * 2) / t / 1e9))
                    xr = rn(16 * h * w_, C)
                    t = timeit(lambda: F.layer_norm(xr, (C,), gmm, bt), a.reps, flush)
                    rows.append(dict(kind="torch_layernorm", C=C, rows=16 * h * w_, ms=t * 1e3,
                                     gbs=(2 * xr.numel() * 2) / t / 1e9))
    if "tattn" in what or "attn" in what:
        for C, hw in ((320, H * W), (640, H * W // 4), (1280, H * W // 16)):
            heads, T = C // 64, 16
            qkv = rn(T * hw, 3 * C)
            out = torch.empty(T * hw, C, device=dev, dtype=BF)
            ld = 3 * C
            op = lambda off: abi.operand(qkv, hw, T * hw * ld, ld, hw * ld, off)
            t = timeit(lambda: abi.attention(op(0), op(C), op(2 * C), out, (hw, T * hw * C, C, hw * C), hw, heads, T, T,
                                             0.125, abi.BF16), a.reps, flush)
            rows.append(dict(kind="attn_temporal", pixels=hw, heads=heads, ms=t * 1e3, gbs=4 * hw * 16 * C * 2 / t / 1e9))
    for r in rows:
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}))
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
