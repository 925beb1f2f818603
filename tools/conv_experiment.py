"""Tile-shape experiment: time representative conv / GEMM shapes (batched forward sizes) under the current
PB_TC_PAIR / PB_TC_BN environment.  Prints one JSON line per shape."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]
from pandora_b200 import abi
BF = torch.bfloat16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3
B = 32
tag = dict(pair=os.environ.get("PB_TC_PAIR", "1"), bn=os.environ.get("PB_TC_BN", "auto"))
for (cin, cout, H, W) in [(320, 320, 40, 64), (640, 640, 20, 32), (1280, 1280, 10, 16), (2560, 1280, 10, 16)]:
    M = B * H * W
    x, w, c = rn(M, cin), rn(cout, 9 * cin), torch.empty(M, cout, device=dev)
    t = timeit(lambda: abi.gemm(x, w, c, mode=abi.A_CONV3X3, conv=(cin, H, W, H, W, 1, 0), tc=True))
    print(json.dumps(dict(tag, kind="conv3x3", cin=cin, cout=cout, H=H, us=round(t * 1e6, 1), tflops=round(2 * M * cout * 9 * cin / t / 1e12, 1))))
for (M, N, K) in [(81920, 320, 1280), (20480, 640, 2560), (5120, 1280, 5120), (5120, 3840, 1280)]:
    x, w, c = rn(M, K), rn(N, K), torch.empty(M, N, device=dev, dtype=BF)
    t = timeit(lambda: abi.gemm(x, w, c, tc=True))
    print(json.dumps(dict(tag, kind="gemm", M=M, N=N, K=K, us=round(t * 1e6, 1), tflops=round(2 * M * N * K / t / 1e12, 1))))
