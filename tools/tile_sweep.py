"""Tile-width sweep of the tcgen05 GEMM family on the shapes of the batched forward (profiling build -DPB_TC_TUNING,
which honours PB_TC_FORCE_BN; the product library has no such switch).

    python tools/tile_sweep.py --build
    PANDORA_B200_LIB=open-pandora_b200/lib/tuning/libpandora_b200.so python tools/tile_sweep.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]

if "--build" in sys.argv:
    from pandora_b200.build import build_library
    print(build_library(defines=("PB_TC_TUNING",), variant="tuning"))
    sys.exit(0)

if "--child" not in sys.argv:
    res = {}
    for bn in (0, 64, 128, 160, 256, 320):
        env = dict(os.environ)
        if bn:
            env["PB_TC_FORCE_BN"] = str(bn)
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True).stdout
        for line in out.splitlines():
            if line.startswith("{"):
                d = json.loads(line)
                res.setdefault(d["shape"], {})[bn] = d["us"]
    print(f"{'shape':52s}" + "".join(f"{('auto' if b == 0 else b):>8}" for b in (0, 64, 128, 160, 256, 320)))
    for shape, r in res.items():
        print(f"{shape:52s}" + "".join(f"{r.get(b, float('nan')):8.1f}" for b in (0, 64, 128, 160, 256, 320)))
    sys.exit(0)

import torch  # noqa: E402
from pandora_b200 import abi  # noqa: E402
from bench_kernels import timeit  # noqa: E402

BF = torch.bfloat16
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(BF)
rf = lambda *s: torch.randn(*s, device=dev, generator=g)
FR = 32


def lin(M, N, K, out=torch.float32, res=True):
    x, w, c, b = rn(M, K), rn(N, K), torch.empty(M, N, device=dev, dtype=out), rf(N)
    r = rf(M, N) if res else None
    t = timeit(lambda: abi.gemm(x, w, c, bias=b, residual=r, tc=True), 7, flush)
    print(json.dumps(dict(shape=f"linear M{M} N{N} K{K} {'res ' if res else ''}{'f32' if out == torch.float32 else 'bf16'}",
                          us=t * 1e6)))


def conv(cin, cout, h, w_):
    M = FR * h * w_
    x, w, c, b, r = rn(M, cin), rn(cout, 9 * cin), torch.empty(M, cout, device=dev), rf(cout), rf(M, cout)
    t = timeit(lambda: abi.gemm(x, w, c, bias=b, residual=r, mode=abi.A_CONV3X3, conv=(cin, h, w_, h, w_, 1, 0), tc=True),
               7, flush)
    print(json.dumps(dict(shape=f"conv3x3 {cin}->{cout} @{h}x{w_} x{FR} res f32", us=t * 1e6)))


def convt(C, hw):
    M = FR * hw
    x, w, c, b = rn(M, C), rn(C, 3 * C), torch.empty(M, C, device=dev, dtype=BF), rf(C)
    t = timeit(lambda: abi.gemm(x, w, c, bias=b, mode=abi.A_CONVT3, convt=(C, 16, hw), tc=True), 7, flush)
    print(json.dumps(dict(shape=f"convT3 {C} @{hw} px x2 clips bf16", us=t * 1e6)))


for (M, C) in ((81920, 320), (20480, 640), (5120, 1280)):
    lin(M, C, C)
    lin(M, C, C, BF, False)
    lin(M, 3 * C, C, BF, False)
    lin(M, C, 4 * C)
lin(1280, 1280, 1280)
for cin, cout, h, w_ in ((320, 320, 40, 64), (640, 640, 20, 32), (1280, 640, 20, 32), (1280, 1280, 10, 16), (2560, 1280, 10, 16)):
    conv(cin, cout, h, w_)
for C, hw in ((320, 2560), (640, 640), (1280, 160)):
    convt(C, hw)
