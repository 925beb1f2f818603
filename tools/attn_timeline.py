"""Per-phase timeline of ONE CTA of the multi-tile self-attention kernel (profiling build, -DPB_ATT_TRACE).

    python tools/attn_timeline.py --build          # here (nvcc): lib/trace/libpandora_b200.so
    PANDORA_B200_LIB=open-pandora_b200/lib/trace/libpandora_b200.so python tools/attn_timeline.py [--n 2560]

Prints, for the softmax warp 0 of query tile 0 and 1 of CTA (0,0,0): mean SM-clock cycles per phase of an iteration
    wait S | tcgen05.ld S | row max | wait turn (ping-pong) | exponentials | wait PV(j-1) | store P + signal
and the issuer's issue times relative to the events it reacts to.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]

if "--build" in sys.argv:
    from pandora_b200.build import build_library
    print(build_library(defines=("PB_ATT_TRACE",), variant="trace"))
    sys.exit(0)

import torch  # noqa: E402
from pandora_b200 import abi  # noqa: E402

n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 2560
heads, B, MAXE = 5, 32, 4096
C = heads * 64
lib = abi.load()
lib.pb_debug_attn_trace.restype, lib.pb_debug_attn_trace.argtypes = ctypes.c_int, [ctypes.c_void_p]
buf = torch.zeros(3, MAXE, dtype=torch.int64, device="cuda")
qkv = torch.randn(B * n, 3 * C, device="cuda").to(torch.bfloat16)
o = torch.empty(B * n, C, device="cuda", dtype=torch.bfloat16)
ld = 3 * C
op = lambda off: abi.operand(qkv, 1, n * ld, 0, ld, off)
run = lambda: abi.attention(op(0), op(C), op(2 * C), o, (1, n * C, 0, C), B, heads, n, n, 0.125, abi.BF16, False, tc=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
assert lib.pb_debug_attn_trace(buf.data_ptr()) == 0
run()
torch.cuda.synchronize()
lib.pb_debug_attn_trace(None)
ev = [[(int(v) >> 8, int(v) & 255) for v in row.tolist() if v] for row in buf.cpu()]
t0 = min(e[0][0] for e in ev if e)
names = {10: "loop top", 11: "S landed", 12: "S in registers", 13: "row max done", 14: "turn granted", 15: "exp done",
         16: "PV(j-1) seen", 17: "P stored"}
phases = [("wait S", 10, 11), ("tcgen05.ld S", 11, 12), ("row max", 12, 13), ("wait turn", 13, 14),
          ("exponentials", 14, 15), ("wait PV(j-1)", 15, 16), ("store P + signal", 16, 17)]
for t in (0, 1):
    e = ev[1 + t]
    its, cur = [], {}
    for c, tag in e:
        cur[tag] = c
        if tag == 17:
            its.append(cur)
            cur = {}
    print(f"query tile {t}: {len(its)} key tiles, first at +{its[0][10] - t0} clk, last P at +{its[-1][17] - t0} clk, "
          f"mean period {(its[-1][17] - its[0][17]) / max(1, len(its) - 1):.0f} clk")
    mid = its[2:-1]
    for name, a, b in phases:
        d = [i.get(b, i[15 if b == 16 else b]) - i[a] for i in mid if a in i]
        print(f"    {name:18s} mean {sum(d) / len(d):7.0f}  max {max(d):6d}")
iss = ev[0]
kinds = {0: "Q landed", 1: "QK tile0", 2: "QK tile1", 3: "PV tile0", 4: "PV tile1", 5: "QK ready seen", 6: "PV ready seen"}
print("issuer (first 24 events):", ", ".join(f"{kinds[k]}@{c - t0}" for c, k in iss[:24] if k < 5))
for t in (0, 1):
    sm = [c for c, k in ev[1 + t] if k == 17]           # P stored, iteration j
    pv = [c for c, k in iss if k == 3 + t]
    sf = [c for c, k in ev[1 + t] if k == 12]           # S in registers (s_free)
    qk = [c for c, k in iss if k == 1 + t]
    d1 = [b - a for a, b in zip(sm, pv)]
    d2 = [b - a for a, b in zip(sf, qk[1:])]
    print(f"tile {t}: PV issue lags 'P stored' by mean {sum(d1) / len(d1):.0f} clk (max {max(d1)}); "
          f"next QK issue lags 's_free' by mean {sum(d2) / max(1, len(d2)):.0f} clk (max {max(d2) if d2 else 0})")

if "--raw" in sys.argv:
    merged = sorted([(c - t0, "issuer", kinds[k]) for c, k in iss] +
                    [(c - t0, f"softmax{t}", names[k]) for t in (0, 1) for c, k in ev[1 + t]])
    lo = merged[len(merged) // 2][0]
    for c, who, what in merged:
        if lo <= c < lo + 9000:
            print(f"{c:8d}  {who:9s} {what}")
