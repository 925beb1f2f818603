"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        agg[name][0] += 1
        agg[name][1] += v_us
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {n} | {us / 1e3:.3f} | {us / tot:.1%} | {us / n:.1f} |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.3f} | | |")


if __name__ == "__main__":
    main(sys.argv[1])
