"""Key metrics of every launch in an `.ncu-rep` (from `ncu --set full`) as a markdown table + a JSON list.

    python tools/ncu_key_metrics.py gpurun_out/r2c_targets.ncu-rep [labels.json] > profiles/<name>.md

`labels.json` (optional): [{"launch": "...", "algorithmic_bytes": N, "flops": N}, ...] in launch order; adds the label,
the algorithmic bytes and the derived GB/s / TFLOP/s columns.  Writes `<out>.json` next to stdout when `--json PATH`.
"""
import csv
import io
import json
import subprocess
import sys

COLS = {
    "time_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = csv.reader(io.StringIO(out))
    head, units = next(rd), next(rd)
    find = lambda m: next((i for i, h in enumerate(head) if h == m), None) or next(
        (i for i, h in enumerate(head) if h.endswith("." + m)), None)
    idx = {k: find(m) for k, m in COLS.items()}
    ik, ig, ib = head.index("Kernel Name"), head.index("Grid Size"), head.index("Block Size")
    rows = []
    for r in rd:
        d = dict(kernel=r[ik].split("(")[0].replace("void ", ""), grid=r[ig], block=r[ib])
        for k, i in idx.items():
            if i is None or r[i] == "":
                d[k] = None
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                d[k] = None
                continue
            u = units[i]
            if k == "time_us":
                v *= TO_US.get(u, 1)
            elif k in ("dram_rd", "dram_wr"):
                v *= TO_BYTES.get(u, 1)
            d[k] = v
        rows.append(d)
    return rows


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep = args[0]
    labels = json.load(open(args[1])) if len(args) > 1 else []
    rows = read(rep)
    print("| # | kernel | launch | grid x block | time us | DRAM read MB | DRAM write MB | algorithmic MB | "
          "DRAM/alg | GB/s (alg) | TFLOP/s | L2 % | DRAM % | issue % | warps % | regs |")
    print("|" + "---|" * 16)
    f = lambda v, p=1: "-" if v is None else f"{v:.{p}f}"
    for i, d in enumerate(rows):
        lab = labels[i] if i < len(labels) else {}
        alg, fl = lab.get("algorithmic_bytes"), lab.get("flops")
        dram = (d["dram_rd"] or 0) + (d["dram_wr"] or 0)
        d.update(launch=lab.get("launch", ""), algorithmic_bytes=alg, flops=fl, dram_bytes=dram)
        print(f"| {i + 1} | `{d['kernel']}` | {d['launch']} | {d['grid']} x {d['block']} | {f(d['time_us'])} | "
              f"{f(d['dram_rd'] and d['dram_rd'] / 1e6)} | {f(d['dram_wr'] and d['dram_wr'] / 1e6)} | "
              f"{f(alg and alg / 1e6)} | {f(alg and dram / alg, 2)} | {f(alg and alg / d['time_us'] / 1e3, 0)} | "
              f"{f(fl and fl / d['time_us'] / 1e6, 0)} | {f(d['lts_pct'])} | {f(d['dram_pct'])} | "
              f"{f(d['issue_pct'])} | {f(d['warps_pct'])} | {f(d['regs'], 0)} |")
    for a in sys.argv[1:]:
        if a.startswith("--json="):
            json.dump(dict(source=rep, launches=rows), open(a.split("=", 1)[1], "w"), indent=1)


if __name__ == "__main__":
    main()
