#!/usr/bin/env python3
"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof_gemm.ncu-rep profiles/r01_gemm_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("void ", "").replace("knn::<unnamed>::", "")[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{n}` | {c} | {t / 1e3:.3f} | {100 * t / tot:.2f}% |\n")
        f.write(f"| total | {sum(a[0] for a in agg.values())} | {tot / 1e3:.3f} | 100% |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(dst, "w") as f:
        f.write("| # | kernel | " + " | ".join(c.replace("|", "/") for c in cols) + " |\n")
        f.write("|---|---|" + "---:|" * len(cols) + "\n")
        f.write("| | unit | " + " | ".join(units[idx[c]] for c in cols) + " |\n")
        for n, r in enumerate(data):
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")[-48:]
            f.write(f"| {n} | `{name}` | " + " | ".join(r[idx[c]] for c in cols) + " |\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
