#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt>     per-kernel launch count / total / share
  python tools/ncu_summary.py full <report.ncu-rep> <out.txt>       key counters per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        tot.setdefault(k, [0, 0.0])
        tot[k][0] += 1
        tot[k][1] += v
    s = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        f.write(f"{'kernel':32s} {'launches':>8s} {'total ms':>12s} {'ms/launch':>11s} {'share':>7s}\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:32s} {n:8d} {ms:12.3f} {ms / n:11.4f} {100 * ms / s:6.2f}%\n")
        f.write(f"{'total':32s} {sum(v[0] for v in tot.values()):8d} {s:12.3f}\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        for d in data:
            f.write("\n" + d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "") + "\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"  {k:70s} {d[idx[k]]:>18s} {units[idx[k]]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
