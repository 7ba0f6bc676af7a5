#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (text, committed).

  python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
  python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep           > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi:
            d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print(f"# {path}: {sum(len(v) for v in d.values())} launches, total {tot/1e6:.3f} ms (gpu__time_duration.sum, "
          "cold-cache + serialised: compare shares)")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{100*sum(v)/tot:6.2f}%  n={len(v):5d}  mean={sum(v)/len(v)/1e3:9.2f} us  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    print(f"# {path}: ncu --set full, {len(rows)-2} launch(es)")
    for r in rows[2:]:
        print("kernel:", r[H.index("Kernel Name")][:140])
        for k in KEYS:
            if k in H:
                i = H.index(k)
                print(f"  {k:72s} {r[i]:>18s} {U[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
