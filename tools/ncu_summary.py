#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run in the build container; the .ncu-rep files come back in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv  > profiles/X_launches.md
    python tools/ncu_summary.py full     gpurun_out/X_prof.ncu-rep  > profiles/X_ncu_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        key = (name, d[idx["launch__grid_size"]])
        if key in seen:
            continue
        seen.add(key)
        print(f"### `{name}`  grid {d[idx['launch__grid_size']]} x {d[idx['launch__block_size']]}\n")
        for k in KEYS:
            if k in idx:
                print(f"- {k} = {d[idx[k]]} {units[idx[k]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
