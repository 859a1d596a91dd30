#!/usr/bin/env python
"""Summarises ncu output brought back in gpurun_out/ (runs here, no GPU).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep

`launches`: per-kernel count / mean duration / share of the listed launches
            (the `--metrics gpu__time_duration.sum` pass).
`full`    : per-launch DRAM bytes, DRAM %, occupancy, registers, top stall
            reasons out of an `--set full` report (needs `ncu` on PATH).
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(unit, 1.0)
        agg.setdefault(name, []).append(us)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | min us | max us | share of listed time |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {100 * sum(v) / tot:.1f} % |")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        print(f"\n### `{name}`\n")
        print("| metric | value |")
        print("|---|---:|")
        for key, label in WANT:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
