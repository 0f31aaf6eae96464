#!/usr/bin/env python3
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_c2.csv          -> markdown table
    python tools/summarize_ncu.py full gpurun_out/prof_c2.ncu-rep              -> key metrics per kernel
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].replace("bv::", "")
        val = float(r[hdr.index("Metric Value")].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        print("| %s | %d | %.1f | %.2f | %.1f%% |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("| **total** | %d | %.1f | | 100%% |" % (sum(v[0] for v in agg.values()), tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    names = []
    cols = []
    for r in rows[2:]:
        names.append(r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("bv::", ""))
        cols.append(r)
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in cols:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3f" % f).rstrip("0").rstrip(".") if abs(f) < 1e6 else "%.3e" % f
            except ValueError:
                pass
            vals.append(v + (" " + units[i] if units[i] else ""))
        print("| %s | " % label + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
