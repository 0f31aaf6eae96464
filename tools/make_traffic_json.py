#!/usr/bin/env python3
"""profiles/r02_traffic.json from the ncu captures of tools/ncu_traffic.sh (gpurun_out/r02_perkernel_*.csv: kernel replay,
caches kept; gpurun_out/r02_traffic_*_range.csv: range replay of one whole 16-frame step with side streams on)."""
import collections
import csv
import io
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
H, W, FRAMES = 1242, 2208, 16


def rows(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def val(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)


res = {"source": "ncu 2025 on B200, tools/ncu_traffic.sh + tools/traffic_run.py: one 16-frame step of 2208x1242 frames out of a 526 MB ring, "
                 "side streams on.  kernels.*: kernel replay with --cache-control none (kernels serialised, caches kept between them: "
                 "each chunk's passes run back to back, the best case for L2 residency).  steady_state.*: --replay-mode range over "
                 "the whole step (kernels concurrent as in production, caches untouched): what the step really moves through HBM.",
       "kernels": {}, "steady_state": {}}
for what in ("c2", "fused"):
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows(os.path.join(OUT, "r02_perkernel_%s.csv" % what)):
        k = r["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
        agg[k][r["Metric Name"]] += val(r)
        if r["Metric Name"] == "gpu__time_duration.sum":
            cnt[k] += 1
    for k, d in agg.items():
        n = cnt[k]
        res["kernels"].setdefault(k, {
            "workload": what, "launches_in_step": n, "frames_per_launch": FRAMES // n,
            "dram_bytes_per_launch": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n,
            "dram_read_bytes_per_launch": d["dram__bytes_read.sum"] / n, "dram_write_bytes_per_launch": d["dram__bytes_write.sum"] / n,
            "warp_instructions_per_launch": d["smsp__inst_executed.sum"] / n,
            "thread_instr_per_px": d["smsp__inst_executed.sum"] * 32 / (FRAMES * H * W),
            "lsu_instr_per_px": d["smsp__inst_executed_pipe_lsu.sum"] * 32 / (FRAMES * H * W),
            "alu_instr_per_px": d["smsp__inst_executed_pipe_alu.sum"] * 32 / (FRAMES * H * W),
            "fma_instr_per_px": d["smsp__inst_executed_pipe_fma.sum"] * 32 / (FRAMES * H * W),
            "smem_wavefronts_per_px": d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"] * 32 / (FRAMES * H * W),
            "ncu_us_per_launch": d["gpu__time_duration.sum"] / n / 1e3})
    m = {r["Metric Name"]: val(r) for r in rows(os.path.join(OUT, "r02_traffic_%s_range.csv" % what))}
    algo = (6 if what == "c2" else 4) * H * W
    res["steady_state"][what] = {
        "dram_read_bytes_per_frame": m["dram__bytes_read.sum"] / FRAMES, "dram_write_bytes_per_frame": m["dram__bytes_write.sum"] / FRAMES,
        "dram_bytes_per_frame": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / FRAMES,
        "algorithmic_bytes_per_frame": algo,
        "ratio_to_algorithmic": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / FRAMES / algo,
        "l2_bytes_per_frame": m["lts__t_bytes.sum"] / FRAMES, "thread_instr_per_px": m["smsp__inst_executed.sum"] * 32 / (FRAMES * H * W)}
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(res["steady_state"], indent=1))
for k, v in res["kernels"].items():
    print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()})
