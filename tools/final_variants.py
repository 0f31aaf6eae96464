#!/usr/bin/env python3
"""A-B timing of pass 3 of balance -> BGR2LAB (BV_OPT_FINAL_SV_TABLES): 0 byte S'/V' stretch tables (int -> float per pixel),
1 float32 s / v tables, 2 8-byte {s, 1-s} / {v, trunc(255 v)} tables.  Checks that the three produce identical bytes.
    python tools/final_variants.py > gpurun_out/r02_final_variants.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

NAMES = {0: "byte tables (default)", 1: "float32 s, v", 2: "{s,1-s} / {v,trunc(255v)}"}


def timed(ctx, fn, reps=40):
    for _ in range(5):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    for _ in range(reps):
        fn()
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    ctx.sync()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def main():
    ctx = bv.Context(0)
    h, w, n = 1242, 2208, 16
    ring = [ctx.upload(np.stack([synth.gen_underwater(h, w, 100 * r + i) for i in range(n)])) for r in range(4)]  # 4 x 131 MB > L2
    desc = ctx.make_stage(balance={}, cvt="bgr2lab")
    order = (0, 1, 2, 0, 1, 2)
    ref = None
    for v in order:
        ctx.set_option("final_sv_tables", v)
        out = {}
        state = {"i": 0}

        def step():
            out.update(ctx.stage(desc, ring[state["i"] % 4], want=("converted",), out=out))
            state["i"] += 1
        t = timed(ctx, step)
        out.update(ctx.stage(desc, ring[0], want=("converted",), out=out))
        got = ctx.download(out["converted"])
        if ref is None:
            ref = got
        for _ in range(3):
            step()
        ctx.profile(True)
        for _ in range(4):
            step()
        prof = ctx.profile_dump()
        ctx.profile(False)
        print("C2 16 x 2208x1242  variant %d (%-26s): %7.2f us/frame = %6.0f frames/s | per-kernel us/launch %s | identical to variant 0: %s"
              % (v, NAMES[v], t / n, n / t * 1e6, {k: round(x["ms"] * 1e3 / x["launches"], 2) for k, x in sorted(prof.items())},
                 np.array_equal(got, ref)), flush=True)
    # odd width: the row tail goes through the scalar rounding path, the whole groups through the tables
    frames = ctx.upload(np.stack([synth.gen_underwater(484, 656, 7 + i) for i in range(4)]))
    ref = None
    for v in (0, 1, 2):
        ctx.set_option("final_sv_tables", v)
        got = ctx.download(ctx.stage(desc, frames, want=("converted",))["converted"])
        if ref is None:
            ref = got
        print("656x484 (width % 32 = 16: row tails take the scalar-rounding path, whole groups the tables): variant", v, "identical:", np.array_equal(got, ref), flush=True)
    ctx.set_option("final_sv_tables", 0)
    ctx.close()


if __name__ == "__main__":
    main()
