#!/usr/bin/env python3
"""A-B timing of the binary-morphology chain (BV_OPT_MORPH_VARIANT): 0 register-rolling warps, 1 shared-memory tile
filled with plain loads (the r01 kernel), 2 the same tile filled by one TMA bulk copy.  Run on the GPU box:
    python tools/morph_variants.py > gpurun_out/r02_morph_variants.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

NAMES = {0: "register-rolling warps", 1: "smem tile, plain loads (r01)", 2: "smem tile, TMA bulk copy"}


def timed(ctx, fn, reps=30):
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
    k5 = np.ones((5, 5), np.uint8)
    for (h, w, n) in [(1242, 2208, 16), (1080, 1920, 16), (2160, 3840, 8), (480, 640, 64)]:
        masks = ctx.upload(np.stack([synth.mask_blobs(h, w, 40 + i, sigma=6.0) for i in range(min(n, 4))] * (n // min(n, 4)))[..., None])
        frames = ctx.upload(np.stack([synth.gen_underwater(h, w, 50 + i) for i in range(min(n, 4))] * (n // min(n, 4))))
        d_fused = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)])
        d_c3 = ctx.make_stage(cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
        d_oc = ctx.make_stage(cvt="bgr2lab", lo=(0, 150, 0), hi=(255, 255, 255), morph=[("open", 5, 5, 1), ("close", 5, 5, 1)])
        ref = None
        for v in (1, 2, 0):
            ctx.set_option("morph_variant", v)
            out = {}
            t_morph = timed(ctx, lambda: ctx.morph(masks, "open", k5))
            t_fused = timed(ctx, lambda: out.update(ctx.stage(d_fused, frames, want=("mask",), out=out)))
            o3 = {}
            t_c3 = timed(ctx, lambda: o3.update(ctx.stage(d_c3, frames, want=("mask", "labels", "blobs"), max_blobs=4096, out=o3)))
            o4 = {}
            t_oc = timed(ctx, lambda: o4.update(ctx.stage(d_oc, frames, want=("mask",), out=o4)))
            got = (ctx.download(out["mask"]), ctx.download(o4["mask"]), ctx.download(ctx.morph(masks, "open", k5)))
            if ref is None:
                ref = got
            same = all(np.array_equal(a, b) for a, b in zip(got, ref))
            print("%dx%d x%d  variant %d (%-30s): morph OPEN 5x5 on uint8 masks %7.2f us/frame | balance+HSV+inRange+OPEN %7.2f | "
                  "HSV+inRange+OPEN+label %7.2f | LAB+inRange+OPEN+CLOSE %7.2f us/frame | identical to variant 1: %s"
                  % (w, h, n, v, NAMES[v], t_morph / n, t_fused / n, t_c3 / n, t_oc / n, same), flush=True)
    ctx.set_option("morph_variant", 0)
    # per-kernel time of the chain itself inside the fused stage
    frames = ctx.upload(np.stack([synth.gen_underwater(1242, 2208, 50 + i) for i in range(16)]))
    d_fused = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)])
    for v in (1, 2, 0):
        ctx.set_option("morph_variant", v)
        out = {}
        for _ in range(3):
            out.update(ctx.stage(d_fused, frames, want=("mask",), out=out))
        ctx.profile(True)
        for _ in range(4):
            out.update(ctx.stage(d_fused, frames, want=("mask",), out=out))
        prof = ctx.profile_dump()
        ctx.profile(False)
        print("fused stage 16 x 2208x1242, variant %d, serialised per-kernel us/frame:" % v,
              {k: round(x["ms"] * 1e3 / 4 / 16, 2) for k, x in sorted(prof.items())}, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
