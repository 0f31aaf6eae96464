#!/usr/bin/env python3
"""Sweep of BV_OPT_MORPH_WARPS (row strips per SM of the register-rolling morphology) on the stages that use it.
    python tools/morph_warps.py > gpurun_out/r02_morph_warps.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402


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
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ctx = bv.Context(0)
    cases = [
        ("fused 16x2208x1242", np.stack([synth.gen_underwater(1242, 2208, 10 + i) for i in range(16)]),
         dict(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",)),
        ("C3 16x1080p", np.stack([synth.gen_underwater(1080, 1920, 10 + i) for i in range(16)]),
         dict(cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True), ("mask", "labels", "blobs")),
        ("C5 8x4K", np.stack([synth.gen_c5_frame(100 + i) for i in range(8)]),
         dict(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True), ("mask", "labels", "blobs")),
        ("C1 64x480p", np.stack([synth.gen_underwater(480, 640, 10 + i) for i in range(64)]),
         dict(cvt="bgr2lab", lo=(0, 150, 0), hi=(255, 255, 255), morph=[("open", 5, 5, 1), ("close", 5, 5, 1)]), ("mask",)),
    ]
    for name, host, kw, want in cases:
        frames = ctx.upload(host)
        desc = ctx.make_stage(**kw)
        ref = None
        for wps in (8, 6, 4, 3, 2, 1, 8):
            ctx.set_option("morph_warps", wps)
            out = {}
            t = timed(ctx, lambda: out.update(ctx.stage(desc, frames, want=want, max_blobs=4096, out=out)))
            got = ctx.download(out["mask"])
            if ref is None:
                ref = got
            print("%-20s morph_warps %d: %8.1f us/step (%7.0f frames/s), mask identical: %s"
                  % (name, wps, t, frames.shape[0] / t * 1e6, np.array_equal(got, ref)), flush=True)
        del frames
    ctx.close()


if __name__ == "__main__":
    main()
