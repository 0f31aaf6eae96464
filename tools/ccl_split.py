#!/usr/bin/env python3
"""Where ccl_final_kernel's time goes: labels + blobs, labels only, blobs only, against a plain fill of the label image.
    python tools/ccl_split.py > gpurun_out/r02_ccl_split.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402


def prof(ctx, fn, n):
    for _ in range(3):
        fn()
    ctx.profile(True)
    for _ in range(4):
        fn()
    p = ctx.profile_dump()
    ctx.profile(False)
    return {k: round(v["ms"] * 1e3 / 4 / n, 2) for k, v in sorted(p.items()) if k.startswith("ccl")}


def main():
    ctx = bv.Context(0)
    for name, h, w, n, gen in (("4K C5 frames", 2160, 3840, 8, lambda i: synth.gen_c5_frame(100 + i)),
                               ("1080p", 1080, 1920, 16, lambda i: synth.gen_underwater(1080, 1920, 10 + i))):
        frames = ctx.upload(np.stack([gen(i) for i in range(n)]))
        desc = ctx.make_stage(balance={} if h == 2160 else None, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)])
        mask = ctx.stage(desc, frames, want=("mask",))["mask"]
        ctx.sync()   # ctx.stage() only enqueues on the context's stream
        frac = float((mask != 0).float().mean())
        for want in (("labels", "blobs"), ("labels",), ("blobs",)):
            mb, wl = (4096 if "blobs" in want else 0), ("labels" in want)
            res = prof(ctx, lambda: ctx.label(mask, max_blobs=mb, want_labels=wl), n)
            print("%s (%.2f %% set): %-20s per-kernel us/frame %s" % (name, 100 * frac, "+".join(want), res), flush=True)
        lab = torch.empty((n, h, w), dtype=torch.int32, device="cuda")
        for _ in range(3):
            lab.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lab.zero_()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print("%s: plain zero fill of the int32 label image: %.2f us/frame (%.0f GB/s)" % (name, us / n, lab.numel() * 4 / us / 1e3), flush=True)
        del frames, mask, lab
    ctx.close()


if __name__ == "__main__":
    main()
