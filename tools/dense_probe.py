#!/usr/bin/env python3
"""Worst case for the blob moments: a threshold that keeps most of a 3840x2160 frame (one giant
component plus holes).  Prints device-resident frames/s of balance -> HSV -> inRange -> OPEN -> label."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

ctx = bv.Context(0)
ring = ctx.upload(np.stack([synth.gen_underwater(2160, 3840, 3200 + i) for i in range(8)]))
for name, lo, hi in (("dense", (0, 40, 60), (179, 255, 255)), ("bins", (10, 20, 60), (30, 100, 255))):
    desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=lo, hi=hi, morph=[("open", 5, 5, 1)], label=True)
    out = {}
    for _ in range(3):
        out.update(ctx.stage(desc, ring, want=("mask", "labels", "blobs"), max_blobs=8192, out=out))
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    for _ in range(10):
        out.update(ctx.stage(desc, ring, want=("mask", "labels", "blobs"), max_blobs=8192, out=out))
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    ctx.sync()
    ms = e0.elapsed_time(e1) / 10
    n, tabs = ctx.blobs_to_numpy(out["blobs"], out["n_blobs"])
    print("%s: %.3f ms per 8 frames = %.0f frames/s; blobs per frame %s; largest %d px" %
          (name, ms, 8e3 / ms, n.tolist()[:3], max(int(t["m00"].max()) if len(t) else 0 for t in tabs)))
