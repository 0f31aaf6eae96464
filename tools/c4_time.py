#!/usr/bin/env python3
"""C4 timing: 16 x 2208x1242 (and a mixed-camera batch) -> 640x640 fp16 letterbox, CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

ctx = bv.Context(0)
sets = {"16x2208x1242": [(1242, 2208)] * 16,
        "mixed 6x2208x1242 + 5x1920x1080 + 5x1280x720": [(1242, 2208)] * 6 + [(1080, 1920)] * 5 + [(720, 1280)] * 5}
for name, shapes in sets.items():
    imgs = [ctx.upload(synth.gen_underwater(h, w, 10 + i)) for i, (h, w) in enumerate(shapes)]
    for _ in range(5):
        ctx.letterbox(imgs)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    with torch.cuda.stream(ctx.torch_stream):
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            ctx.letterbox(imgs)
        e1.record(ctx.torch_stream)
    ctx.sync()
    us = e0.elapsed_time(e1) * 1e3 / reps
    src = sum(h * w * 3 for h, w in shapes)
    alg = src + len(shapes) * 3 * 640 * 640 * 2
    ctx.profile(True)
    for _ in range(10):
        ctx.letterbox(imgs)
    prof = ctx.profile_dump()
    ctx.profile(False)
    print("   kernels:", {k: round(v["ms"] * 1e3 / v["launches"], 2) for k, v in prof.items()}, "us/launch")
    print("%-50s %7.1f us/batch  %8.0f img/s  algorithmic %.0f GB/s" % (name, us, len(shapes) / us * 1e6, alg / us / 1e3))
