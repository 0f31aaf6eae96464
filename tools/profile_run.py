#!/usr/bin/env python3
"""Short, profiler-friendly run of one workload (used under ncu; numbers printed here are never
bench values).  python tools/profile_run.py [c2|fused|c3|c5|c4] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic input generator only)

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = bv.Context(0)
if what in ("c2", "fused"):
    frames = ctx.upload(np.stack([synth.gen_underwater(1242, 2208, 10 + i) for i in range(8)]))
    if what == "c2":
        desc, want = ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",)
    else:
        desc, want = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",)
    out = {}
    for _ in range(steps):
        out.update(ctx.stage(desc, frames, want=want, out=out))
elif what in ("c3", "c5"):
    h, w = (1080, 1920) if what == "c3" else (2160, 3840)
    frames = ctx.upload(np.stack([synth.gen_underwater(h, w, 10 + i) for i in range(4)]))
    desc = ctx.make_stage(balance=({} if what == "c5" else None), cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255),
                          morph=[("open", 5, 5, 1)], label=True)
    out = {}
    for _ in range(steps):
        out.update(ctx.stage(desc, frames, want=("mask", "labels", "blobs"), max_blobs=4096, out=out))
else:
    imgs = [ctx.upload(synth.gen_underwater(1242, 2208, 10 + i)) for i in range(16)]
    for _ in range(steps):
        ctx.letterbox(imgs)
ctx.sync()
print("done", what, steps, "launches", ctx.launches)
