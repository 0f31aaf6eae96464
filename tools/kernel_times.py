#!/usr/bin/env python3
"""Per-kernel CUDA-event times (library profiler, serialised launches) of the main workloads.
    python tools/kernel_times.py [c2 fused c5 c3]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

ctx = bv.Context(0)
which = sys.argv[1:] or ["c2", "fused", "c5", "c3"]
for what in which:
    h, w, n = {"c2": (1242, 2208, 16), "fused": (1242, 2208, 16), "c5": (2160, 3840, 8), "c3": (1080, 1920, 16)}[what]
    base = np.stack([synth.gen_underwater(h, w, 3000 + i) for i in range(4)])
    frames = ctx.upload(np.concatenate([np.roll(base, 5 * k, axis=2) for k in range(n // 4)]))
    if what == "c2":
        desc, want, kw = ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",), {}
    elif what == "fused":
        desc, want, kw = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",), {}
    else:
        desc = ctx.make_stage(balance=({} if what == "c5" else None), cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255),
                              morph=[("open", 5, 5, 1)], label=True)
        want, kw = ("mask", "labels", "blobs"), {"max_blobs": 4096}
    out = {}
    for _ in range(3):
        out.update(ctx.stage(desc, frames, want=want, out=out, **kw))
    ctx.sync()
    ctx.profile(True)
    reps = 10
    for _ in range(reps):
        out.update(ctx.stage(desc, frames, want=want, out=out, **kw))
    prof = ctx.profile_dump()
    ctx.profile(False)
    tot = sum(v["ms"] for v in prof.values())
    print("%s: %d frames %dx%d, %.1f us/frame serialised" % (what, n, w, h, tot * 1e3 / reps / n))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("   %-28s %3d launches/step  %7.2f us/frame  %5.1f %%" % (k, v["launches"] // reps, v["ms"] * 1e3 / reps / n, 100 * v["ms"] / tot))
    del frames, out
