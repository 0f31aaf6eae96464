#!/usr/bin/env python3
"""Per-kernel CUDA-event times (bv_profile, kernels serialised by the events' own ordering on each stream) of the side
workloads: C3 (1080p HSV inRange -> OPEN -> labels + moments), C5 (4K balance -> ... -> labels), C1, the fused mask stage.
    python tools/stage_profile.py > gpurun_out/r02_stage_profile.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402


def run(ctx, name, desc, frames, want, max_blobs=4096, reps=20):
    out = {}
    for _ in range(5):
        out.update(ctx.stage(desc, frames, want=want, max_blobs=max_blobs, out=out))
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    for _ in range(reps):
        out.update(ctx.stage(desc, frames, want=want, max_blobs=max_blobs, out=out))
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    ctx.sync()
    n = frames.shape[0]
    us = e0.elapsed_time(e1) / reps * 1e3 / n
    ctx.profile(True)
    for _ in range(4):
        out.update(ctx.stage(desc, frames, want=want, max_blobs=max_blobs, out=out))
    prof = ctx.profile_dump()
    ctx.profile(False)
    tot = sum(v["ms"] for v in prof.values())
    print("%s: %.2f us/frame in a real step (%.0f frames/s); per-kernel us/frame (events, sum %.2f):" % (name, us, 1e6 / us, tot * 1e3 / 4 / n))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("    %-34s %7.2f us/frame  %3d launches/step  %7.2f us/launch" % (k, v["ms"] * 1e3 / 4 / n, v["launches"] // 4, v["ms"] * 1e3 / v["launches"]))
    sys.stdout.flush()


def main():
    ctx = bv.Context(0)
    f1080 = ctx.upload(np.stack([synth.gen_underwater(1080, 1920, 10 + i) for i in range(16)]))
    run(ctx, "C3 16 x 1920x1080 HSV inRange -> OPEN 5x5 -> labels + moments",
        ctx.make_stage(cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True), f1080,
        ("mask", "labels", "blobs"))
    del f1080
    f4k = ctx.upload(np.stack([synth.gen_c5_frame(100 + i) for i in range(8)]))
    run(ctx, "C5 8 x 3840x2160 balance -> HSV inRange -> OPEN 5x5 -> labels + moments",
        ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True), f4k,
        ("mask", "labels", "blobs"))
    del f4k
    fz = ctx.upload(np.stack([synth.gen_underwater(1242, 2208, 10 + i) for i in range(16)]))
    run(ctx, "fused 16 x 2208x1242 balance -> HSV inRange -> OPEN 5x5 -> mask",
        ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), fz, ("mask",))
    del fz
    f480 = ctx.upload(np.stack([synth.gen_underwater(480, 640, 10 + i) for i in range(64)]))
    run(ctx, "C1 64 x 640x480 LAB a-channel inRange -> OPEN -> CLOSE",
        ctx.make_stage(cvt="bgr2lab", lo=(0, 150, 0), hi=(255, 255, 255), morph=[("open", 5, 5, 1), ("close", 5, 5, 1)]), f480, ("mask",))
    ctx.close()


if __name__ == "__main__":
    main()
