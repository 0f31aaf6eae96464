#!/usr/bin/env python3
"""C3 / C5 / C1: host enqueue time vs device time per step, for 1 / 2 / 4 side streams and several chunk sizes.
    python tools/c3_overlap.py > gpurun_out/r02_c3_overlap.log"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402


def measure(ctx, desc, frames, want, reps=30):
    out = {}
    for _ in range(5):
        out.update(ctx.stage(desc, frames, want=want, max_blobs=4096, out=out))
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    t0 = time.perf_counter()
    for _ in range(reps):
        out.update(ctx.stage(desc, frames, want=want, max_blobs=4096, out=out))
    t_host = time.perf_counter() - t0
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    ctx.sync()
    n0 = ctx.launches
    ctx.stage(desc, frames, want=want, max_blobs=4096, out=out)
    ctx.sync()
    return e0.elapsed_time(e1) / reps * 1e3, t_host / reps * 1e6, ctx.launches - n0


def main():
    ctx = bv.Context(0)
    cases = [
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
        ctx.set_option("side_streams", 0)
        ctx.set_option("l2_chunk_mb", 0)
        dev, hostt, launches = measure(ctx, desc, frames, want)
        print("%-12s library defaults      : device %8.1f us/step (%7.0f frames/s), host enqueue %7.1f us/step, %3d launches"
              % (name, dev, frames.shape[0] / dev * 1e6, hostt, launches), flush=True)
        for side in (1, 2, 4):
            for l2 in (8, 17, 33, 66, 132, 264):
                ctx.set_option("side_streams", side)
                ctx.set_option("l2_chunk_mb", l2)
                dev, hostt, launches = measure(ctx, desc, frames, want)
                print("%-12s side %d  chunk %3d MB: device %8.1f us/step (%7.0f frames/s), host enqueue %7.1f us/step, %3d launches"
                      % (name, side, l2, dev, frames.shape[0] / dev * 1e6, hostt, launches), flush=True)
        del frames
    ctx.close()


if __name__ == "__main__":
    main()
