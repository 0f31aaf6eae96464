#!/usr/bin/env python3
"""In-process sweep of the colour-balance tuning knobs (run under gpurun): device-resident frames/s
of C2 (balance -> LAB image) and of the fused mask stage (balance -> HSV -> inRange -> OPEN).
    python tools/tune.py > gpurun_out/tune.log"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

H, W, BATCH, RING = 1242, 2208, 16, 48
ctx = bv.Context(0)
base = np.stack([synth.gen_underwater(H, W, 2000 + i) for i in range(8)])
ring = ctx.upload(np.concatenate([np.roll(base, 7 * k, axis=2) for k in range(RING // 8)]))
c2 = ctx.make_stage(balance={}, cvt="bgr2lab")
fused = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)])


def run(desc, want, steps=24, warmup=3):
    out = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(ctx.torch_stream):
        for i in range(warmup + steps):
            if i == warmup:
                ev[0].record(ctx.torch_stream)
            o = (i * BATCH) % RING
            out.update(ctx.stage(desc, ring[o:o + BATCH], want=want, out=out))
        ev[1].record(ctx.torch_stream)
    ctx.sync()
    return steps * BATCH / (ev[0].elapsed_time(ev[1]) * 1e-3)


grid = list(itertools.product([1, 2, 4], [8, 17, 33, 66], [0, 2, 4, 8], [0, 2, 4, 8]))
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    grid = [(4, 33, 0, 0), (1, 66, 0, 0), (2, 17, 0, 0), (4, 17, 0, 0)]
print("side l2_mb hist_bps final_bps -> C2 fps | fused fps", flush=True)
for side, l2, hb, fb in grid:
    for k, v in (("side_streams", side), ("l2_chunk_mb", l2), ("hist_bps", hb), ("final_bps", fb)):
        ctx.set_option(k, v)
    print("%d %2d %d %d -> %6.0f | %6.0f" % (side, l2, hb, fb, run(c2, ("converted",)), run(fused, ("mask",))), flush=True)
