#!/usr/bin/env python3
"""Knob sweep for C5 (3840x2160 balance -> HSV inRange -> OPEN -> labels + moments), device-resident frames/s."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

ctx = bv.Context(0)
base = np.stack([synth.gen_underwater(2160, 3840, 4000 + i) for i in range(4)])
ring = ctx.upload(np.concatenate([base, np.roll(base, 9, axis=2), np.roll(base, 17, axis=1), np.roll(base, 5, axis=2)]))
desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)


def run(steps=12, warmup=2, batch=8):
    out = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(ctx.torch_stream):
        for i in range(warmup + steps):
            if i == warmup:
                ev[0].record(ctx.torch_stream)
            o = (i * batch) % 16
            out.update(ctx.stage(desc, ring[o:o + batch], want=("mask", "labels", "blobs"), max_blobs=4096, out=out))
        ev[1].record(ctx.torch_stream)
    ctx.sync()
    return steps * batch / (ev[0].elapsed_time(ev[1]) * 1e-3)


print("side l2_mb -> C5 fps", flush=True)
for side, l2 in itertools.product([1, 2, 3, 4], [25, 33, 50, 100]):
    ctx.set_option("side_streams", side)
    ctx.set_option("l2_chunk_mb", l2)
    print("%d %3d -> %6.0f" % (side, l2, run()), flush=True)
