#!/usr/bin/env python3
"""C5 per-GPU batch at N = 8 (one stream, 4 consecutive 3840x2160 frames per call): chunking of the balance passes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

ctx = bv.Context(0)
base = synth.gen_c5_frame(3200)
for nb in (1, 2, 4, 8):
    frames = ctx.upload(np.stack([np.roll(base, 37 * i, axis=1) for i in range(nb)]))
    desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
    for mb in (0, 25, 50, 100):
        ctx.set_option("l2_chunk_mb", mb)
        out = {}
        for _ in range(3):
            out.update(ctx.stage(desc, frames, want=("mask", "labels", "blobs"), max_blobs=8192, out=out))
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.torch_stream):
            e0.record()
        for _ in range(10):
            out.update(ctx.stage(desc, frames, want=("mask", "labels", "blobs"), max_blobs=8192, out=out))
        with torch.cuda.stream(ctx.torch_stream):
            e1.record()
        ctx.sync()
        us = e0.elapsed_time(e1) * 1e3 / 10 / nb
        print("batch %d x 4K, l2_chunk_mb=%3d: %.1f us/frame (%.0f frames/s)" % (nb, mb, us, 1e6 / us), flush=True)
