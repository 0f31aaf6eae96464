#!/usr/bin/env python3
"""C2 / fused stage: side streams (chunks in flight) x chunk size, 32 frames per call.
    python tools/side_sweep.py > gpurun_out/r02_side_sweep.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

H, W, B = 1242, 2208, 32
ctx = bv.Context(0)
base = np.stack([synth.gen_underwater(H, W, 10 + i) for i in range(8)])
ring = ctx.upload(np.stack([np.roll(base[i % 8], 31 * i, axis=1) for i in range(64)]))
descs = {"C2": (ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",)),
         "fused": (ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",))}
for name, (desc, want) in descs.items():
    for side in (4, 6, 8):
        for l2 in (17, 25, 33, 50):
            ctx.set_option("side_streams", side)
            ctx.set_option("l2_chunk_mb", l2)
            outs = {}
            views = [ring[:B], ring[B:]]
            for s in range(3):
                outs.update(ctx.stage(desc, views[s % 2], want=want, out=outs))
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ctx.torch_stream):
                e0.record()
            reps = 20
            for s in range(reps):
                outs.update(ctx.stage(desc, views[s % 2], want=want, out=outs))
            with torch.cuda.stream(ctx.torch_stream):
                e1.record()
            ctx.sync()
            us = e0.elapsed_time(e1) * 1e3 / (reps * B)
            print("%-6s side %d chunk %2d MB (%d frames): %6.2f us/frame = %7.0f frames/s" % (name, side, l2, (l2 << 20) // (H * W * 3), us, 1e6 / us), flush=True)
ctx.close()
