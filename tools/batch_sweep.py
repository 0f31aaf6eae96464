#!/usr/bin/env python3
"""C2 / fused stage device-resident throughput against frames per call (the join at the end of a call drains all side
streams; more frames per call amortise it).  python tools/batch_sweep.py > gpurun_out/r02_batch_sweep.log"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

H, W = 1242, 2208
ctx = bv.Context(0)
base = np.stack([synth.gen_underwater(H, W, 10 + i) for i in range(8)])
ring = ctx.upload(np.stack([np.roll(base[i % 8], 31 * i, axis=1) for i in range(128)]))   # 1.05 GB
descs = {"C2": (ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",)),
         "fused": (ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",))}
for name, (desc, want) in descs.items():
    for batch in (4, 8, 16, 32, 64, 128):
        nb = 128 // batch
        outs = {}
        views = [ring[k * batch:(k + 1) * batch] for k in range(nb)]
        reps = max(2, 512 // batch)
        for s in range(3):
            outs.update(ctx.stage(desc, views[s % nb], want=want, out=outs))
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.torch_stream):
            e0.record()
        for s in range(reps):
            outs.update(ctx.stage(desc, views[s % nb], want=want, out=outs))
        with torch.cuda.stream(ctx.torch_stream):
            e1.record()
        ctx.sync()
        us = e0.elapsed_time(e1) * 1e3 / (reps * batch)
        print("%-6s %3d frames per call: %6.2f us/frame = %7.0f frames/s" % (name, batch, us, 1e6 / us), flush=True)
        del outs
ctx.close()
