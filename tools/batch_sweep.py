#!/usr/bin/env python3
"""Is the C2 step launch-bound?  Device-resident frames/s over frames per call."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

H, W, RING = 1242, 2208, 64
ctx = bv.Context(0)
base = np.stack([synth.gen_underwater(H, W, 2000 + i) for i in range(8)])
ring = ctx.upload(np.concatenate([np.roll(base, 7 * k, axis=2) for k in range(RING // 8)]))
for name, desc, want in (("C2", ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",)),
                         ("fused", ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",))):
    for batch, chunk_mb in [(b, c) for b in (1, 2, 4, 8, 16, 32, 64) for c in ((0, 9, 17) if 1 < b <= 16 else (0,))]:
        ctx.set_option("l2_chunk_mb", chunk_mb)
        out = {}
        steps = max(4, 512 // batch)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.stream(ctx.torch_stream):
            for i in range(3 + steps):
                if i == 3:
                    ev[0].record(ctx.torch_stream)
                o = (i * batch) % (RING - batch + 1)
                out.update(ctx.stage(desc, ring[o:o + batch], want=want, out=out))
            ev[1].record(ctx.torch_stream)
        ctx.sync()
        print("%-5s frames/call %2d chunk_mb %2d -> %6.0f frames/s" % (name, batch, chunk_mb, steps * batch / (ev[0].elapsed_time(ev[1]) * 1e-3)), flush=True)
