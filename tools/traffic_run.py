#!/usr/bin/env python3
"""One full 16-frame step (side streams ON, 64-frame ring as in bench.py) between cudaProfilerStart / Stop, for a
steady-state DRAM-traffic capture:
    ncu --replay-mode range --cache-control none --clock-control none --profile-from-start off \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum python tools/traffic_run.py c2
(range replay keeps the kernels of the step concurrent and does not flush the caches between them, which kernel replay
does; numbers printed here are never bench values).  python tools/traffic_run.py [c2|fused] [warm-up steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic input generator only)

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = bv.Context(0)
base = [synth.gen_underwater(1242, 2208, 2000 + i) for i in range(8)]
ring = ctx.upload(np.stack([np.roll(base[i % 8], 31 * (i // 8), axis=1) for i in range(64)]))     # 526 MB > L2
if what == "c2":
    desc, want = ctx.make_stage(balance={}, cvt="bgr2lab"), ("converted",)
else:
    desc, want = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)]), ("mask",)
out = {}
for s in range(warm):
    b = s % 4
    out.update(ctx.stage(desc, ring[b * 16:(b + 1) * 16], want=want, out=out))
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
b = warm % 4
out.update(ctx.stage(desc, ring[b * 16:(b + 1) * 16], want=want, out=out))
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", what, "frames in the profiled range: 16, launches so far", ctx.launches)
