#!/usr/bin/env python3
"""C2 end to end through bv_stage_host_submit / _wait for one BV_HOST_CHUNK_MB (read once per process):
    for mb in 8 16 25 33 66 132; do BV_HOST_CHUNK_MB=$mb python tools/e2e_pipe_sweep.py; done"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

H, W, B = 1242, 2208, 16
ctx = bv.Context(0)
pin_in = bv.PinnedArray((2, B, H, W, 3))
base = np.stack([synth.gen_underwater(H, W, 10 + i) for i in range(4)])
for k in range(2):
    for i in range(B):
        pin_in.array[k, i] = np.roll(base[i % 4], 17 * (i + k), axis=1)
outs = [{"converted": bv.PinnedArray((B, H, W, 3)).array} for _ in range(2)]
desc = ctx.make_stage(balance={}, cvt="bgr2lab")


def run(steps, slots):
    for s in range(steps):
        ctx.stage_host(desc, pin_in.array[s % 2], want=("converted",), out=outs[s % 2], slot=(s % 2 if slots == 2 else 0))
    ctx.stage_host_wait(0)
    ctx.stage_host_wait(1)


res = []
for slots in (1, 2):
    run(4, slots)
    t0 = time.perf_counter()
    run(40, slots)
    dt = time.perf_counter() - t0
    res.append(40 * B / dt)
print("BV_HOST_CHUNK_MB=%s: one batch in flight %7.0f frames/s, two in flight %7.0f frames/s"
      % (os.environ.get("BV_HOST_CHUNK_MB", "default(16)"), res[0], res[1]), flush=True)
ctx.close()
