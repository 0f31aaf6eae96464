#!/usr/bin/env python3
"""bins-module end to end (frames in, blob tables out) against the host chunk size of bv_stage_host.
    for mb in 16 33 66 132; do BV_HOST_CHUNK_MB=$mb python tools/bins_e2e_sweep.py; done"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

H, W, B = 1242, 2208, 16
ctx = bv.Context(0)
pin = bv.PinnedArray((B, H, W, 3))
pin.array[...] = np.stack([synth.gen_underwater(H, W, 2000 + i) for i in range(4)] * 4)
desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
out = {}
for _ in range(3):
    out = ctx.stage_host(desc, pin.array, want=("blobs",), max_blobs=1024, out=out)
t0 = time.perf_counter()
for _ in range(10):
    out = ctx.stage_host(desc, pin.array, want=("blobs",), max_blobs=1024, out=out)
dt = (time.perf_counter() - t0) / 10
print("BV_HOST_CHUNK_MB=%s: bins module end to end %.0f frames/s (%.3f ms per 16 frames)" % (os.environ.get("BV_HOST_CHUNK_MB", "16"), B / dt, dt * 1e3))
