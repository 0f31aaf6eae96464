#!/usr/bin/env python3
"""Small batches: aggregate frames/s of K module threads, each with its own context (stream), 1 or 2 frames per call."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

H, W = 1242, 2208
base = np.stack([synth.gen_underwater(H, W, 2000 + i) for i in range(4)])
for per_call in (1, 2):
    for k in (1, 2, 4, 8):
        ctxs = [bv.Context(0) for _ in range(k)]
        rings = [c.upload(base) for c in ctxs]
        descs = [c.make_stage(balance={}, cvt="bgr2lab") for c in ctxs]
        calls = 300
        barrier = threading.Barrier(k + 1)

        def worker(i):
            c, ring, desc, out = ctxs[i], rings[i], descs[i], {}
            for j in range(5):
                out.update(c.stage(desc, ring[:per_call], want=("converted",), out=out))
            c.sync()
            barrier.wait()
            for j in range(calls):
                o = (j * per_call) % (4 - per_call + 1)
                out.update(c.stage(desc, ring[o:o + per_call], want=("converted",), out=out))
            c.sync()
            barrier.wait()

        ts = [threading.Thread(target=worker, args=(i,)) for i in range(k)]
        for t in ts:
            t.start()
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        dt = time.perf_counter() - t0
        for t in ts:
            t.join()
        print("%d frame(s)/call, %d contexts -> %6.0f frames/s aggregate (%.1f us per call per context)" %
              (per_call, k, k * calls * per_call / dt, dt / calls * 1e6), flush=True)
        for c in ctxs:
            c.close()
