#!/usr/bin/env python3
"""Where the end-to-end leg (bv_stage_host, C2, 16 frames 2208x1242 from pinned memory) spends its time:
 * plain copies of the same pinned buffers in 8.2 MB pieces, both directions at once, with the dependency structure of the
   pipeline (download k after upload k) but NO kernels: the ceiling for this chunking;
 * the library's own device time stamps per chunk (BV_HOST_TIMELINE=1), for a few chunk sizes.
    python tools/e2e_timeline.py 2> gpurun_out/r02_e2e_timeline.log"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BV_HOST_TIMELINE"] = "1" if "--timeline" in sys.argv else os.environ.get("BV_HOST_TIMELINE", "")
if not os.environ["BV_HOST_TIMELINE"]:
    del os.environ["BV_HOST_TIMELINE"]
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

H, W, B = 1242, 2208, 16
ctx = bv.Context(0)
frames = np.stack([synth.gen_underwater(H, W, 2000 + i) for i in range(4)] * 4)
pin_in = bv.PinnedArray((B, H, W, 3))
pin_in.array[...] = frames
pin_out = bv.PinnedArray((B, H, W, 3))
desc = ctx.make_stage(balance={}, cvt="bgr2lab")
out = {"converted": pin_out.array}


def run(n=8):
    for _ in range(2):
        ctx.stage_host(desc, pin_in.array, want=("converted",), out=out)
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.stage_host(desc, pin_in.array, want=("converted",), out=out)
    return B * n / (time.perf_counter() - t0)


if "--timeline" in sys.argv:
    ctx.stage_host(desc, pin_in.array, want=("converted",), out=out)
    ctx.stage_host(desc, pin_in.array, want=("converted",), out=out)
    sys.exit(0)

# plain copy pipeline, pieces of `pf` frames
t_in, t_out = torch.from_numpy(pin_in.array), torch.from_numpy(pin_out.array)
d_in = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
for pf in (1, 2, 4, 16):
    def pipe():
        evs = []
        for k in range(0, B, pf):
            with torch.cuda.stream(s_up):
                d_in[k:k + pf].copy_(t_in[k:k + pf], non_blocking=True)
                e = torch.cuda.Event()
                e.record()
            with torch.cuda.stream(s_dn):
                s_dn.wait_event(e)
                t_out[k:k + pf].copy_(d_in[k:k + pf], non_blocking=True)
    pipe()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        pipe()
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 8
    print("plain copy pipeline, %2d-frame pieces: %.3f ms per 16 frames -> %.0f frames/s (%.1f GB/s both ways)"
          % (pf, dt * 1e3, B / dt, 2 * B * H * W * 3 / dt / 1e9), file=sys.stderr)
print("bv_stage_host C2 (default chunking): %.0f frames/s" % run(), file=sys.stderr)
