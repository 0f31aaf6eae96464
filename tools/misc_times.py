#!/usr/bin/env python3
"""CUDA-event times of the secondary kernels on 2208x1242 frames (blur, warp, LAB2BGR, LUV, HSI balance, contours + rectangles)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from cuauv_vision_pipeline_b200 import transform  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

ctx = bv.Context(0)
H, W, N = 1242, 2208, 8
frames = ctx.upload(np.stack([synth.gen_underwater(H, W, 5000 + i) for i in range(N)]))
mask = ctx.upload(np.stack([synth.mask_blobs(H, W, i, sigma=5.0, pct=72) for i in range(2)]))


def timed(name, fn, reps=10, per=N):
    for _ in range(2):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            fn()
        e1.record(ctx.torch_stream)
    ctx.sync()
    print("%-46s %8.1f us/frame" % (name, e0.elapsed_time(e1) * 1e3 / reps / per), flush=True)


timed("GaussianBlur 3x3", lambda: ctx.gaussian_blur(frames, (3, 3)))
timed("GaussianBlur 5x5", lambda: ctx.gaussian_blur(frames, (5, 5)))
timed("GaussianBlur 31x31", lambda: ctx.gaussian_blur(frames, (31, 31)))
m = transform.rotation_matrix_2d((W / 2, H / 2), 10, 1)
timed("warpAffine rotate 10 deg (replicate)", lambda: ctx.warp_affine(frames, m, border="replicate"))
_K = np.array([[904.66192735, 0.0, 481.17596262], [0.0, 902.84000422, 404.82437525], [0.0, 0.0, 1.0]]) * (W / 964.0)
_K[2, 2] = 1.0
_maps_f = transform.init_undistort_rectify_map(_K, [0.1, -0.2, 0.001, 0.002, 0.05], None, None, (W, H))
_maps_q = transform.init_undistort_rectify_map(_K, [0.1, -0.2, 0.001, 0.002, 0.05], None, None, (W, H), fixed=True)
timed("undistortion maps (float32 pair, once per camera)", lambda: transform.init_undistort_rectify_map(
    _K, [0.1, -0.2, 0.001, 0.002, 0.05], None, None, (W, H)), per=1)
timed("remap undistort, float32 maps", lambda: ctx.remap(frames, *_maps_f))
timed("remap undistort, fixed-point maps", lambda: ctx.remap(frames, *_maps_q))
timed("BGR2LAB", lambda: ctx.cvt_color(frames, "bgr2lab"))
timed("LAB2BGR", lambda: ctx.cvt_color(frames, "lab2bgr"))
timed("BGR2LUV", lambda: ctx.cvt_color(frames, "bgr2luv"))
timed("balance default flags", lambda: ctx.color_balance(frames))
timed("balance + HSI branch", lambda: ctx.color_balance(frames, hsi_contrast_correct=True), reps=3)


def contours():
    t, nb, pts, npts = ctx.outer_contours(mask, max_contours=4096, max_points=200000)
    return t, nb, pts


timed("outer contours + vertex lists (2 masks)", contours, per=2)
t, nb, pts = contours()
timed("minAreaRect of all contours (2 masks)", lambda: bv.runtime.check(bv.runtime.lib.bv_min_area_rects(
    ctx.handle, bv.runtime.ffi.cast("bv_contour *", t.data_ptr()), bv.runtime.ffi.cast("int32_t *", nb.data_ptr()),
    bv.runtime.ffi.cast("int32_t *", pts.data_ptr()), 2, 4096, 200000,
    bv.runtime.ffi.cast("bv_rrect *", ctx.empty((2, 4096, 24), torch.uint8).data_ptr()))), per=2)
print("contours per mask:", ctx.download(nb).tolist())

# per-kernel split of the contour call (library profiler, serialised launches)
ctx.profile(True)
for _ in range(5):
    contours()
prof = ctx.profile_dump()
ctx.profile(False)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print("   %-28s %3d launches/call  %8.1f us/call" % (k, v["launches"] // 5, v["ms"] * 1e3 / 5))
