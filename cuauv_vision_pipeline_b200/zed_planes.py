"""GPU versions of the ZED auxiliary-plane conversions around the hot path:
capture_sources/zed.py:49-50 / zed.cpp:54-91 (RGBA -> RGB, normals -> [0,1]) and the display casts
of modules/poster.py:41-47, modules/record.py:106-113, modules/calibrate.py:111."""
import numpy as np
import torch

from ._host import ctx_for as _ctx, like_input, to_device


def _dev(ctx, x, dtype):
    return to_device(ctx, x, dtype)       # ordered behind the caller's stream when x is a CUDA tensor


def _back(ctx, x, t):
    return like_input(ctx, x, t)


def to_rgb(x):
    """capture_sources/zed.py:49-50 == cv2.cvtColor(x, cv2.COLOR_RGBA2RGB)."""
    ctx = _ctx(x)
    return _back(ctx, x, ctx.rgba_to_rgb(_dev(ctx, x, np.uint8)))


def normals_to_rgb01(normals_xyzw):
    """capture_sources/zed.cpp:73-91: float32 [H,W,4] -> float32 [H,W,3], (v + 1) * 0.5."""
    ctx = _ctx(normals_xyzw)
    return _back(ctx, normals_xyzw, ctx.normals_to_rgb01(_dev(ctx, normals_xyzw, np.float32)))


def depth_to_u8(depth, min_distance, max_distance, clip_before_scale=False):
    """modules/poster.py:41-44 (clip_before_scale=False) / modules/record.py:106-109 (True)."""
    ctx = _ctx(depth)
    return _back(ctx, depth, ctx.f32_to_u8(_dev(ctx, depth, np.float32), sub=np.float32(min_distance),
                                           div=np.float32(max_distance - min_distance),
                                           clip_before_scale=clip_before_scale))


def normal_to_u8(normal):
    """modules/poster.py:47, modules/record.py:113, modules/calibrate.py:111: clip(normal * 255, 0, 255)."""
    ctx = _ctx(normal)
    return _back(ctx, normal, ctx.f32_to_u8(_dev(ctx, normal, np.float32)))


def channel_means(img):
    """modules/auto_calibrate_zed.py:82: np.mean(img, axis=(0, 1))."""
    ctx = _ctx(img)
    return ctx.channel_means(_dev(ctx, img, np.uint8))
