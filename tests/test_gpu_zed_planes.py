"""GPU parity for the ZED auxiliary-plane conversions against the
reference's own numpy / cv2 expressions."""
import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1242, 2208), (479, 641), (3, 5)])
def test_rgba_to_rgb(ctx, shape):
    from cuauv_vision_pipeline_b200 import zed_planes
    x = np.random.default_rng(1).integers(0, 256, shape + (4,), dtype=np.uint8)
    assert np.array_equal(zed_planes.to_rgb(x), cv2.cvtColor(x, cv2.COLOR_RGBA2RGB))     # capture_sources/zed.py:49-50


def test_normals_to_rgb01(ctx):
    from cuauv_vision_pipeline_b200 import zed_planes
    n = np.random.default_rng(2).uniform(-1, 1, (240, 320, 4)).astype(np.float32)
    want = ((n[..., :3] + np.float32(1.0)) * np.float32(0.5)).astype(np.float32)          # zed.cpp:86-88
    assert np.array_equal(zed_planes.normals_to_rgb01(n), want)


@pytest.mark.parametrize("shape", [(720, 1280), (479, 641)])
def test_depth_and_normal_display_casts(ctx, shape):
    from cuauv_vision_pipeline_b200 import zed_planes
    rng = np.random.default_rng(3)
    depth = rng.uniform(-1.0, 25.0, shape).astype(np.float32)
    lo, hi = 0.3, 20.0
    # modules/poster.py:41-44
    want = (depth - lo) / (hi - lo)
    want = np.clip(want * 255, 0, 255).astype(np.uint8)
    assert np.array_equal(zed_planes.depth_to_u8(depth, lo, hi), want)
    # modules/record.py:106-109
    want2 = (depth - lo) / (hi - lo)
    want2 = (np.clip(want2, 0.0, 1.0) * 255).astype(np.uint8)
    assert np.array_equal(zed_planes.depth_to_u8(depth, lo, hi, clip_before_scale=True), want2)
    normal = rng.uniform(-0.2, 1.2, shape + (3,)).astype(np.float32)
    assert np.array_equal(zed_planes.normal_to_u8(normal), np.clip(normal * 255, 0, 255).astype(np.uint8))  # poster.py:47
    # invalid depth is defined here (NaN -> 0, +inf -> 255) instead of numpy's undefined cast
    bad = np.array([[np.nan, np.inf, -np.inf, 1.0]], np.float32)
    assert zed_planes.depth_to_u8(bad, 0.0, 2.0).tolist() == [[0, 255, 0, 127]]


def test_channel_means(ctx):
    from cuauv_vision_pipeline_b200 import zed_planes
    from oracle import synth
    img = synth.gen_underwater(1080, 1920, 5)
    got = zed_planes.channel_means(img)
    want = img.reshape(-1, 3).sum(axis=0, dtype=np.int64) / (1080 * 1920)                # exact; np.mean agrees to ~1e-13
    assert np.array_equal(got, want)
    assert np.allclose(got, np.mean(img, axis=(0, 1)), rtol=0, atol=1e-9)                # auto_calibrate_zed.py:82
