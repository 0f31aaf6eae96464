"""Pins the arithmetic specification (oracle/spec_np.py) against the installed cv2: every 8-bit
conversion over all 2^24 colours, resize over mixed shapes, morphology / inRange semantics."""
import hashlib

import cv2
import numpy as np
import pytest

from oracle import spec_np as S
from oracle import synth, ccl, letterbox, cv_ops


@pytest.fixture(scope="module")
def all_colors():
    return synth.all_colors_image()


@pytest.mark.parametrize("name,fn,code", [("hsv", S.bgr2hsv, cv2.COLOR_BGR2HSV), ("lab", S.bgr2lab, cv2.COLOR_BGR2LAB),
                                          ("gray", S.bgr2gray, cv2.COLOR_BGR2GRAY),
                                          ("ycrcb", S.bgr2ycrcb, cv2.COLOR_BGR2YCrCb),
                                          ("hls", S.bgr2hls, cv2.COLOR_BGR2HLS)])
def test_conversion_all_colors(all_colors, name, fn, code):
    ref = cv2.cvtColor(all_colors, code)
    for y0 in range(0, 4096, 512):
        assert np.array_equal(fn(all_colors[y0:y0 + 512]), ref[y0:y0 + 512]), name


def test_lab_tables_checksums():
    g, c = S.lab_tables()
    assert hashlib.sha256(g.astype("<u2").tobytes()).hexdigest()[:16] == "8bfeace00785402e"
    assert hashlib.sha256(c.astype("<u2").tobytes()).hexdigest()[:16] == "8d9ac99d93fa1ed9"


def _all_hsv():
    hh, ss, vv = np.meshgrid(np.arange(180), np.arange(256), np.arange(256), indexing="ij")
    return np.stack([hh, ss, vv], -1).astype(np.uint8).reshape(-1, 3)


@pytest.mark.parametrize("width", [4096, 63, 33])
def test_hsv2bgr_vector_and_tail_rule(width):
    flat = _all_hsv()
    n = (flat.shape[0] // width) * width
    im = np.ascontiguousarray(flat[:n].reshape(-1, width, 3))
    step = max(1, im.shape[0] // 400)       # a stride of rows keeps the run short; every (S,V) pair still appears
    im = np.ascontiguousarray(im[::step])
    assert np.array_equal(S.hsv2bgr_rows(im), cv2.cvtColor(im, cv2.COLOR_HSV2BGR))


@pytest.mark.parametrize("shape", [(1242, 2208, 360, 640, 3), (479, 641, 777, 333, 3), (480, 640, 960, 1280, 1),
                                   (100, 100, 37, 53, 3), (720, 1280, 360, 640, 3), (1080, 1920, 459, 816, 3)])
def test_resize_linear(shape):
    sh, sw, dh, dw, c = shape
    im = np.random.default_rng(sh + dw).integers(0, 256, (sh, sw, c), dtype=np.uint8)
    if c == 1:
        im = im[..., 0]
    assert np.array_equal(S.resize_linear(im, dw, dh), cv2.resize(im, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_morphology_border_semantics():
    m = synth.mask_random(40, 50, 3, 0.6)
    k = np.ones((5, 5), np.uint8)
    pad = cv2.copyMakeBorder(m, 2, 2, 2, 2, cv2.BORDER_CONSTANT, value=255)
    er = np.min(np.stack([pad[dy:dy + 40, dx:dx + 50] for dy in range(5) for dx in range(5)]), 0)
    assert np.array_equal(cv2.erode(m, k), er)
    pad0 = cv2.copyMakeBorder(m, 2, 2, 2, 2, cv2.BORDER_CONSTANT, value=0)
    di = np.max(np.stack([pad0[dy:dy + 40, dx:dx + 50] for dy in range(5) for dx in range(5)]), 0)
    assert np.array_equal(cv2.dilate(m, k), di)
    assert np.array_equal(cv2.morphologyEx(m, cv2.MORPH_OPEN, k, iterations=2),
                          cv2.dilate(cv2.erode(m, k, iterations=2), k, iterations=2))


def test_canonical_labels_and_moments():
    m = synth.mask_blobs(120, 200, 2, sigma=4.0)
    n, lab, tab = ccl.label_and_moments(m)
    assert n > 3
    firsts = [int(np.flatnonzero(lab.reshape(-1) == i)[0]) for i in range(1, n + 1)]
    assert firsts == sorted(firsts)
    for i in (1, n // 2, n):
        mm = ccl.cv2_moments_of_label(lab, i)
        for k in ccl.MOMENT_KEYS:
            assert int(mm[k]) == int(tab[k][i - 1])
    n2, _, stats, _ = cv2.connectedComponentsWithStats((m != 0).astype(np.uint8), connectivity=8)
    assert n2 - 1 == n and sorted(stats[1:, cv2.CC_STAT_AREA].tolist()) == sorted(tab["m00"].tolist())


def test_letterbox_geometry():
    assert letterbox.letterbox_geometry(1242, 2208) == (640, 360, 140, 140, 0, 0)
    assert letterbox.letterbox_geometry(720, 1280) == (640, 360, 140, 140, 0, 0)
    assert letterbox.letterbox_geometry(480, 640) == (640, 480, 80, 80, 0, 0)
    out = letterbox.yolo_input([synth.gen_underwater(90, 160, 1)], 64, 64)
    assert out.shape == (1, 3, 64, 64) and out.dtype == np.float16
    assert float(out[0, 0, 0, 0]) == pytest.approx(114 / 255, abs=1e-3)


def test_preprocessor_point_ops_semantics():
    img = synth.gen_random_bgr(16, 16, 0)
    assert np.array_equal(cv_ops.contrast(img, 1.7), np.minimum(np.floor(img * 1.7), 255).astype(np.uint8))
    assert np.array_equal(cv_ops.brightness(img, -40), np.clip(img.astype(int) - 40, 0, 255).astype(np.uint8))
    assert np.array_equal(cv_ops.channel_bias(img, 2, 25)[..., 2], np.clip(img[..., 2].astype(int) + 25, 0, 255))


def test_color_distance_restatement_matches_reference_function_golden():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv_calls_72x128.npz"))
    split = list(cv2.split(z["bgr2lab"]))
    cases = [dict(color=(120, 150, 140), distance=30), dict(color=(60, 128, 128), distance=45, weights=(0.2, 1, 1)),
             dict(color=(200, 110, 170), distance=25.5, ignore_channels=[0]),
             dict(color=(10, 240, 20), distance=400, weights=(3, 1, 2))]
    for k, kw in enumerate(cases):
        m, d = cv_ops.thresh_color_distance(split, **kw)
        assert np.array_equal(m, z["tcd%d_mask" % k]) and np.array_equal(d, z["tcd%d_dist" % k])


def test_hue_interval_property_of_the_hsv_round_trip():
    """The property the CUDA fast path for `balance -> BGR2HSV -> inRange` rests on (balance.cu, "hue-interval
    table"), pinned on cv2 itself: for every (S, V) the hues H for which
    inRange(BGR2HSV(HSV2BGR(H, S, V))) passes form ONE cyclic interval of [0, 180), and S2, V2 of the
    round trip do not depend on H."""
    cube = np.empty((180, 65536, 3), np.uint8)
    cube[..., 0] = np.arange(180, dtype=np.uint8)[:, None]
    sv = np.arange(65536)
    cube[..., 1] = (sv >> 8).astype(np.uint8)[None, :]
    cube[..., 2] = (sv & 255).astype(np.uint8)[None, :]
    back = cv2.cvtColor(cv2.cvtColor(cube, cv2.COLOR_HSV2BGR), cv2.COLOR_BGR2HSV)     # width 65536: vector path throughout
    assert (back[..., 1] == back[0:1, :, 1]).all() and (back[..., 2] == back[0:1, :, 2]).all()
    for lo, hi in [((10, 20, 60), (30, 100, 255)), ((0, 0, 0), (5, 255, 255)), ((170, 0, 0), (179, 255, 255)),
                   ((0, 0, 0), (179, 30, 40)), ((37, 5, 9), (121, 250, 251)), ((90, 0, 0), (90, 255, 255))]:
        m = cv2.inRange(back, np.array(lo), np.array(hi)) > 0                          # [180, 65536]
        rises = (m & ~np.roll(m, 1, axis=0)).sum(axis=0)
        assert int(rises.max()) <= 1, (lo, hi)


@pytest.mark.parametrize("n", [1, 3, 5, 7, 9, 11, 13, 17, 21, 31, 63, 101, 201])
def test_gaussian_blur_8u_model_vs_cv2(n):
    """modules/preprocessor.py:110-114: cv2.GaussianBlur(mat, (n, n), 0) on uint8, 8.8 fixed point."""
    rng = np.random.default_rng(n)
    assert np.allclose(S.gaussian_kernel_f64(n), cv2.getGaussianKernel(n, 0)[:, 0], rtol=0, atol=1e-15)
    for shape in [(37, 53, 3), (120, 160, 3), (64, 48), (5, 4, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(S.gaussian_blur_8u(img, (n, n)), cv2.GaussianBlur(img, (n, n), 0)), shape
    img = synth.gen_underwater(96, 128, n)
    assert np.array_equal(S.gaussian_blur_8u(img, (n, 3), 1.7, 0.9), cv2.GaussianBlur(img, (n, 3), 1.7, sigmaY=0.9))


def test_warp_affine_8u_model_vs_cv2():
    """modules/preprocessor.py:130-135 (rotate, BORDER_REPLICATE) and 144-149 (translate, constant 0)."""
    rng = np.random.default_rng(3)
    for shape in [(120, 160, 3), (97, 131, 3), (64, 80)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        h, w = shape[:2]
        for ang in (5, 30, -47.5, 90, 180, 0.3, 359):
            m = cv2.getRotationMatrix2D((w / 2, h / 2), ang, 1)
            assert np.array_equal(S.rotation_matrix_2d((w / 2, h / 2), ang, 1), m)
            assert np.array_equal(S.warp_affine_8u(img, m, (w, h), "replicate"),
                                  cv2.warpAffine(img, m, (w, h), borderMode=cv2.BORDER_REPLICATE)), (shape, ang)
        for tx, ty in ((5, 0), (0, 7), (-13, 4), (3.5, -2.25), (200, 0)):
            m = np.float32([[1, 0, tx], [0, 1, ty]])
            assert np.array_equal(S.warp_affine_8u(img, m, (w, h)), cv2.warpAffine(img, m, (w, h))), (shape, tx, ty)
    img = synth.gen_underwater(90, 120, 4)
    m = np.array([[0.8, 0.3, 10.5], [-0.2, 1.1, -4.25]])
    assert np.array_equal(S.warp_affine_8u(img, m, (150, 70), "constant", (7, 99, 200)),
                          cv2.warpAffine(img, m, (150, 70), borderValue=(7, 99, 200)))


@pytest.mark.parametrize("seed,pre,n", [(1, 0, 1000), (2, 1, 1001), (3, 3, 7), (4, 0, 1), (5, 1, 1), (6, 311, 5883),
                                        (7, 2, 60000)])
def test_legacy_randn_replay_matches_numpy(seed, pre, n):
    """The restated MT19937 + polar method (oracle/mt_gauss_np.py) gives numpy.random.randn's values and leaves
    the generator state numpy leaves (key, position, cached second value), from aligned and unaligned positions,
    with and without a cached value at the start."""
    from oracle import mt_gauss_np
    np.random.seed(seed)
    if pre:
        np.random.randn(pre)
    state = np.random.get_state()
    ref = np.random.randn(n)
    after_ref = np.random.get_state()
    got, after = mt_gauss_np.randn_replay(state, n)
    assert np.array_equal(got, ref)
    assert np.array_equal(after[1], after_ref[1]) and tuple(after[2:]) == tuple(after_ref[2:])


_CAM_K = np.array([[904.66192735, 0.0, 481.17596262], [0.0, 902.84000422, 404.82437525], [0.0, 0.0, 1.0]])
_CAM_D = np.array([0.48525658, 2.02550297, 0.03807578, -0.02152142, -3.30299241])   # lib/configs/1_camera_matrix_params.yaml


def test_remap_linear_8u_model_vs_cv2():
    """cv2.remap(INTER_LINEAR) on uint8 ("bilinear remap fixed-point model, not yet pinned"): float32 maps
    are rounded to 1/32 pixel as RemapInvoker does, then the same 32 x 32 x 4 int16 weights as warpAffine."""
    img = synth.gen_underwater(240, 320, 3)
    rng = np.random.default_rng(1)
    mx = (np.arange(320, dtype=np.float32)[None, :] + rng.normal(0, 4, (240, 320))).astype(np.float32)
    my = (np.arange(240, dtype=np.float32)[:, None] + rng.normal(0, 4, (240, 320))).astype(np.float32)
    for border, mode in (("constant", cv2.BORDER_CONSTANT), ("replicate", cv2.BORDER_REPLICATE)):
        assert np.array_equal(S.remap_linear_8u(img, mx, my, border), cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=mode))
        assert np.array_equal(S.remap_linear_8u(img[..., 1], mx, my, border),
                              cv2.remap(np.ascontiguousarray(img[..., 1]), mx, my, cv2.INTER_LINEAR, borderMode=mode))
    m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    xy, frac = S.fixed_point_maps(mx, my)
    assert np.array_equal(xy, m1) and np.array_equal(frac, m2)
    assert np.array_equal(S.remap_linear_8u(img, xy, frac), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    assert np.array_equal(S.remap_linear_8u(img, mx, my, "constant", (7, 8, 9)),
                          cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(7, 8, 9)))


def test_undistort_maps_model_vs_cv2():
    """cv2.initUndistortRectifyMap / cv2.undistort with the reference's camera file: identical maps (float32 and the fixed-point
    pair), identical undistorted image; also with a rotation and 8 distortion coefficients."""
    for size in ((964, 724), (640, 480)):
        new_k, _ = cv2.getOptimalNewCameraMatrix(_CAM_K, _CAM_D, size, 1)
        for nk in (_CAM_K, new_k):
            rx, ry = cv2.initUndistortRectifyMap(_CAM_K, _CAM_D, None, nk, size, cv2.CV_32FC1)
            gx, gy = S.init_undistort_rectify_map(_CAM_K, _CAM_D, nk, size)
            assert np.array_equal(rx, gx) and np.array_equal(ry, gy)
            r1, r2 = cv2.initUndistortRectifyMap(_CAM_K, _CAM_D, None, nk, size, cv2.CV_16SC2)
            g1, g2 = S.init_undistort_rectify_map(_CAM_K, _CAM_D, nk, size, fixed=True)
            assert np.array_equal(r1, g1) and np.array_equal(r2, g2)
            img = synth.gen_underwater(size[1], size[0], 5)
            assert np.array_equal(S.remap_linear_8u(img, g1, g2), cv2.undistort(img, _CAM_K, _CAM_D, None, nk))
    rot, _ = cv2.Rodrigues(np.array([0.02, -0.03, 0.01]))
    d8 = np.array([0.1, -0.2, 0.001, 0.002, 0.05, 0.01, -0.02, 0.003])
    rx, ry = cv2.initUndistortRectifyMap(_CAM_K, d8, rot, _CAM_K, (320, 240), cv2.CV_32FC1)
    gx, gy = S.init_undistort_rectify_map(_CAM_K, d8, _CAM_K, (320, 240), rotation=rot)
    assert np.array_equal(rx, gx) and np.array_equal(ry, gy)
