"""GPU parity: colour conversions, inRange, thresholds, LUT -- CUDA kernels through the C ABI
against the reference's own cv2 calls (oracle/cv_ops.py).  Bit-exact unless a tolerance is stated."""
import cv2
import numpy as np
import pytest
import torch

from oracle import cv_ops, synth

pytestmark = pytest.mark.gpu

CODES = [("bgr2hsv", cv2.COLOR_BGR2HSV), ("bgr2lab", cv2.COLOR_BGR2LAB), ("bgr2ycrcb", cv2.COLOR_BGR2YCrCb),
         ("bgr2gray", cv2.COLOR_BGR2GRAY), ("lab2bgr", cv2.COLOR_LAB2BGR)]


@pytest.fixture(scope="module")
def all_colors():
    return synth.all_colors_image()


@pytest.mark.parametrize("name,cvc", CODES)
def test_cvt_all_2_24_colours(ctx, all_colors, name, cvc):
    d = ctx.upload(all_colors)
    got = ctx.download(ctx.cvt_color(d, name))
    assert np.array_equal(got, cv2.cvtColor(all_colors, cvc))


@pytest.mark.parametrize("name,cvc", CODES[:3])
def test_cvt_split_planes(ctx, name, cvc):
    img = synth.gen_underwater(242, 368, 3)
    conv, planes = ctx.cvt_color(ctx.upload(img), name, split=True)
    ref = cv2.cvtColor(img, cvc)
    assert np.array_equal(ctx.download(conv), ref)
    for k, p in enumerate(planes):
        assert np.array_equal(ctx.download(p), ref[..., k])


@pytest.mark.parametrize("width", [4096, 100, 63, 33])
def test_hsv2bgr_every_hsv_and_tail_rule(ctx, width):
    hh, ss, vv = np.meshgrid(np.arange(180), np.arange(256), np.arange(256), indexing="ij")
    flat = np.stack([hh, ss, vv], -1).astype(np.uint8).reshape(-1, 3)
    n = (flat.shape[0] // width) * width
    im = np.ascontiguousarray(flat[:n].reshape(-1, width, 3))
    got = ctx.download(ctx.cvt_color(ctx.upload(im), "hsv2bgr"))
    assert np.array_equal(got, cv2.cvtColor(im, cv2.COLOR_HSV2BGR))


def test_hls_all_colors(ctx, all_colors):
    """P1 conversion, bit-exact over all 2^24 colours since round 2 (cv2's vector path wraps a negative hue inside the
    multiply-add, which decides three rounding ties); the row-tail path on an odd width."""
    got = ctx.download(ctx.cvt_color(ctx.upload(all_colors), "bgr2hls"))
    assert np.array_equal(got, cv2.cvtColor(all_colors, cv2.COLOR_BGR2HLS))
    tail = np.ascontiguousarray(all_colors[:64, :4000].reshape(-1, 25, 3))
    assert np.array_equal(ctx.download(ctx.cvt_color(ctx.upload(tail), "bgr2hls")), cv2.cvtColor(tail, cv2.COLOR_BGR2HLS))


@pytest.mark.parametrize("shape", [(479, 641), (1, 1), (3, 5), (17, 16), (1243, 2209)])
@pytest.mark.parametrize("name,cvc", CODES)
def test_cvt_odd_shapes(ctx, shape, name, cvc):
    img = synth.gen_random_bgr(shape[0], shape[1], 11)
    assert np.array_equal(ctx.download(ctx.cvt_color(ctx.upload(img), name)), cv2.cvtColor(img, cvc))


def test_cvt_batch_and_misaligned_views(ctx):
    batch = np.stack([synth.gen_random_bgr(37, 53, s) for s in range(5)])
    d = ctx.upload(batch)
    got = ctx.download(ctx.cvt_color(d, "bgr2lab"))
    for i in range(5):
        assert np.array_equal(got[i], cv2.cvtColor(batch[i], cv2.COLOR_BGR2LAB))
    # frame 1 of this batch starts at an odd byte offset: exercises the scalar kernels
    one = d[1]
    assert one.data_ptr() % 16 != 0
    assert np.array_equal(ctx.download(ctx.cvt_color(one.contiguous(), "bgr2hsv")), cv2.cvtColor(batch[1], cv2.COLOR_BGR2HSV))


def test_gray2bgr_and_bgr2rgb(ctx):
    img = synth.gen_random_bgr(40, 72, 2)
    g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(ctx.download(ctx.cvt_color(ctx.upload(g), "gray2bgr")), cv2.cvtColor(g, cv2.COLOR_GRAY2BGR))
    assert np.array_equal(ctx.download(ctx.cvt_color(ctx.upload(img), "bgr2rgb")), img[..., ::-1])


def test_mirror_of_utils_color(ctx):
    """Same call shapes as utils/color.py: (converted, [planes])."""
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_underwater(120, 160, 8)
    for fn, name in ((color.bgr_to_lab, "bgr2lab"), (color.bgr_to_hsv, "bgr2hsv"), (color.bgr_to_ycrcb, "bgr2ycrcb")):
        conv, planes = fn(img)
        rconv, rplanes = cv_ops.convert(img, name)
        assert np.array_equal(conv, rconv) and len(planes) == 3
        assert all(np.array_equal(a, b) for a, b in zip(planes, rplanes))
    gray, _ = color.bgr_to_gray(img)
    assert np.array_equal(gray, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    assert np.array_equal(color.hsv_to_bgr(hsv)[0], cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
    lab_a = rplanes[1] if False else cv_ops.convert(img, "bgr2lab")[1][1]
    assert np.array_equal(color.range_threshold(lab_a, 120, 140), cv2.inRange(lab_a, 120, 140))


@pytest.mark.parametrize("shape", [(1080, 1920), (479, 641), (5, 7)])
def test_in_range_three_channel(ctx, shape):
    img = synth.gen_underwater(shape[0], shape[1], 4)
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    lo, hi = np.array([10, 20, 60]), np.array([30, 100, 255])        # modules/bins.py:14-15
    assert np.array_equal(ctx.download(ctx.in_range(ctx.upload(hsv), lo, hi)), cv2.inRange(hsv, lo, hi))
    # fused convert + inRange
    from cuauv_vision_pipeline_b200.runtime import ffi, lib, check, _u8ptr
    d = ctx.upload(img)
    m = ctx.empty(shape)
    lo8, hi8 = lo.astype(np.uint8), hi.astype(np.uint8)
    check(lib.bv_cvt_in_range(ctx.handle, _u8ptr(d), _u8ptr(m), 1, shape[0], shape[1], lib.BV_BGR2HSV,
                              ffi.from_buffer("uint8_t[]", lo8), ffi.from_buffer("uint8_t[]", hi8)))
    assert np.array_equal(ctx.download(m), cv2.inRange(hsv, lo, hi))


def test_in_range_single_channel_and_edge_bounds(ctx):
    img = synth.gen_random_bgr(123, 77, 9)[..., 0].copy()
    d = ctx.upload(img)
    for lo, hi in ((0, 255), (100, 100), (200, 50), (0, 0), (255, 255), (-5, 300), (256, 300)):
        assert np.array_equal(ctx.download(ctx.in_range(d, lo, hi)), cv2.inRange(img, lo, hi)), (lo, hi)


@pytest.mark.parametrize("kind,cvt", [("binary", cv2.THRESH_BINARY), ("binary_inv", cv2.THRESH_BINARY_INV),
                                      ("trunc", cv2.THRESH_TRUNC), ("tozero", cv2.THRESH_TOZERO),
                                      ("tozero_inv", cv2.THRESH_TOZERO_INV)])
def test_thresholds(ctx, kind, cvt):
    img = synth.gen_random_bgr(97, 131, 1)[..., 1].copy()
    d = ctx.upload(img)
    for t in (0, 1, 127, 128, 254, 255):
        mv = 255 if kind.startswith("binary") else 0
        assert np.array_equal(ctx.download(ctx.threshold(d, t, mv, kind)), cv2.threshold(img, t, mv, cvt)[1]), t


def test_threshold_mirrors(ctx):
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_random_bgr(64, 64, 3)[..., 2].copy()
    assert np.array_equal(color.binary_threshold(img, 99), cv_ops.binary_threshold(img, 99))
    assert np.array_equal(color.binary_threshold_inv(img, 99), cv_ops.binary_threshold_inv(img, 99))
    assert np.array_equal(color.max_threshold(img, 99), cv_ops.max_threshold(img, 99))
    assert np.array_equal(color.above_threshold(img, 99), cv_ops.above_threshold(img, 99))
    assert np.array_equal(color.below_threshold(img, 99), cv_ops.below_threshold(img, 99))


@pytest.mark.parametrize("shape", [(242, 368), (479, 641)])
def test_apply_lut_carries_preprocessor_point_ops(ctx, shape):
    from cuauv_vision_pipeline_b200.preprocessor import point_lut
    img = synth.gen_random_bgr(shape[0], shape[1], 6)
    ref = cv_ops.brightness(cv_ops.contrast(cv_ops.channel_bias(img, 2, 25), 1.7), -40)
    got = ctx.download(ctx.apply_lut(ctx.upload(img), point_lut(25, 0, 0, 1.7, -40)))
    assert np.array_equal(got, ref)
    one = img[..., 0].copy()
    lut1 = np.arange(255, -1, -1, dtype=np.uint8)
    assert np.array_equal(ctx.download(ctx.apply_lut(ctx.upload(one), lut1)), lut1[one])


TCD_CASES = [dict(color=(120, 150, 140), distance=30), dict(color=(60, 128, 128), distance=45, weights=(0.2, 1, 1)),
             dict(color=(200, 110, 170), distance=25.5, ignore_channels=[0]),
             dict(color=(10, 240, 20), distance=400, weights=(3, 1, 2))]


def test_thresh_color_distance_golden_from_reference_function(ctx):
    """utils/color.py:66-103 (P1): golden outputs of the reference function itself."""
    import os
    from cuauv_vision_pipeline_b200 import color
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv_calls_72x128.npz"))
    split = list(cv2.split(z["bgr2lab"]))
    for k, kw in enumerate(TCD_CASES):
        mask, dist = color.thresh_color_distance(split, **kw)
        assert np.array_equal(mask, z["tcd%d_mask" % k]) and np.array_equal(dist, z["tcd%d_dist" % k]), k


@pytest.mark.parametrize("shape", [(479, 641), (1080, 1920)])
def test_thresh_color_distance_vs_oracle(ctx, shape):
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_underwater(shape[0], shape[1], 13)
    split = list(cv2.split(cv2.cvtColor(img, cv2.COLOR_BGR2LAB)))
    for kw in TCD_CASES + [dict(color=(0, 0, 0), distance=500, weights=(1, 1, 1))]:
        mask, dist = color.thresh_color_distance(split, **kw)
        rmask, rdist = cv_ops.thresh_color_distance(split, **kw)
        assert np.array_equal(mask, rmask) and np.array_equal(dist, rdist), kw
    # auto_distance_percentile (utils/color.py:98-99): min(np.percentile(dists, p), distance**2)
    for kw in (dict(color=(120, 150, 140), distance=300, auto_distance_percentile=50),
               dict(color=(60, 128, 128), distance=45, auto_distance_percentile=1.5, weights=(0.2, 1, 1)),
               dict(color=(200, 110, 170), distance=1000, auto_distance_percentile=99.9, ignore_channels=[0]),
               dict(color=(100, 140, 135), distance=5, auto_distance_percentile=33.3),    # distance**2 is the smaller one
               dict(color=(90, 120, 130), distance=1e4, auto_distance_percentile=100)):
        mask, dist = color.thresh_color_distance(split, **kw)
        rmask, rdist = cv_ops.thresh_color_distance(split, **kw)
        assert np.array_equal(mask, rmask) and np.array_equal(dist, rdist), kw


def test_select_kth_is_the_order_statistic(ctx):
    rng = np.random.default_rng(5)
    for vals in (rng.normal(0, 100, 100003).astype(np.float32), (rng.random(4097) ** 3 * 1e6).astype(np.float32),
                 np.array([3.0, -0.0, 0.0, -7.5, 3.0, np.inf, -np.inf, 1e-40], np.float32), np.zeros(10, np.float32)):
        d = ctx.upload(vals)
        s = np.sort(vals)
        for k in (0, 1, len(vals) // 2, len(vals) - 2, len(vals) - 1):
            got = ctx.select_kth(d, k)
            assert got == s[k] and np.signbit(got) == np.signbit(s[k]), (k, got, s[k])


@pytest.mark.parametrize("shape", [(480, 640, 3), (479, 641, 3), (1242, 2208, 3), (97, 131), (5, 4, 3)])
@pytest.mark.parametrize("n", [3, 5, 9, 13, 31])
def test_gaussian_blur_vs_cv2(ctx, shape, n):
    """modules/preprocessor.py:110-114: cv2.GaussianBlur(mat, (2k+1, 2k+1), 0), bit-exact."""
    from cuauv_vision_pipeline_b200 import transform
    img = synth.gen_underwater(shape[0], shape[1], n)
    if len(shape) == 2:
        img = np.ascontiguousarray(img[..., 1])
    assert np.array_equal(transform.gaussian_blur(img, (n, n)), cv2.GaussianBlur(img, (n, n), 0))


@pytest.mark.parametrize("n", [33, 63, 101, 151, 201])
def test_gaussian_blur_up_to_the_tuner_maximum(ctx, n):
    """PPX_gaussian_blur_kernel spans 1..100, i.e. kernels up to 201 x 201 (modules/preprocessor.py:25,110-114); the
    largest ones leave the shared-memory tile for the two-launch path, also on images smaller than the kernel."""
    for shape in [(240, 320), (97, 131)]:
        img = synth.gen_underwater(shape[0], shape[1], n)
        assert np.array_equal(ctx.download(ctx.gaussian_blur(ctx.upload(img), (n, n))), cv2.GaussianBlur(img, (n, n), 0)), shape
    img = synth.gen_random_bgr(150, 210, n)
    assert np.array_equal(ctx.download(ctx.gaussian_blur(ctx.upload(img), (n, 3))), cv2.GaussianBlur(img, (n, 3), 0))
    grey = np.ascontiguousarray(img[..., 0])
    assert np.array_equal(ctx.download(ctx.gaussian_blur(ctx.upload(grey), (5, n))), cv2.GaussianBlur(grey, (5, n), 0))


def test_gaussian_blur_rectangular_kernels_sigmas_and_batches(ctx):
    img = synth.gen_random_bgr(200, 333, 9)
    for (kw, kh, sx, sy) in [(7, 3, 0, 0), (1, 9, 0, 0), (5, 5, 2.5, 0), (11, 7, 1.2, 3.3), (63, 1, 0, 0)]:
        got = ctx.download(ctx.gaussian_blur(ctx.upload(img), (kw, kh), sx, sy))
        assert np.array_equal(got, cv2.GaussianBlur(img, (kw, kh), sx, sigmaY=sy)), (kw, kh, sx, sy)
    batch = np.stack([synth.gen_underwater(120, 200, s) for s in range(3)])
    got = ctx.download(ctx.gaussian_blur(ctx.upload(batch), (5, 5)))
    for i in range(3):
        assert np.array_equal(got[i], cv2.GaussianBlur(batch[i], (5, 5), 0))


@pytest.mark.parametrize("shape", [(480, 640, 3), (479, 641, 3), (1242, 2208, 3), (97, 131)])
def test_rotate_and_translate_vs_cv2(ctx, shape):
    """modules/preprocessor.py:130-135 (rotate about the centre, BORDER_REPLICATE) and 144-149 (translate)."""
    from cuauv_vision_pipeline_b200 import transform
    img = synth.gen_underwater(shape[0], shape[1], 21)
    if len(shape) == 2:
        img = np.ascontiguousarray(img[..., 2])
    h, w = shape[:2]
    for ang in (10, -33.3, 90, 180, 0.5, 359):
        m = cv2.getRotationMatrix2D((w / 2, h / 2), ang, 1)
        assert np.array_equal(transform.rotation_matrix_2d((w / 2, h / 2), ang, 1), m)
        assert np.array_equal(transform.rotate(img, ang), cv2.warpAffine(img, m, (w, h), borderMode=cv2.BORDER_REPLICATE)), ang
    for tx, ty in ((25, 0), (0, -17), (-300, 40), (3.5, -2.25)):
        m = np.float32([[1, 0, tx], [0, 1, ty]])
        assert np.array_equal(transform.translate(img, tx, ty), cv2.warpAffine(img, m, (w, h))), (tx, ty)


def test_warp_affine_general_matrix_dsize_and_border_value(ctx):
    from cuauv_vision_pipeline_b200 import transform
    img = synth.gen_underwater(300, 400, 5)
    m = np.array([[0.8, 0.3, 10.5], [-0.2, 1.1, -4.25]])
    assert np.array_equal(transform.warp_affine(img, m, (500, 210), "constant", (7, 99, 200)),
                          cv2.warpAffine(img, m, (500, 210), borderValue=(7, 99, 200)))
    assert np.array_equal(transform.warp_affine(img, m, (123, 457), "replicate"),
                          cv2.warpAffine(img, m, (123, 457), borderMode=cv2.BORDER_REPLICATE))
    sing = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 5.0]])              # singular: cv2 uses D = 0
    assert np.array_equal(transform.warp_affine(img, sing), cv2.warpAffine(img, sing, (400, 300)))


def test_lab_to_bgr_mirror_odd_shape_and_round_trip(ctx):
    """utils/color.py:27-29 lab_to_bgr (P1): OpenCV's fixed-point Lab2RGBinteger, bit-exact incl. odd widths and planes."""
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_underwater(479, 641, 31)
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB)
    conv, planes = color.lab_to_bgr(lab)
    ref = cv2.cvtColor(lab, cv2.COLOR_LAB2BGR)
    assert np.array_equal(conv, ref)
    for k in range(3):
        assert np.array_equal(planes[k], ref[..., k])
    rnd = synth.gen_random_bgr(77, 129, 2)                       # arbitrary (L, a, b) triples, out-of-gamut included
    assert np.array_equal(color.lab_to_bgr(rnd)[0], cv2.cvtColor(rnd, cv2.COLOR_LAB2BGR))


@pytest.mark.parametrize("shape,seed", [((480, 640), 1), ((1242, 2208), 2), ((479, 641), 3)])
def test_white_balance_bgr_vs_reference_expression(ctx, shape, seed):
    """utils/color.py:370-379 (a11, P2).  Stated tolerance: identical unless the a / b mean lies within 1e-5 of an
    integer (the reference's float32 pairwise mean vs the exact mean here); identical on these frames."""
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_underwater(shape[0], shape[1], seed)
    lab_img = cv2.cvtColor(img, cv2.COLOR_BGR2LAB).astype(np.float32)
    lab_l, lab_a, lab_b = cv2.split(lab_img)
    lab_a -= np.mean(lab_a) - 128
    lab_b -= np.mean(lab_b) - 128
    ref = cv2.cvtColor(cv2.merge((lab_l, lab_a, lab_b)).astype(np.uint8), cv2.COLOR_LAB2BGR)
    assert np.array_equal(color.white_balance_bgr(img), ref)


@pytest.mark.parametrize("shape,ksize,seed", [((480, 640), 15, 1), ((1242, 2208), 31, 2), ((479, 641), 4, 3),
                                              ((37, 53), 101, 4), ((64, 64), 1, 5)])
def test_white_balance_bgr_blur_bit_exact(ctx, shape, ksize, seed):
    """utils/color.py:381-391 (a11, P2), the reference's own expressions; kernels larger than the image and even
    sizes (rounded up to odd by the reference) included."""
    from cuauv_vision_pipeline_b200 import color
    img = synth.gen_underwater(shape[0], shape[1], seed) if seed != 4 else synth.gen_random_bgr(shape[0], shape[1], seed)
    k = 2 * (ksize // 2) + 1
    lab_img = cv2.cvtColor(img, cv2.COLOR_BGR2LAB).astype(np.float32)
    lab_l, lab_a, lab_b = cv2.split(lab_img)
    lab_a_avg = cv2.blur(lab_a, (k, k), 0, borderType=cv2.BORDER_REPLICATE)
    lab_b_avg = cv2.blur(lab_b, (k, k), 0, borderType=cv2.BORDER_REPLICATE)
    lab_a -= lab_a_avg - 128
    lab_b -= lab_b_avg - 128
    with np.errstate(invalid="ignore"):
        ref = cv2.cvtColor(cv2.merge((lab_l, lab_a, lab_b)).astype(np.uint8), cv2.COLOR_LAB2BGR)
    assert np.array_equal(color.white_balance_bgr_blur(img, ksize), ref)


def test_bgr2luv_all_colors(ctx, all_colors):
    """utils/color.py:30 bgr_to_luv / modules/preprocessor.py:76-80 (P1).  Bit-exact over all 2^24 colours since round 2: the
    node table is the host libm's plus 97 fitted entries (csrc/luv_fix.inc) that make it OpenCV's own.  (The table is built
    with powf / cbrtf of the host at run time: with a libm that rounds other nodes differently the stated bound is <= 1 LSB.)"""
    from cuauv_vision_pipeline_b200 import color
    got = ctx.download(ctx.cvt_color(ctx.upload(all_colors), "bgr2luv"))
    ref = cv2.cvtColor(all_colors, cv2.COLOR_BGR2LUV)
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert int(d.max()) <= 1, "stated bound"
    assert int((d != 0).any(axis=2).sum()) == 0, "expected: identical bytes"
    img = synth.gen_underwater(479, 641, 5)
    conv, planes = color.bgr_to_luv(img)
    assert np.array_equal(conv, cv2.cvtColor(img, cv2.COLOR_BGR2LUV))
    for k in range(3):
        assert np.array_equal(planes[k], conv[..., k])


_CAM_K = np.array([[904.66192735, 0.0, 481.17596262], [0.0, 902.84000422, 404.82437525], [0.0, 0.0, 1.0]])
_CAM_D = np.array([0.48525658, 2.02550297, 0.03807578, -0.02152142, -3.30299241])   # lib/configs/1_camera_matrix_params.yaml


@pytest.mark.parametrize("shape", [(240, 320), (479, 641), (1242, 2208)])
def test_remap_bit_exact_vs_cv2(ctx, shape):
    """bv_remap == cv2.remap(INTER_LINEAR) on uint8: float32 maps and the fixed-point pair, both border modes, 1 and 3
    channels, a batch sharing one pair of maps, maps leaving the frame on every side, a destination smaller than the source."""
    from cuauv_vision_pipeline_b200 import transform
    h, w = shape
    img = synth.gen_underwater(h, w, 8)
    rng = np.random.default_rng(h)
    mx = (np.arange(w, dtype=np.float32)[None, :] * 1.03 - 9 + rng.normal(0, 3, (h, w))).astype(np.float32)
    my = (np.arange(h, dtype=np.float32)[:, None] * 1.02 - 7 + rng.normal(0, 3, (h, w))).astype(np.float32)
    for border, mode in (("constant", cv2.BORDER_CONSTANT), ("replicate", cv2.BORDER_REPLICATE)):
        assert np.array_equal(transform.remap(img, mx, my, border=border), cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=mode))
    gray = np.ascontiguousarray(img[..., 2])
    assert np.array_equal(transform.remap(gray, mx, my), cv2.remap(gray, mx, my, cv2.INTER_LINEAR))
    assert np.array_equal(transform.remap(img, mx, my, border_value=(7, 8, 9)),
                          cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(7, 8, 9)))
    m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    assert np.array_equal(transform.remap(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    small = (slice(0, h // 2), slice(0, w // 3))
    assert np.array_equal(transform.remap(img, np.ascontiguousarray(mx[small]), np.ascontiguousarray(my[small])),
                          cv2.remap(img, np.ascontiguousarray(mx[small]), np.ascontiguousarray(my[small]), cv2.INTER_LINEAR))
    batch = np.stack([img, img[::-1].copy()])
    got = ctx.download(ctx.remap(ctx.upload(batch), torch.from_numpy(mx).to(ctx.device), torch.from_numpy(my).to(ctx.device)))
    for k in range(2):
        assert np.array_equal(got[k], cv2.remap(batch[k], mx, my, cv2.INTER_LINEAR))


def test_undistort_with_the_reference_camera_file(ctx):
    """include/camera_filters.hpp:6-11 (maps built once per camera, remap per frame) with lib/configs/1_camera_matrix_params.yaml:
    the device maps equal cv2.initUndistortRectifyMap's (float32 and fixed point), the undistorted frame equals cv2.undistort's."""
    from cuauv_vision_pipeline_b200 import transform
    for size in ((964, 724), (640, 480)):
        new_k, _ = cv2.getOptimalNewCameraMatrix(_CAM_K, _CAM_D, size, 1)
        img = synth.gen_underwater(size[1], size[0], 9)
        for nk in (None, new_k):
            ref_k = _CAM_K if nk is None else nk
            mx, my = transform.init_undistort_rectify_map(_CAM_K, _CAM_D, None, nk, size)
            rx, ry = cv2.initUndistortRectifyMap(_CAM_K, _CAM_D, None, ref_k, size, cv2.CV_32FC1)
            assert np.array_equal(mx.cpu().numpy(), rx) and np.array_equal(my.cpu().numpy(), ry)
            assert np.array_equal(transform.remap(img, mx, my), cv2.remap(img, rx, ry, cv2.INTER_LINEAR))
            f1, f2 = transform.init_undistort_rectify_map(_CAM_K, _CAM_D, None, nk, size, fixed=True)
            r1, r2 = cv2.initUndistortRectifyMap(_CAM_K, _CAM_D, None, ref_k, size, cv2.CV_16SC2)
            assert np.array_equal(f1.cpu().numpy(), r1) and np.array_equal(f2.cpu().numpy().view(np.uint16), r2)
            ref = cv2.undistort(img, _CAM_K, _CAM_D, None, ref_k)
            assert np.array_equal(transform.undistort(img, _CAM_K, _CAM_D, nk), ref)
            assert np.array_equal(transform.undistort(img, _CAM_K, _CAM_D, nk, maps=(f1, f2)), ref)     # maps reused across frames
    rot, _ = cv2.Rodrigues(np.array([0.02, -0.03, 0.01]))
    d8 = np.array([0.1, -0.2, 0.001, 0.002, 0.05, 0.01, -0.02, 0.003])
    mx, my = transform.init_undistort_rectify_map(_CAM_K, d8, rot, _CAM_K, (320, 240))
    rx, ry = cv2.initUndistortRectifyMap(_CAM_K, d8, rot, _CAM_K, (320, 240), cv2.CV_32FC1)
    assert np.array_equal(mx.cpu().numpy(), rx) and np.array_equal(my.cpu().numpy(), ry)


def test_init_undistort_map_flow(ctx):
    """The flow include/camera_filters.hpp:6-11 declares: camera file -> optimal new camera matrix -> maps (once), remap per
    frame; equals getOptimalNewCameraMatrix + initUndistortRectifyMap(CV_16SC2) + remap of cv2."""
    from cuauv_vision_pipeline_b200 import transform
    w, h = 964, 724
    img = synth.gen_underwater(h, w, 4)
    for alpha in (0.0, 1.0):
        maps = transform.init_undistort_map((_CAM_K, _CAM_D), w, h, alpha=alpha)
        new_k, _ = cv2.getOptimalNewCameraMatrix(_CAM_K, _CAM_D, (w, h), alpha)
        r1, r2 = cv2.initUndistortRectifyMap(_CAM_K, _CAM_D, None, new_k, (w, h), cv2.CV_16SC2)
        assert np.array_equal(transform.remap(img, *maps), cv2.remap(img, r1, r2, cv2.INTER_LINEAR))


def test_mirrors_order_device_tensors_against_the_callers_stream(ctx):
    """a CUDA tensor produced on the caller's torch stream must be complete before the context's own
    stream reads it, and a returned tensor must be complete before the caller's stream uses it."""
    import torch
    from cuauv_vision_pipeline_b200 import color
    from cuauv_vision_pipeline_b200.color_balance import balance
    img = synth.gen_underwater(1080, 1920, 5)
    want_lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB)
    side = torch.cuda.Stream()
    pinned = torch.from_numpy(img).pin_memory()
    for rep in range(5):
        with torch.cuda.stream(side):
            # a long producer on the caller's stream: the frame is only correct after 40 passes of in-place arithmetic
            d = pinned.to("cuda", non_blocking=True)
            acc = d.to(torch.int32)
            for _ in range(40):
                acc = acc + 3
            for _ in range(40):
                acc = acc - 3
            d = acc.to(torch.uint8)
            lab, planes = color.bgr_to_lab(d)            # consumer: the context's stream
            host = (lab.to(torch.int32) + 0).to(torch.uint8).cpu()   # caller's stream again
            bal = balance(d)
            bal_host = bal.cpu()
        side.synchronize()
        assert np.array_equal(host.numpy(), want_lab)
        if rep == 0:
            first = bal_host.numpy().copy()
        assert np.array_equal(bal_host.numpy(), first)


def test_in_range_reads_its_bounds_like_cv2(ctx):
    """a bare scalar bound on a 3-channel image is cv::Scalar(v, 0, 0, 0); fractional bounds are cvRound'ed
    (half to even); bounds outside [0, 255] saturate or empty the range."""
    from cuauv_vision_pipeline_b200 import color
    img = synth.all_colors_image()[:512, :512].copy()
    grey = np.ascontiguousarray(img[..., 1])
    cases3 = [(10, 20), ((10, 10, 10), (20, 20, 20)), ((10.5, 0.5, 1.5), (20.5, 99.5, 254.6)), ((-3.2, 0, 0), (5, 300.7, 255)),
              ((0, 0, 0), (0, 0, 0)), ((20.6, 0, 0), (20.4, 255, 255)), (0, 255)]
    for lo, hi in cases3:
        want = cv2.inRange(img, np.array(lo, dtype=np.float64) if np.ndim(lo) else lo, np.array(hi, dtype=np.float64) if np.ndim(hi) else hi)
        assert np.array_equal(color.range_threshold(img, lo, hi), want), (lo, hi)
    for lo, hi in [(10, 20), (10.5, 20.5), (10.4, 20.6), (-3.2, 5.0), (250.1, 300.7), (19.999, 20.0001), (20.6, 20.4)]:
        assert np.array_equal(color.range_threshold(grey, lo, hi), cv2.inRange(grey, lo, hi)), (lo, hi)
    # make_stage uses the same convention
    d = ctx.make_stage(cvt=None, lo=10, hi=20)
    assert np.array_equal(ctx.download(ctx.stage(d, ctx.upload(img), want=("mask",))["mask"]), cv2.inRange(img, 10, 20))
