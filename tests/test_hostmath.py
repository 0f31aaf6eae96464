"""The per-pixel arithmetic of the CUDA kernels (csrc/pixel_math.cuh, __host__ __device__) compiled
for the host and swept against cv2: all 2^24 colours per conversion, every (H<180,S,V) for
HSV2BGR incl. the 32-px vector / tail rounding rule, and the fixed-point bilinear resize.  This
checks the exact source the device runs, in a container without a GPU; the GPU tests repeat the
sweeps through the real kernels."""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmath", "hostmath.cpp")
OUT = os.path.join(HERE, "hostmath", "_build", "libhostmath.so")


@pytest.fixture(scope="module")
def hm():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    csrc = os.path.join(HERE, "..", "cuauv_vision_pipeline_b200", "csrc")
    deps = [SRC, os.path.join(csrc, "pixel_math.cuh"), os.path.join(csrc, "luv_fix.inc")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def convert(hm, img, code, one=False):
    h, w = img.shape[:2]
    out = np.empty((h, w) if one else (h, w, 3), np.uint8)
    rc = hm.hm_convert(img.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_size_t(h * w), w, code)
    assert rc == 0
    return out


@pytest.mark.parametrize("name,code,cvc", [("hsv", 0, cv2.COLOR_BGR2HSV), ("lab", 1, cv2.COLOR_BGR2LAB),
                                           ("gray", 2, cv2.COLOR_BGR2GRAY), ("ycrcb", 3, cv2.COLOR_BGR2YCrCb),
                                           ("lab2bgr", 8, cv2.COLOR_LAB2BGR)])
def test_device_math_all_colors(hm, name, code, cvc):
    img = synth.all_colors_image()
    assert np.array_equal(convert(hm, img, code, one=(code == 2)), cv2.cvtColor(img, cvc))


@pytest.mark.parametrize("width", [4096, 100, 63, 33])
def test_device_math_hsv2bgr(hm, width):
    hh, ss, vv = np.meshgrid(np.arange(180), np.arange(256), np.arange(256), indexing="ij")
    flat = np.stack([hh, ss, vv], -1).astype(np.uint8).reshape(-1, 3)
    n = (flat.shape[0] // width) * width
    im = np.ascontiguousarray(flat[:n].reshape(-1, width, 3))
    assert np.array_equal(convert(hm, im, 4), cv2.cvtColor(im, cv2.COLOR_HSV2BGR))


def test_device_math_hls_all_colors(hm):
    """BGR2HLS is P1: identical to cv2 on all 2^24 colours in the vector path since round 2 (cv2 wraps a negative hue with
    the product still unrounded, fma(g - b, k, 360): three colours sit on that rounding tie)."""
    img = synth.all_colors_image()
    mine = convert(hm, img, 5)
    ref = cv2.cvtColor(img, cv2.COLOR_BGR2HLS)
    assert np.array_equal(mine, ref), int((mine != ref).any(axis=2).sum())
    # row tail (width % 32 columns): cv2's scalar path, separately rounded multiply and add
    tail = np.ascontiguousarray(img[:64, :4000].reshape(-1, 25, 3))
    assert np.array_equal(convert(hm, tail, 5), cv2.cvtColor(tail, cv2.COLOR_BGR2HLS))


def test_hsv_division_tables(hm):
    sdiv = (ctypes.c_int * 256)()
    hdiv = (ctypes.c_int * 256)()
    hm.hm_hsv_tables(sdiv, hdiv)
    i = np.arange(1, 256, dtype=np.float64)
    assert list(sdiv)[1:] == np.rint((255 << 12) / i).astype(int).tolist() and sdiv[0] == 0
    assert list(hdiv)[1:] == np.rint((180 << 12) / (6 * i)).astype(int).tolist() and hdiv[0] == 0


@pytest.mark.parametrize("shape", [(1242, 2208, 360, 640, 3), (479, 641, 777, 333, 3), (480, 640, 960, 1280, 1),
                                   (100, 100, 37, 53, 3), (2160, 3840, 360, 640, 3)])
def test_device_math_resize(hm, shape):
    sh, sw, dh, dw, c = shape
    im = np.random.default_rng(7).integers(0, 256, (sh, sw, c), dtype=np.uint8)
    ref = cv2.resize(im, (dw, dh), interpolation=cv2.INTER_LINEAR).reshape(dh, dw, c)
    out = np.empty((dh, dw, c), np.uint8)
    hm.hm_resize(im.ctypes.data_as(ctypes.c_void_p), sh, sw, out.ctypes.data_as(ctypes.c_void_p), dh, dw, c)
    assert np.array_equal(out, ref)


def test_device_math_bgr2luv_all_colors(hm):
    """utils/color.py:30 bgr_to_luv (P1).  OpenCV interpolates a 33^3 int16 node table that it fills with its softfloat
    pow / cubeRoot; here the table comes from the host libm plus 97 fitted node values (csrc/luv_fix.inc: round 1 fitted 94
    node by node, round 2 solved the rest as integer programs over neighbouring nodes).  With this image's glibc all 2^24
    colours are identical to cv2; the stated bound for a libm that rounds other nodes differently is <= 1 LSB."""
    img = synth.all_colors_image()
    got = convert(hm, img, 9).astype(np.int16)
    ref = cv2.cvtColor(img, cv2.COLOR_BGR2LUV).astype(np.int16)
    d = np.abs(got - ref)
    assert int(d.max()) <= 1
    assert int((d != 0).any(axis=2).sum()) == 0, int((d != 0).any(axis=2).sum())
    print("BGR2LUV colours off by one:", int((d != 0).any(axis=2).sum()))
