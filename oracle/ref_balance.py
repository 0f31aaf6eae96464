"""TEST INFRASTRUCTURE ONLY -- the reference's own `process_frame`, compiled unmodified.

Loads oracle/_ref/libauv-color-balance-ref.so (built by oracle/Makefile from
/root/reference/utils/color_correction/color_balance.cpp) and calls it with the exact marshalling
of the reference's `balance()` (modules/color_balance.py:93-110): a flattened copy of the frame, a
`c_int8*` pointer, python ints / bools passed without argtypes.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libauv-color-balance-ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(
                f"{REF_LIB_PATH} missing: run `make -C oracle` in a container that has /root/reference")
        _lib = ctypes.CDLL(REF_LIB_PATH)
    return _lib


def balance(mat, equalize_rgb=True, rgb_contrast_correct=False,
            hsv_contrast_correct=True, hsi_contrast_correct=False,
            rgb_extrema_clipping=True, adaptive_cast_correction=False,
            horizontal_blocks=1, vertical_blocks=1):
    """Same signature, defaults and marshalling as the reference `balance()`."""
    lib = _load()
    rows = mat.shape[0]
    cols = mat.shape[1]
    depth = 3
    c_int8_p = ctypes.POINTER(ctypes.c_int8)
    if hsi_contrast_correct:
        # The reference's HSI branch is not repeatable on a cold start: its eight conversion threads share an unsynchronised
        # memo table (rgb_to_hsi_cache, color_balance.cpp:180-185 read, 208-212 write), so a thread can see entry [0] of a
        # colour as valid while [1] / [2] are still being written by another one -- one pixel of the frame then carries a
        # stale S or I (seen on ~4 % of first calls in a process; tests/test_gpu_balance_stage.py reports which side moved).
        # A first pass over a throw-away copy fills the table for every colour of this frame; the second pass only reads
        # complete entries and is what the oracle returns.
        warm = mat.flatten()
        lib.process_frame(warm.ctypes.data_as(c_int8_p), rows, cols, depth, equalize_rgb,
                          rgb_contrast_correct, hsv_contrast_correct, hsi_contrast_correct,
                          rgb_extrema_clipping, adaptive_cast_correction,
                          horizontal_blocks, vertical_blocks)
    data = mat.flatten()
    data_p = data.ctypes.data_as(c_int8_p)
    lib.process_frame(data_p, rows, cols, depth, equalize_rgb,
                      rgb_contrast_correct, hsv_contrast_correct, hsi_contrast_correct,
                      rgb_extrema_clipping, adaptive_cast_correction,
                      horizontal_blocks, vertical_blocks)
    return np.ctypeslib.as_array(data_p, (rows, cols, depth)).astype(np.uint8)


# ---- timed reference arm of bench.py -------------------------------------------------------------------------------
FAST_LIB_PATH = os.path.join(_HERE, "_ref", "libauv-color-balance-ref-avx2.so")
_fast = None
_hook_keepalive = None


def _cpu_has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2 " in f.read().replace("\n", " ")
    except OSError:
        return False


def fastest():
    """(library, description): the reference's translation unit as fast as it gets on this host -- the -O3 -mavx2 build when
    the CPU has AVX2, with cv::cvtColor delegated to REAL OpenCV (cv2.cvtColor, SIMD) through the shim's hook.  Used only as
    the timed CPU baseline; `balance()` above (scalar shim, -O2) stays the parity oracle."""
    global _fast, _hook_keepalive
    if _fast is not None:
        return _fast
    import cv2
    path, what = REF_LIB_PATH, "-O2"
    if os.path.exists(FAST_LIB_PATH) and _cpu_has_avx2():
        path, what = FAST_LIB_PATH, "-O3 -mavx2"
    lib = ctypes.CDLL(path)
    try:
        hook_t = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int)

        def hook(src, dst, rows, cols, code):
            n = rows * cols * 3
            s = np.ctypeslib.as_array(ctypes.cast(src, ctypes.POINTER(ctypes.c_uint8)), (n,)).reshape(rows, cols, 3)
            d = np.ctypeslib.as_array(ctypes.cast(dst, ctypes.POINTER(ctypes.c_uint8)), (n,)).reshape(rows, cols, 3)
            if src == dst:
                d[...] = cv2.cvtColor(s, code)
            else:
                cv2.cvtColor(s, code, dst=d)
        _hook_keepalive = hook_t(hook)
        ctypes.c_void_p.in_dll(lib, "bv_shim_cvtcolor_hook").value = ctypes.cast(_hook_keepalive, ctypes.c_void_p).value
        what += ", cv::cvtColor = real cv2.cvtColor (OpenCV %s SIMD)" % cv2.__version__
    except (ValueError, AttributeError):
        what += ", scalar cv shim"
    _fast = (lib, what)
    return _fast


def balance_timed(mat):
    """balance() with the default flags through `fastest()` (same marshalling)."""
    lib, _ = fastest()
    rows, cols = mat.shape[0], mat.shape[1]
    data = mat.flatten()
    data_p = data.ctypes.data_as(ctypes.POINTER(ctypes.c_int8))
    lib.process_frame(data_p, rows, cols, 3, True, False, True, False, True, False, 1, 1)
    return np.ctypeslib.as_array(data_p, (rows, cols, 3)).astype(np.uint8)
