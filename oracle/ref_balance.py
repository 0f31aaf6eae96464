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
    data = mat.flatten()
    data_p = data.ctypes.data_as(c_int8_p)
    lib.process_frame(data_p, rows, cols, depth, equalize_rgb,
                      rgb_contrast_correct, hsv_contrast_correct, hsi_contrast_correct,
                      rgb_extrema_clipping, adaptive_cast_correction,
                      horizontal_blocks, vertical_blocks)
    return np.ctypeslib.as_array(data_p, (rows, cols, depth)).astype(np.uint8)
