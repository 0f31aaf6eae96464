"""GPU mirror of the reference's modules/color_balance.py::balance (lines 93-110)."""
import ctypes
import os

import numpy as np

from ._host import ctx_for, to_device, like_input, is_device

_LEGACY_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libauv-color-balance.so")


def balance(mat, equalize_rgb=True, rgb_contrast_correct=False,
            hsv_contrast_correct=True, hsi_contrast_correct=False,
            rgb_extrema_clipping=True, adaptive_cast_correction=False,
            horizontal_blocks=1, vertical_blocks=1):
    """Same signature and defaults as the reference.  Accepts np.uint8[H,W,3] (returns a new numpy
    array, as the reference does) or a CUDA uint8 tensor [H,W,3] / [B,H,W,3] (returns a tensor)."""
    ctx = ctx_for(mat)
    out = ctx.color_balance(to_device(ctx, mat), equalize_rgb=equalize_rgb,
                            rgb_contrast_correct=rgb_contrast_correct,
                            hsv_contrast_correct=hsv_contrast_correct,
                            hsi_contrast_correct=hsi_contrast_correct,
                            rgb_extrema_clipping=rgb_extrema_clipping,
                            adaptive_cast_correction=adaptive_cast_correction,
                            horizontal_blocks=horizontal_blocks, vertical_blocks=vertical_blocks)
    return like_input(ctx, mat, out)


def balance_legacy(mat, equalize_rgb=True, rgb_contrast_correct=False,
                   hsv_contrast_correct=True, hsi_contrast_correct=False,
                   rgb_extrema_clipping=True, adaptive_cast_correction=False,
                   horizontal_blocks=1, vertical_blocks=1):
    """The reference's balance() body verbatim in behaviour (flattened copy, c_int8 pointer, no
    argtypes), bound to the CUDA-backed libauv-color-balance.so: proves the legacy `process_frame`
    symbol is a drop-in for modules/color_balance.py:93-110."""
    lib = ctypes.CDLL(_LEGACY_LIB)
    rows, cols, depth = mat.shape[0], mat.shape[1], 3
    c_int8_p = ctypes.POINTER(ctypes.c_int8)
    data = mat.flatten()
    data_p = data.ctypes.data_as(c_int8_p)
    rc = lib.process_frame(data_p, rows, cols, depth, equalize_rgb, rgb_contrast_correct,
                           hsv_contrast_correct, hsi_contrast_correct, rgb_extrema_clipping,
                           adaptive_cast_correction, horizontal_blocks, vertical_blocks)
    if rc != 0:
        lib.bv_last_error.restype = ctypes.c_char_p
        raise RuntimeError("process_frame failed (%d): %s" % (rc, lib.bv_last_error().decode()))
    return np.ctypeslib.as_array(data_p, (rows, cols, depth)).astype(np.uint8)
