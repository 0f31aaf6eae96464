"""GPU mirror of the reference's utils/color.py (same function names, arguments and return
shapes).  Citations are to /root/reference/utils/color.py."""
import numpy as np

from ._host import ctx_for, to_device, like_input


def _convert_colorspace(code):
    """utils/color.py:11-23: returns (converted image, list of split channels)."""
    def _inner(mat):
        ctx = ctx_for(mat)
        src = to_device(ctx, mat)
        if code == "bgr2gray":
            conv = ctx.cvt_color(src, code)
            out = like_input(ctx, mat, conv)
            return out, [out]                      # cv2.split of a 1-channel image
        if code == "gray2bgr":
            conv = ctx.cvt_color(src, code)
            out = like_input(ctx, mat, conv)
            return out, [out[..., k] for k in range(3)]
        conv, planes = ctx.cvt_color(src, code, split=True)
        return like_input(ctx, mat, conv), [like_input(ctx, mat, p) for p in planes]
    return _inner


# utils/color.py:26-32.  bgr_to_luv and lab_to_bgr have no pinned arithmetic model yet
#  and are not provided.
bgr_to_lab = _convert_colorspace("bgr2lab")
bgr_to_hsv = _convert_colorspace("bgr2hsv")
bgr_to_hls = _convert_colorspace("bgr2hls")
bgr_to_ycrcb = _convert_colorspace("bgr2ycrcb")
bgr_to_gray = _convert_colorspace("bgr2gray")
gray_to_bgr = _convert_colorspace("gray2bgr")
hsv_to_bgr = _convert_colorspace("hsv2bgr")


def range_threshold(mat, min, max):  # noqa: A002 (reference argument names)
    """utils/color.py:105-121 == cv2.inRange(mat, min, max); scalars or per-channel sequences."""
    ctx = ctx_for(mat)
    return like_input(ctx, mat, ctx.in_range(to_device(ctx, mat), min, max))


def _thresh(kind, maxval=255):
    def fn(mat, threshold):
        ctx = ctx_for(mat)
        mv = maxval if maxval is not None else 0
        return like_input(ctx, mat, ctx.threshold(to_device(ctx, mat), threshold, mv, kind))
    return fn


binary_threshold = _thresh("binary")            # utils/color.py:124-137
binary_threshold_inv = _thresh("binary_inv")    # utils/color.py:140-153
max_threshold = _thresh("trunc", 0)             # utils/color.py:156-169
above_threshold = _thresh("tozero", 0)          # utils/color.py:172-185
below_threshold = _thresh("tozero_inv", 0)      # utils/color.py:188-201
