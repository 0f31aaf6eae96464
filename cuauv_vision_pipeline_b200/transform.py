"""GPU mirror of the reference's utils/transform.py (kernels, morphology, resize)."""
import numpy as np

from ._host import ctx_for, to_device, like_input


def rect_kernel(x, y=None):
    """utils/transform.py:54-77 (== cv2.getStructuringElement(MORPH_RECT, (x, y)))."""
    if y is None:
        y = x
    if x <= 0 or y <= 0:
        raise ValueError("x and y must be positive integers")
    return np.ones((y, x), np.uint8)


def elliptic_kernel(x, y=None):
    """utils/transform.py:27-51.  Restates cv2.getStructuringElement(MORPH_ELLIPSE): row i spans
    |dx| <= round(c * sqrt(r^2 - dy^2) / r) around the centre (OpenCV morph.cpp)."""
    if y is None:
        y = x
    if x % 2 == 0 or y % 2 == 0 or x <= 0 or y <= 0:
        raise ValueError("x and y must be odd positive integers")
    k = np.zeros((y, x), np.uint8)
    r, c = y // 2, x // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    for i in range(y):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            j1, j2 = max(c - dx, 0), min(c + dx + 1, x)
            k[i, j1:j2] = 1
    return k


def _morph(op):
    def fn(mat, kernel, iterations=1):
        ctx = ctx_for(mat)
        return like_input(ctx, mat, ctx.morph(to_device(ctx, mat), op, kernel, iterations))
    return fn


erode = _morph("erode")                        # utils/transform.py:80-94
dilate = _morph("dilate")                      # utils/transform.py:97-112
morph_remove_noise = _morph("open")            # utils/transform.py:115-129
morph_close_holes = _morph("close")            # utils/transform.py:132-146
morph_borders = _morph("gradient")             # utils/transform.py:149-164


def resize(mat, width, height):
    """utils/transform.py:167-179 == cv2.resize(mat, (width, height)) (INTER_LINEAR)."""
    ctx = ctx_for(mat)
    return like_input(ctx, mat, ctx.resize(to_device(ctx, mat), int(width), int(height)))
