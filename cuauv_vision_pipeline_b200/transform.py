"""GPU mirror of the reference's utils/transform.py (kernels, morphology, resize) and of the smoothing /
warping steps of modules/preprocessor.py:110-149."""
import math

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


def gaussian_blur(mat, ksize, sigma_x=0.0, sigma_y=0.0):
    """cv2.GaussianBlur(mat, ksize, sigma_x, sigma_y) on uint8 (modules/preprocessor.py:110-114)."""
    ctx = ctx_for(mat)
    if isinstance(ksize, int):
        ksize = (ksize, ksize)
    return like_input(ctx, mat, ctx.gaussian_blur(to_device(ctx, mat), ksize, sigma_x, sigma_y))


def rotation_matrix_2d(center, angle, scale):
    """cv2.getRotationMatrix2D (float64)."""
    a = angle * (math.pi / 180.0)
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    cx, cy = float(center[0]), float(center[1])
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def warp_affine(mat, matrix, dsize=None, border="constant", border_value=(0, 0, 0)):
    """cv2.warpAffine(mat, matrix, dsize, flags=INTER_LINEAR, borderMode, borderValue) on uint8."""
    ctx = ctx_for(mat)
    return like_input(ctx, mat, ctx.warp_affine(to_device(ctx, mat), matrix, dsize, border, border_value))


def rotate(mat, angle):
    """modules/preprocessor.py:130-135: rotation about the image centre, BORDER_REPLICATE."""
    h, w = mat.shape[0], mat.shape[1]
    return warp_affine(mat, rotation_matrix_2d((w / 2, h / 2), angle, 1), (w, h), border="replicate")


def translate(mat, tx, ty):
    """modules/preprocessor.py:144-149: float32 translation matrix, default border (constant 0)."""
    m = np.float32([[1, 0, tx], [0, 1, ty]])
    return warp_affine(mat, m, (mat.shape[1], mat.shape[0]))
