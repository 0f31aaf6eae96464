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


# ---- lens undistortion (include/camera_filters.hpp:6-11; lib/configs/*_camera_matrix_params.yaml) -------------------
def load_camera_matrix_params(path_or_text):
    """Reads an OpenCV FileStorage YAML as the reference ships them (lib/configs/1_camera_matrix_params.yaml: a 3x3
    camera matrix `M` and the distortion row `D`) without OpenCV.  Returns (M float64 [3,3], D float64 [n])."""
    import os
    import re
    text = path_or_text
    if "\n" not in path_or_text and os.path.exists(path_or_text):
        with open(path_or_text) as f:
            text = f.read()
    out = {}
    for name, rows, cols, data in re.findall(
            r"(\w+):\s*!!opencv-matrix\s*rows:\s*(\d+)\s*cols:\s*(\d+)\s*dt:\s*\w+\s*data:\s*\[([^\]]*)\]", text):
        out[name] = np.array([float(v) for v in data.replace("\n", " ").split(",") if v.strip()], np.float64).reshape(int(rows), int(cols))
    if "M" not in out or "D" not in out:
        raise ValueError("no camera matrix M / distortion D in the file")
    return out["M"], out["D"].ravel()


def init_undistort_rectify_map(camera_matrix, dist_coeffs, rotation, new_camera_matrix, size, fixed=False, like=None):
    """cv2.initUndistortRectifyMap(camera_matrix, dist_coeffs, rotation, new_camera_matrix, size, CV_32FC1) as device maps
    (fixed=True: the CV_16SC2 + CV_16UC1 pair cv2.undistort computes).  The 3x3 inverse is taken on the host."""
    ctx = ctx_for(like)
    new_k = np.asarray(camera_matrix if new_camera_matrix is None else new_camera_matrix, np.float64)[:3, :3]
    rot = np.eye(3) if rotation is None else np.asarray(rotation, np.float64)
    return ctx.undistort_maps(camera_matrix, dist_coeffs, np.linalg.inv(new_k @ rot), size, fixed=fixed)


def remap(mat, map1, map2, border="constant", border_value=(0, 0, 0)):
    """cv2.remap(mat, map1, map2, cv2.INTER_LINEAR, borderMode=...); maps as returned by init_undistort_rectify_map, or
    host arrays (float32 x / y planes, or int16 xy + uint16 fractions), which are uploaded."""
    import torch
    ctx = ctx_for(mat)

    def dev(m):
        return m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m)).to(ctx.device)
    return like_input(ctx, mat, ctx.remap(to_device(ctx, mat), dev(map1), dev(map2), border=border, border_value=border_value))


def undistort(mat, camera_matrix, dist_coeffs, new_camera_matrix=None, maps=None):
    """cv2.undistort(mat, camera_matrix, dist_coeffs, None, new_camera_matrix): fixed-point maps from the float64
    coordinates, bilinear remap, constant zero border.  Pass `maps` (from init_undistort_rectify_map(..., fixed=True))
    to reuse them across frames, which is what the reference's optimal_camera_matrix struct is for."""
    if maps is None:
        h, w = (mat.shape[0], mat.shape[1]) if len(mat.shape) <= 3 else (mat.shape[1], mat.shape[2])
        maps = init_undistort_rectify_map(camera_matrix, dist_coeffs, None, new_camera_matrix, (w, h), fixed=True, like=mat)
    return remap(mat, maps[0], maps[1])
