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


def get_optimal_new_camera_matrix(camera_matrix, dist_coeffs, size, alpha, new_size=None):
    """cv2.getOptimalNewCameraMatrix(camera_matrix, dist_coeffs, size, alpha, new_size)[0] without OpenCV (host-side set-up
    arithmetic, float64, identical to cv2's result): a 9 x 9 grid over the image is undistorted into normalised coordinates
    (five fixed-point iterations of the inverse distortion model), the inscribed and the circumscribed rectangle of the
    grid are each mapped onto the viewport, and `alpha` blends the two projections (0: only valid pixels, 1: all
    source pixels kept).  size / new_size = (width, height)."""
    K = np.asarray(camera_matrix, np.float64)
    k = np.zeros(12)
    d = np.asarray(dist_coeffs if dist_coeffs is not None else [], np.float64).ravel()
    k[:d.size] = d
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    n = 9
    w, h = size
    pts = np.empty((n, n, 2))
    for gy in range(n):
        for gx in range(n):
            x0 = x = (gx * (w - 1) / (n - 1) - cx) * ifx
            y0 = y = (gy * (h - 1) / (n - 1) - cy) * ify
            for _ in range(5):
                r2 = x * x + y * y
                icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
                if icdist < 0:
                    x, y = x0, y0
                    break
                dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
                dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
                x, y = (x0 - dx) * icdist, (y0 - dy) * icdist
            pts[gy, gx] = (x, y)
    inner = (pts[:, 0, 0].max(), pts[0, :, 1].max(), pts[:, n - 1, 0].min(), pts[n - 1, :, 1].min())      # x0, y0, x1, y1
    outer = (pts[..., 0].min(), pts[..., 1].min(), pts[..., 0].max(), pts[..., 1].max())
    nw, nh = new_size if new_size else size
    out = K.copy()
    proj = []
    for x0, y0, x1, y1 in (inner, outer):
        fxr, fyr = (nw - 1) / (x1 - x0), (nh - 1) / (y1 - y0)
        proj.append((fxr, fyr, -fxr * x0, -fyr * y0))
    (fx0, fy0, cx0, cy0), (fx1, fy1, cx1, cy1) = proj
    out[0, 0] = fx0 * (1 - alpha) + fx1 * alpha
    out[1, 1] = fy0 * (1 - alpha) + fy1 * alpha
    out[0, 2] = cx0 * (1 - alpha) + cx1 * alpha
    out[1, 2] = cy0 * (1 - alpha) + cy1 * alpha
    return out


def init_undistort_map(params, width, height, alpha=0.0, fixed=True, like=None):
    """What include/camera_filters.hpp:11 declares (`initUndistortMap(&maps, name, width, height)`; the reference has no
    definition): read the camera file (path, YAML text or an (M, D) pair), take the optimal new camera matrix for
    `alpha`, build the two undistortion maps on the device.  Returns (map1, map2) for transform.remap."""
    m, d = load_camera_matrix_params(params) if isinstance(params, str) else params
    new_m = get_optimal_new_camera_matrix(m, d, (width, height), alpha)
    return init_undistort_rectify_map(m, d, None, new_m, (width, height), fixed=fixed, like=like)


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
