"""TEST INFRASTRUCTURE ONLY -- the literal cv2 / numpy calls the reference makes on the hot path.

Each function restates one reference function (file:line under /root/reference) by making the
same OpenCV call.  OpenCV is a third-party dependency of the reference (configure.py:27-33); the
installed cv2 4.13.0 is the executable oracle.
"""
import numpy as np
import cv2

# utils/color.py:26-32 ---------------------------------------------------------------------------
_CODES = {
    "bgr2lab": cv2.COLOR_BGR2LAB, "bgr2hsv": cv2.COLOR_BGR2HSV, "bgr2hls": cv2.COLOR_BGR2HLS,
    "bgr2ycrcb": cv2.COLOR_BGR2YCrCb, "bgr2luv": cv2.COLOR_BGR2LUV, "bgr2gray": cv2.COLOR_BGR2GRAY,
    "gray2bgr": cv2.COLOR_GRAY2BGR, "lab2bgr": cv2.COLOR_LAB2BGR, "hsv2bgr": cv2.COLOR_HSV2BGR,
}


def convert(mat, name):
    """utils/color.py:11-23 `_convert_colorspace`: returns (converted, split planes)."""
    conv = cv2.cvtColor(mat, _CODES[name])
    return conv, cv2.split(conv)


def range_threshold(mat, lo, hi):
    """utils/color.py:105-121; modules/bins.py:16 (3-channel bounds as arrays)."""
    return cv2.inRange(mat, lo, hi)


def thresh_color_distance(split, color, distance, auto_distance_percentile=None, ignore_channels=(), weights=(1, 1, 1)):
    """utils/color.py:66-103: weighted squared distance accumulated in a float32 image (numpy evaluates
    weight * square in float64 and rounds on the in-place add), thresholded with
    cv2.inRange(dists, 0, limit), limit = distance**2 or min(np.percentile(dists, p), distance**2) (98-101);
    second result np.uint8(np.sqrt(dists))."""
    w = list(weights)
    for idx in ignore_channels:
        w[idx] = 0
    w = np.asarray(w, dtype=np.float64) / np.linalg.norm(weights)
    dists = np.zeros(split[0].shape, dtype=np.float32)
    for i in range(3):
        if i in ignore_channels:
            continue
        dists += w[i] * (np.float32(split[i]) - color[i]) ** 2
    if auto_distance_percentile:
        limit = min(np.percentile(dists, auto_distance_percentile), distance ** 2)
    else:
        limit = distance ** 2
    with np.errstate(invalid="ignore"):
        return cv2.inRange(dists, 0, float(limit)), (np.sqrt(dists).astype(np.int64) & 0xFF).astype(np.uint8)


def binary_threshold(mat, t):            # utils/color.py:124-137
    return cv2.threshold(mat, t, 255, cv2.THRESH_BINARY)[1]


def binary_threshold_inv(mat, t):        # utils/color.py:140-153
    return cv2.threshold(mat, t, 255, cv2.THRESH_BINARY_INV)[1]


def max_threshold(mat, t):               # utils/color.py:156-169
    return cv2.threshold(mat, t, 0, cv2.THRESH_TRUNC)[1]


def above_threshold(mat, t):             # utils/color.py:172-185
    return cv2.threshold(mat, t, 0, cv2.THRESH_TOZERO)[1]


def below_threshold(mat, t):             # utils/color.py:188-201
    return cv2.threshold(mat, t, 0, cv2.THRESH_TOZERO_INV)[1]


# utils/transform.py ------------------------------------------------------------------------------
def rect_kernel(x, y=None):              # utils/transform.py:54-77
    return cv2.getStructuringElement(cv2.MORPH_RECT, (x, x if y is None else y))


def elliptic_kernel(x, y=None):          # utils/transform.py:27-51
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (x, x if y is None else y))


def erode(mat, kernel, iterations=1):    # utils/transform.py:80-94
    return cv2.erode(mat, kernel, iterations=iterations)


def dilate(mat, kernel, iterations=1):   # utils/transform.py:97-112
    return cv2.dilate(mat, kernel, iterations=iterations)


def morph_remove_noise(mat, kernel, iterations=1):   # utils/transform.py:115-129
    return cv2.morphologyEx(mat, cv2.MORPH_OPEN, kernel, iterations=iterations)


def morph_close_holes(mat, kernel, iterations=1):    # utils/transform.py:132-146
    return cv2.morphologyEx(mat, cv2.MORPH_CLOSE, kernel, iterations=iterations)


def morph_borders(mat, kernel, iterations=1):        # utils/transform.py:149-164
    return cv2.morphologyEx(mat, cv2.MORPH_GRADIENT, kernel, iterations=iterations)


def resize(mat, width, height):          # utils/transform.py:167-179
    return cv2.resize(mat, (width, height))


# utils/feature.py --------------------------------------------------------------------------------
def outer_contours(mat):                 # utils/feature.py:5-21
    contours, _ = cv2.findContours(mat, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    return contours


def contour_centroid(contour):           # utils/feature.py:240-252
    moments = cv2.moments(contour)
    m00 = max(1e-10, moments["m00"])
    return int(moments["m10"] / m00), int(moments["m01"] / m00)


def contour_area(contour):               # utils/feature.py:255-265
    return cv2.contourArea(contour, oriented=False)


# modules/preprocessor.py -------------------------------------------------------------------------
def channel_bias(mat, channel, bias):
    """preprocessor.py:89-103: split, cv2.add(scalar, plane) (saturating), merge."""
    planes = list(cv2.split(mat))
    planes[channel] = cv2.add(bias, planes[channel])
    # cv2.add(scalar, plane) returns the plane shape in cv2 >= 4.5; keep it 2-D for merge
    planes[channel] = np.asarray(planes[channel]).reshape(mat.shape[:2]).astype(np.uint8)
    return cv2.merge(planes)


def contrast(mat, c):
    """preprocessor.py:104-106: float64 multiply, clip, truncating cast."""
    return np.clip(mat * c, 0., 255.).astype(np.uint8)


def brightness(mat, b):
    """preprocessor.py:107-109."""
    return np.clip(mat + float(b), 0., 255.).astype(np.uint8)


def ellipse_erode(mat, k):
    """preprocessor.py:120-124 with PPX_erode_kernel = k."""
    return cv2.erode(mat, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * k + 1, 2 * k + 1)))


def ellipse_dilate(mat, k):
    """preprocessor.py:125-129."""
    return cv2.dilate(mat, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * k + 1, 2 * k + 1)))


def resize_ratio(mat, ratio):
    """preprocessor.py:140-143: truncating int() on both dimensions."""
    return cv2.resize(mat, (int(mat.shape[1] * ratio), int(mat.shape[0] * ratio)))


# modules/bins.py / modules/red_buoy.py pipelines ------------------------------------------------
def bins_mask(img, lo=(10, 20, 60), hi=(30, 100, 255)):
    """modules/bins.py:13-24: BGR2HSV -> inRange -> OPEN 5x5.  Returns (mask, cleaned)."""
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mask = cv2.inRange(hsv, np.array(lo), np.array(hi))
    cleaned = morph_remove_noise(mask, rect_kernel(5))
    return mask, cleaned


def buoy_mask(img, lo, hi):
    """modules/red_buoy.py:21-34: LAB a-channel -> inRange -> OPEN 5x5 -> CLOSE 5x5."""
    _, (_, lab_a, _) = convert(img, "bgr2lab")
    threshed = cv2.inRange(lab_a, lo, hi)
    cleaned = morph_close_holes(morph_remove_noise(threshed, rect_kernel(5)), rect_kernel(5))
    return threshed, cleaned
