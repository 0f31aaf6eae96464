"""TEST INFRASTRUCTURE ONLY -- declared oracle for labelling + per-blob raster moments.

The reference never labels components: it extracts blobs with cv2.findContours(RETR_EXTERNAL)
(utils/feature.py:20) and takes polygon moments (utils/feature.py:250-252, 265).  SURVEY.md 8c
declares the oracle for the CUDA labelling path:

    cv2.connectedComponentsWithStats(mask, connectivity=8, ltype=CV_32S)

with labels canonicalised by ascending raster index of each component's first pixel (cv2's own
numbering is not raster order), plus exact integer raster moments per label, i.e. what
cv2.moments((labels == i), binaryImage=True) returns, recomputed here in int64/python-int so that
third-order moments beyond 2^53 stay exact.
"""
import numpy as np
import cv2

MOMENT_KEYS = ("m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03")


def canonical_labels(mask):
    """Returns (n_blobs, labels int32[H,W]) with labels 1..n in first-pixel raster order, 0 = bg."""
    binary = (np.asarray(mask) != 0).astype(np.uint8)
    n, lab = cv2.connectedComponents(binary, connectivity=8, ltype=cv2.CV_32S)
    if n <= 1:
        return 0, np.zeros(binary.shape, np.int32)
    flat = lab.reshape(-1)
    first = np.full(n, flat.size, np.int64)
    # first raster index of each cv2 label
    np.minimum.at(first, flat, np.arange(flat.size, dtype=np.int64))
    order = np.argsort(first[1:], kind="stable") + 1          # cv2 labels sorted by first pixel
    remap = np.zeros(n, np.int32)
    remap[order] = np.arange(1, n, dtype=np.int32)
    return n - 1, remap[lab]


def blob_table(labels, n):
    """Exact integer raster moments, bounding boxes and areas per canonical label.

    Returns dict of int64 arrays of length n (index i-1 for label i): the ten raster moments up to
    third order, and bbox x0,y0,x1,y1 (inclusive)."""
    h, w = labels.shape
    ys, xs = np.nonzero(labels)
    lab = labels[ys, xs].astype(np.int64) - 1
    xs = xs.astype(np.int64)
    ys = ys.astype(np.int64)
    out = {}

    def acc(vals):
        # int64 accumulation is exact: sum x^3 over a full 4K frame is ~1.2e17 < 2^63
        r = np.zeros(n, np.int64)
        np.add.at(r, lab, vals)
        return r
    out["m00"] = acc(np.ones_like(xs))
    out["m10"] = acc(xs)
    out["m01"] = acc(ys)
    out["m20"] = acc(xs * xs)
    out["m11"] = acc(xs * ys)
    out["m02"] = acc(ys * ys)
    out["m30"] = acc(xs * xs * xs)
    out["m21"] = acc(xs * xs * ys)
    out["m12"] = acc(xs * ys * ys)
    out["m03"] = acc(ys * ys * ys)
    x0 = np.full(n, w, np.int64)
    y0 = np.full(n, h, np.int64)
    x1 = np.full(n, -1, np.int64)
    y1 = np.full(n, -1, np.int64)
    np.minimum.at(x0, lab, xs)
    np.minimum.at(y0, lab, ys)
    np.maximum.at(x1, lab, xs)
    np.maximum.at(y1, lab, ys)
    out.update(x0=x0, y0=y0, x1=x1, y1=y1)
    return out


def label_and_moments(mask):
    n, lab = canonical_labels(mask)
    return n, lab, blob_table(lab, n)


def cv2_moments_of_label(labels, i):
    """The literal cv2 call of the declared oracle, used to pin blob_table on small cases."""
    return cv2.moments((labels == i).astype(np.uint8), binaryImage=True)


def reference_contour_path(mask):
    """The literal reference path (utils/feature.py:20,250-252,265): outer contours, polygon
    centroid and area.  Reported next to the raster numbers, not gated (SURVEY.md finding 4)."""
    contours, _ = cv2.findContours(np.ascontiguousarray(mask), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    res = []
    for c in contours:
        m = cv2.moments(c)
        m00 = max(1e-10, m["m00"])
        res.append(((int(m["m10"] / m00), int(m["m01"] / m00)), cv2.contourArea(c, oriented=False)))
    return res
