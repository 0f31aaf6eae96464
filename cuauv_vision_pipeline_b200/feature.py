"""Blob extraction: the role of the reference's outer_contours / contour_centroid / contour_area
(utils/feature.py:5-21,240-265) as connected-component labelling with exact raster moments."""
import numpy as np

from ._host import ctx_for, to_device, is_device, release_to_caller


class Blob(dict):
    """One labelled component: label, integer raster moments m00..m03 and bbox x0,y0,x1,y1."""

    @property
    def area(self):
        return int(self["m00"])


def label_blobs(mat, max_blobs=4096, want_labels=True):
    """8-connected labelling of `mat != 0`.  Returns (labels int32[H,W] or None, [Blob...]) with
    labels 1..n in raster order of each blob's first pixel."""
    ctx = ctx_for(mat)
    labels, blobs, nb = ctx.label(to_device(ctx, mat), max_blobs=max_blobs, want_labels=want_labels)
    n, tables = ctx.blobs_to_numpy(blobs, nb)
    out = []
    for i, row in enumerate(tables[0]):
        b = Blob({k: int(row[k]) for k in row.dtype.names})
        b["label"] = i + 1
        out.append(b)
    if labels is not None:
        labels = release_to_caller(ctx, labels) if is_device(mat) else ctx.download(labels)
    return labels, out, int(n[0])


class Contour(dict):
    """One outer border (cv2.findContours RETR_EXTERNAL) reduced to its Green's-theorem sums."""


def outer_contours(mat, max_contours=4096, points=False, max_points=None, rects=False):
    """utils/feature.py:5-21 on the GPU: the outermost 8-connected borders of `mat != 0`, in raster
    order of their first pixel (cv2 returns the same set, in its own order).  Each record carries
    what the reference consumes next (contour_centroid / contour_area below); with `points=True`
    also the CHAIN_APPROX_SIMPLE vertices as an int32 [n,1,2] array, exactly the array cv2 returns
    (ready for cv2.minAreaRect as in modules/bins.py:60); with `rects=True` also
    `c["min_area_rect"] = ((cx, cy), (w, h), angle)`, cv2.minAreaRect of those vertices computed on
    the device (min_area_rect below)."""
    from .runtime import CONTOUR_DTYPE
    ctx = ctx_for(mat)
    h, w = mat.shape[-2], mat.shape[-1]
    points = points or rects
    if points and max_points is None:
        max_points = 4 * (h + w) + h * w // 4        # generous: borders of a very ragged mask
    table, nb, pts, npts = ctx.outer_contours(to_device(ctx, mat), max_contours=max_contours,
                                              max_points=max_points if points else 0)
    n = int(ctx.download(nb)[0])
    raw = ctx.download(table)[0, :min(n, max_contours)].copy().view(CONTOUR_DTYPE).reshape(-1)
    keep = [i for i, row in enumerate(raw) if row["external"]]
    out = [Contour({k: int(raw[i][k]) for k in CONTOUR_DTYPE.names}) for i in keep]
    if rects:
        rr = ctx.min_area_rects(table, nb, pts)[0]
        for c, i in zip(out, keep):
            q = rr[i]
            c["min_area_rect"] = ((float(q["cx"]), float(q["cy"])), (float(q["width"]), float(q["height"])),
                                  float(q["angle"])) if q["valid"] else None
    if points:
        needed = int(ctx.download(npts)[0])
        host = ctx.download(pts)[0, :min(needed, max_points)]
        for c in out:
            off = c["point_offset"]
            c["points"] = host[off:off + c["n_simple"]].reshape(-1, 1, 2).copy() if off >= 0 else None
    return out


def _contour_moments(c):
    """m00, m10, m01 exactly as cv2.moments(contour) scales its sums (contourMoments: multiply by
    +-0.5 and +-1/6 in double)."""
    a00, a10, a01 = float(c["a00"]), float(c["a10"]), float(c["a01"])
    if abs(a00) <= 1.1920928955078125e-07:      # FLT_EPSILON: degenerate border, all moments 0
        return 0.0, 0.0, 0.0
    db1_2, db1_6 = (0.5, 0.16666666666666666666666666666667) if a00 > 0 else (-0.5, -0.16666666666666666666666666666667)
    return a00 * db1_2, a10 * db1_6, a01 * db1_6


def contour_centroid(contour):
    """utils/feature.py:240-252."""
    m00, m10, m01 = _contour_moments(contour)
    m00 = max(1e-10, m00)
    return int(m10 / m00), int(m01 / m00)


def contour_area(contour):
    """utils/feature.py:255-265 (cv2.contourArea, oriented=False)."""
    return abs(float(contour["a00"]) * 0.5)


def blob_centroid(blob):
    """Same rounding rule as contour_centroid (utils/feature.py:250-252): int(m10/m00), int(m01/m00)."""
    m00 = max(1e-10, float(blob["m00"]))
    return int(blob["m10"] / m00), int(blob["m01"] / m00)


def blob_area(blob):
    """Raster area in pixels (contour_area, utils/feature.py:255-265, is the polygon area)."""
    return float(blob["m00"])


def min_area_rect(contour):
    """cv2.minAreaRect(contour) (modules/bins.py:60) for a contour returned by
    `outer_contours(..., rects=True)`: ((cx, cy), (w, h), angle), angle in [-90, 0) degrees."""
    r = contour.get("min_area_rect")
    if r is None:
        raise ValueError("contour carries no rectangle: call outer_contours(..., rects=True)")
    return r


def min_enclosing_rect(contour):
    """utils/feature.py:301-312."""
    return min_area_rect(contour)
