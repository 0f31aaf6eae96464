"""Blob extraction: the role of the reference's outer_contours / contour_centroid / contour_area
(utils/feature.py:5-21,240-265) as connected-component labelling with exact raster moments."""
import numpy as np

from ._host import ctx_for, to_device, is_device


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
    if labels is not None and not is_device(mat):
        labels = ctx.download(labels)
    return labels, out, int(n[0])


def blob_centroid(blob):
    """Same rounding rule as contour_centroid (utils/feature.py:250-252): int(m10/m00), int(m01/m00)."""
    m00 = max(1e-10, float(blob["m00"]))
    return int(blob["m10"] / m00), int(blob["m01"] / m00)


def blob_area(blob):
    """Raster area in pixels (contour_area, utils/feature.py:255-265, is the polygon area)."""
    return float(blob["m00"])
