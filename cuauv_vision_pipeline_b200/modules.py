"""Drop-in vision modules: GPU versions of the reference's modules/bins.py::BinDetector,
modules/red_buoy.py::BuoyLAB and modules/color_balance.py::ColorBalance.

They keep the reference contract `process(self, direction, image)` with `image` an
np.uint8[H,W,3] BGR frame (core/base.py:936-942, dispatch 800/807) and publish through
`self.post(name, image)`.  `ModuleBase` is resolved at import time: the reference's own class when
`vision.core.base` is importable (the real CUAUV tree), else the duck-typed stand-in below, because
core/base.py does not import on Python 3.12 (dataclass default at core/base.py:521) and needs the
external `shm` / `auvlog` packages.
"""
import numpy as np

from . import feature
from .runtime import default_context

try:  # pragma: no cover - only inside the real CUAUV software tree
    from vision.core.base import ModuleBase  # type: ignore
except Exception:  # noqa: BLE001
    class ModuleBase:
        """Minimal stand-in with the reference's constructor / post / normalize surface
        (core/base.py:577-667, 846-891)."""

        def __init__(self, video_sources=None, tuners=None, fps=10):
            self.video_sources = list(video_sources or [])
            self.tuners = {t.name: t for t in (tuners or [])} if not isinstance(tuners, dict) else tuners
            self.fps = fps
            self.posted = {}
            self._shape = None

        def post(self, name, image, color_space="BGR"):
            self.posted[name] = np.asarray(image)

        def normalize(self, coord):
            """core/base.py:882-891 via 553-574: ((y - H/2)/W, (x - W/2)/W)."""
            h, w = self._shape
            y, x = coord
            return (y - h / 2) / w, (x - w / 2) / w

        def process(self, direction, image):
            raise NotImplementedError


class BinDetectorGPU(ModuleBase):
    """modules/bins.py:10-81 with the pixel work (13-27) as one fused stage call:
    BGR2HSV -> inRange([10,20,60],[30,100,255]) -> OPEN 5x5 -> blobs (+ the overlay of 19-20 when
    `overlay=True`).  Rectangle filtering (60-69) runs on the blob bounding boxes."""

    lower_beige = (10, 20, 60)
    upper_beige = (30, 100, 255)

    def __init__(self, *a, device=0, overlay=False, want_contours=True, **kw):
        super().__init__(*a, **kw)
        self.ctx = default_context(device)
        self.overlay = overlay
        self.want_contours = want_contours
        self.contours = []
        self.valid_rects = []
        self.desc = self.ctx.make_stage(cvt="bgr2hsv", lo=self.lower_beige, hi=self.upper_beige,
                                        morph=[("open", 5, 5, 1)], label=True)
        self.blobs = []

    def process(self, direction, img):
        self._shape = img.shape[:2]
        out = self.ctx.stage_host(self.desc, img, want=("mask", "blobs"), max_blobs=1024)
        n = int(out["n_blobs"][0])
        table = out["blobs"][0][:min(n, 1024)]
        self.blobs = []
        for i, row in enumerate(table):
            w = int(row["x1"] - row["x0"] + 1)
            h = int(row["y1"] - row["y0"] + 1)
            if w * h < 500:                                   # bins.py:64
                continue
            aspect = max(w, h) / min(w, h)
            if 1.0 <= aspect <= 3.0:                          # bins.py:67-68
                b = feature.Blob({k: int(row[k]) for k in row.dtype.names})
                b["label"] = i + 1
                self.blobs.append(b)
        cleaned = out["mask"]
        # the reference's own next step (bins.py:27): outer contours of the cleaned mask, as the exact
        # vertex arrays cv2.findContours returns -- ready for cv2.minAreaRect (bins.py:60) on the host
        self.contours = feature.outer_contours(cleaned, points=True, rects=True) if self.want_contours else []
        # bins.py:58-69 verbatim on the device-computed rectangles
        self.valid_rects = []
        for contour in self.contours:
            rect = contour["min_area_rect"]
            if rect is None:
                continue
            (center, (w, h), angle) = rect
            if w * h < 500:
                continue
            aspect_ratio = max(w, h) / min(w, h)
            if 1.0 <= aspect_ratio <= 3.0:
                self.valid_rects.append(rect)
        if self.overlay:
            vis = np.repeat(cleaned[..., None], 3, axis=2)
            overlayed = np.clip(np.rint(img * 0.7 + vis * 0.3), 0, 255).astype(np.uint8)   # bins.py:19-20
            self.post("bins", overlayed)
        else:
            self.post("bins", cleaned, "GRAY")
        return self.blobs


class BuoyLABGPU(ModuleBase):
    """modules/red_buoy.py:15-53: LAB a-channel inRange -> OPEN 5x5 -> CLOSE 5x5, blob centroid and
    area.  The reference takes contours of the un-cleaned mask (line 38) and leaves the choice of
    contour open (line 40, "logic omitted"); here the largest blob of the cleaned mask is reported."""

    def __init__(self, *a, device=0, thresh_min=150, thresh_max=255, **kw):
        super().__init__(*a, **kw)
        self.ctx = default_context(device)
        self.thresh = (thresh_min, thresh_max)
        self.result = None
        self.contour_result = None

    def process(self, direction, image):
        self._shape = image.shape[:2]
        lo, hi = self.thresh
        if "thresh_min" in getattr(self, "tuners", {}):
            lo, hi = self.tuners["thresh_min"].value, self.tuners["thresh_max"].value
        threshed_desc = self.ctx.make_stage(cvt="bgr2lab", lo=(0, lo, 0), hi=(255, hi, 255))
        cleaned_desc = self.ctx.make_stage(cvt="bgr2lab", lo=(0, lo, 0), hi=(255, hi, 255),
                                           morph=[("open", 5, 5, 1), ("close", 5, 5, 1)], label=True)
        threshed = self.ctx.stage_host(threshed_desc, image, want=("mask",))["mask"]
        self.post("threshed", threshed, "GRAY")
        out = self.ctx.stage_host(cleaned_desc, image, want=("mask", "blobs"), max_blobs=1024)
        self.post("threshed_cleaned", out["mask"], "GRAY")
        n = min(int(out["n_blobs"][0]), 1024)
        self.result = None
        if n:
            table = out["blobs"][0][:n]
            best = table[int(np.argmax(table["m00"]))]
            x, y = feature.blob_centroid(best)
            ny, nx = self.normalize((y, x))
            self.result = dict(center_x=nx, center_y=ny, area=float(best["m00"]), pixel=(x, y))
        # the literal reference path (red_buoy.py:38-44): outer contours of the UN-cleaned mask, polygon
        # centroid and area; "most likely contour" (line 40, logic omitted upstream) = largest area here
        contours = feature.outer_contours(threshed)
        if contours:
            c = max(contours, key=feature.contour_area)
            cx, cy = feature.contour_centroid(c)
            cny, cnx = self.normalize((cy, cx))
            self.contour_result = dict(center_x=cnx, center_y=cny, area=feature.contour_area(c), pixel=(cx, cy))
        else:
            self.contour_result = None
        return self.result


class ColorBalanceGPU(ModuleBase):
    """modules/color_balance.py:112-121: post the original and the balanced frame."""

    def __init__(self, *a, device=0, **kw):
        super().__init__(*a, **kw)
        self.ctx = default_context(device)
        self.desc = self.ctx.make_stage(balance={})

    def process(self, direction, mat):
        self._shape = mat.shape[:2]
        self.post("orig", mat)
        balanced = self.ctx.stage_host(self.desc, mat, want=("balanced",))["balanced"]
        self.post("balanced", balanced)
        return balanced
