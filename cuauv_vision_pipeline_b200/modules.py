"""Drop-in vision modules: GPU versions of the reference's modules/bins.py::BinDetector,
modules/red_buoy.py::BuoyLAB and modules/color_balance.py::ColorBalance.

They keep the reference contract `process(self, direction, image)` with `image` an
np.uint8[H,W,3] BGR frame (core/base.py:936-942, dispatch 800/807) and publish through
`self.post(name, image, color_space)` exactly the images the reference modules post.

Two seams:

* `bind(ModuleBase)` builds the three classes on top of ANY base class with the reference's
  constructor / post / normalize surface.  Inside the CUAUV tree that is the reference's own
  `vision.core.base.ModuleBase` (tests/test_real_runtime.py drives them through its `__call__` /
  `_loop`, fed by the reference's image_directory capture source over the reference's transport);
  elsewhere it is the duck-typed stand-in below, because core/base.py does not import on
  Python 3.12 (dataclass default at core/base.py:521) and needs the external `shm` / `auvlog`.
* the per-frame pixel work sits in `GpuPixels`: ONE upload per frame, every further step on device
  handles (conversion, threshold, morphology, labels, contours, rectangles), small results and the
  posted images come back.  A module takes its pixel backend as the `pixels=` argument.
"""
import numpy as np

from . import feature
from .runtime import default_context


class StandInModuleBase:
    """Minimal stand-in with the reference's constructor / post / normalize surface
    (core/base.py:577-667, 846-891)."""

    def __init__(self, video_sources=None, tuners=None, fps=10):
        self.video_sources = list(video_sources or [])
        self.tuners = {t.name: t for t in (tuners or [])} if not isinstance(tuners, dict) else tuners
        self.fps = fps
        self.posted = {}
        self.posted_color_space = {}
        self._shape = None

    def post(self, name, image, color_space="BGR"):
        self.posted[name] = np.array(image, np.uint8, copy=True, order="C", ndmin=1)   # core/base.py:863
        self.posted_color_space[name] = color_space.upper()

    def normalize(self, coord):
        """core/base.py:882-891 via 553-574: ((y - H/2)/W, (x - W/2)/W)."""
        h, w = self._shape
        y, x = coord
        return (y - h / 2) / w, (x - w / 2) / w

    def process(self, direction, image):
        raise NotImplementedError


class GpuPixels:
    """The device side of the three modules.  Every method uploads the frame once and keeps all
    intermediates on the device; `h2d_bytes` / `d2h_bytes` count what crossed PCIe (tests assert
    one upload per process() call)."""

    def __init__(self, device=0):
        self.ctx = default_context(device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.uploads = 0
        self._rect5 = np.ones((5, 5), np.uint8)                     # utils/transform.py:54-77 rect_kernel(5)
        self._balance_desc = self.ctx.make_stage(balance={})

    def _up(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        self.h2d_bytes += img.nbytes
        self.uploads += 1
        return self.ctx.upload(img)

    def _down(self, t):
        a = self.ctx.download(t)
        self.d2h_bytes += a.nbytes
        return a

    def _contours(self, mask_dev, rects):
        """Outer contours of a DEVICE mask (no re-upload): vertex lists, polygon centroid / area, optionally minAreaRect."""
        cs = feature.outer_contours(mask_dev, points=True, rects=rects)
        for c in cs:
            c["centroid"] = feature.contour_centroid(c)
            c["area"] = feature.contour_area(c)
            self.d2h_bytes += 72 + (0 if c.get("points") is None else c["points"].nbytes)
        return cs

    def bins(self, img, lo, hi):
        """modules/bins.py:13-27,58-69: HSV inRange mask, OPEN 5x5, outer contours of the cleaned mask with
        their cv2.minAreaRect, and the labelled blobs.  Returns the raw mask (for the overlay of 19-20) on the host."""
        ctx = self.ctx
        d = self._up(img)
        raw = ctx.cvt_in_range(d, "bgr2hsv", lo, hi)
        cleaned = ctx.morph(raw, "open", self._rect5)
        contours = self._contours(cleaned, rects=True)
        _, blobs, nb = ctx.label(cleaned, max_blobs=1024, want_labels=False)
        n, tables = ctx.blobs_to_numpy(blobs, nb)
        self.d2h_bytes += tables[0].nbytes + 4
        return dict(mask=self._down(raw), cleaned_dev=cleaned, contours=contours, blobs=tables[0])

    def buoy(self, img, lo, hi):
        """modules/red_buoy.py:21-44: LAB once, a-channel inRange, OPEN + CLOSE, contours of the un-cleaned mask."""
        ctx = self.ctx
        d = self._up(img)
        _, (_, lab_a, _) = ctx.cvt_color(d, "bgr2lab", split=True)
        threshed = ctx.in_range(lab_a, lo, hi)
        cleaned = ctx.morph(ctx.morph(threshed, "open", self._rect5), "close", self._rect5)
        contours = self._contours(threshed, rects=False)
        _, blobs, nb = ctx.label(cleaned, max_blobs=1024, want_labels=False)
        n, tables = ctx.blobs_to_numpy(blobs, nb)
        self.d2h_bytes += tables[0].nbytes + 4
        return dict(threshed=self._down(threshed), cleaned=self._down(cleaned), contours=contours, blobs=tables[0])

    def balance(self, img):
        """modules/color_balance.py:93-110 balance() with the default flags."""
        d = self._up(img)
        out = self.ctx.stage(self._balance_desc, d, want=("balanced",))["balanced"]
        return self._down(out)


def _int0(a):
    """np.int0 of modules/bins.py:73 (an alias of np.intp that numpy 2 removed): truncation toward zero."""
    return np.asarray(a).astype(np.intp)


def _draw_box(img, rect):
    """cv2.drawContours(overlayed, [np.int0(cv2.boxPoints(rect))], 0, (0, 255, 0), 4) of modules/bins.py:72-74.
    Drawing is visualisation on the host image that gets posted (a few hundred pixels); cv2 is the reference's own
    dependency and is imported only here, when a module actually has a rectangle to draw."""
    import cv2
    cv2.drawContours(img, [_int0(cv2.boxPoints(rect))], 0, (0, 255, 0), 4)


def bind(ModuleBase):
    """The three GPU modules as subclasses of `ModuleBase`.  Returns (BinDetectorGPU, BuoyLABGPU, ColorBalanceGPU)."""

    class BinDetectorGPU(ModuleBase):
        """modules/bins.py:10-81.  Posts "bins": the frame overlaid with the RAW inRange mask (19-20) and the
        accepted rectangles drawn in green (71-74), as the reference does; `self.valid_rects` is its list of
        cv2.minAreaRect tuples (58-69), `self.blobs` the labelled components of the cleaned mask."""

        lower_beige = (10, 20, 60)
        upper_beige = (30, 100, 255)

        def __init__(self, *a, device=0, pixels=None, **kw):
            super().__init__(*a, **kw)
            self.pixels = pixels if pixels is not None else GpuPixels(device)
            self.contours = []
            self.valid_rects = []
            self.blobs = []

        def process(self, direction, img):
            self._shape = img.shape[:2]
            r = self.pixels.bins(img, self.lower_beige, self.upper_beige)
            mask = r["mask"]
            # bins.py:19-20: cv2.addWeighted(img, 0.7, GRAY2BGR(mask), 0.3, 0): float32 products and sum, each rounded on
            # its own, then round-half-even and saturate (identical to cv2 for every (a, b in {0, 255}) pair)
            overlayed = np.clip(np.rint(img.astype(np.float32) * np.float32(0.7) + mask[..., None].astype(np.float32) * np.float32(0.3)),
                                0, 255).astype(np.uint8)
            self.contours = r["contours"]
            self.valid_rects = []
            for contour in self.contours:                      # bins.py:58-69 verbatim on the device's rectangles
                rect = contour["min_area_rect"]
                if rect is None:
                    continue
                (center, (w, h), angle) = rect
                if w * h < 500:
                    continue
                aspect_ratio = max(w, h) / min(w, h)
                if 1.0 <= aspect_ratio <= 3.0:
                    self.valid_rects.append(rect)
            for rect in self.valid_rects:                      # bins.py:71-74
                _draw_box(overlayed, rect)
            self.blobs = []
            for i, row in enumerate(r["blobs"]):
                b = feature.Blob({k: int(row[k]) for k in row.dtype.names})
                b["label"] = i + 1
                self.blobs.append(b)
            self.post("bins", overlayed)                       # bins.py:81
            return self.valid_rects

    class BuoyLABGPU(ModuleBase):
        """modules/red_buoy.py:15-53: LAB a-channel inRange -> OPEN 5x5 -> CLOSE 5x5; posts "threshed" and
        "threshed_cleaned" (GRAY).  The reference takes contours of the un-cleaned mask (line 38) and leaves the
        choice of contour open (line 40, "logic omitted"): here the contour of largest area; `self.result` holds
        what it writes to shm (46-49) from that contour's polygon centroid / area, `self.blob_result` the same from
        the largest labelled blob of the cleaned mask."""

        def __init__(self, *a, device=0, pixels=None, thresh_min=150, thresh_max=255, **kw):
            super().__init__(*a, **kw)
            self.pixels = pixels if pixels is not None else GpuPixels(device)
            self.thresh = (thresh_min, thresh_max)
            self.result = None
            self.blob_result = None

        def _bounds(self):
            lo, hi = self.thresh
            try:                                               # red_buoy.py:25-29: self.tuners["thresh_min"]
                t = self.tuners
                lo, hi = t["thresh_min"], t["thresh_max"]
                lo, hi = getattr(lo, "value", lo), getattr(hi, "value", hi)
            except (KeyError, TypeError):
                pass
            return int(lo), int(hi)

        def process(self, direction, image):
            self._shape = image.shape[:2]
            lo, hi = self._bounds()
            r = self.pixels.buoy(image, lo, hi)
            self.post("threshed", r["threshed"], "GRAY")
            self.post("threshed_cleaned", r["cleaned"], "GRAY")
            self.result = None
            if r["contours"]:
                c = max(r["contours"], key=lambda q: q["area"])
                x, y = c["centroid"]
                ny, nx = self.normalize((y, x))
                self.result = dict(center_x=nx, center_y=ny, area=c["area"], pixel=(x, y))
            self.blob_result = None
            if len(r["blobs"]):
                best = r["blobs"][int(np.argmax(r["blobs"]["m00"]))]
                x, y = feature.blob_centroid(best)
                ny, nx = self.normalize((y, x))
                self.blob_result = dict(center_x=nx, center_y=ny, area=float(best["m00"]), pixel=(x, y))
            return self.result

    class ColorBalanceGPU(ModuleBase):
        """modules/color_balance.py:112-121: post the original and the balanced frame."""

        def __init__(self, *a, device=0, pixels=None, **kw):
            super().__init__(*a, **kw)
            self.pixels = pixels if pixels is not None else GpuPixels(device)

        def process(self, direction, mat):
            self._shape = mat.shape[:2]
            self.post("orig", mat)
            balanced = self.pixels.balance(mat)
            self.post("balanced", balanced)
            return balanced

    return BinDetectorGPU, BuoyLABGPU, ColorBalanceGPU


try:  # pragma: no cover - only inside the real CUAUV software tree
    from vision.core.base import ModuleBase  # type: ignore
except Exception:  # noqa: BLE001
    ModuleBase = StandInModuleBase

BinDetectorGPU, BuoyLABGPU, ColorBalanceGPU = bind(ModuleBase)
