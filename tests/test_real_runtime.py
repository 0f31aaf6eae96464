"""The drop-in modules under the reference's REAL runtime:

  reference capture source  capture_sources/image_directory.py:13-36 (PNG files on disk, cv2.imread)
        -> reference CaptureSource thread, core/capture_source.py:128-234
        -> reference cffi binding BlockAccessor, core/bindings/camera_message_framework.py:117-441
        -> reference transport compiled unmodified (oracle/_ref/libcamera_message_framework.so)
        -> reference ModuleBase.__call__ / _loop, core/base.py:642-844  (line 521 patched in memory)
        -> OUR modules (cuauv_vision_pipeline_b200.modules.bind(ModuleBase)) .process(direction, image)
        -> reference ModuleBase.post -> ModuleManager.post -> a "module_..._post" block, read back here.

Runs in the container that has /root/reference (it does not exist on the GPU box).  The container has no GPU, so the
modules' pixel backend is the cv2 chain of the reference call sites (`OraclePixels`, the same three methods as
modules.GpuPixels); with a CUDA device present the same test uses GpuPixels.  The GPU half of the claim -- GpuPixels
== that cv2 chain, one upload per call -- is tests/test_gpu_balance_stage.py::test_drop_in_modules and
tests/test_cmf_dropin.py on the GPU box.
"""
import os
import signal
import subprocess
import sys
import threading
import time

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refruntime  # noqa: E402
from oracle import cv_ops, ref_balance, color_balance_np, synth  # noqa: E402

pytestmark = pytest.mark.skipif(not refruntime.available(), reason="needs /root/reference and oracle/_ref (CPU container only)")


class OraclePixels:
    """modules.GpuPixels' interface on the reference's own cv2 calls."""

    def __init__(self):
        self.uploads = 0

    @staticmethod
    def _contours(mask, rects):
        out = []
        for c in cv_ops.outer_contours(mask):
            d = dict(points=c, centroid=cv_ops.contour_centroid(c), area=cv_ops.contour_area(c))
            if rects:
                d["min_area_rect"] = cv2.minAreaRect(c)
            out.append(d)
        return out

    @staticmethod
    def _blobs(mask):
        from oracle import ccl
        n, _, tab = ccl.label_and_moments(mask)
        rec = np.zeros(n, dtype=[(k, "<i8") for k in ccl.MOMENT_KEYS] + [(k, "<i4") for k in ("x0", "y0", "x1", "y1")])
        for k in rec.dtype.names:
            rec[k] = tab[k]
        return rec

    def bins(self, img, lo, hi):
        self.uploads += 1
        mask, cleaned = cv_ops.bins_mask(img, lo, hi)
        return dict(mask=mask, cleaned_dev=cleaned, contours=self._contours(cleaned, True), blobs=self._blobs(cleaned))

    def buoy(self, img, lo, hi):
        self.uploads += 1
        th, cl = cv_ops.buoy_mask(img, lo, hi)
        return dict(threshed=th, cleaned=cl, contours=self._contours(th, False), blobs=self._blobs(cl))

    def balance(self, img):
        self.uploads += 1
        return ref_balance.balance(img) if ref_balance.available() else color_balance_np.process_frame_np(img)


def make_pixels():
    import torch
    if torch.cuda.is_available():
        from cuauv_vision_pipeline_b200.modules import GpuPixels
        return GpuPixels(0)
    return OraclePixels()


def reference_bins_post(img):
    """modules/bins.py:11-81 literally (np.int0 spelled np.intp)."""
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mask = cv2.inRange(hsv, np.array([10, 20, 60]), np.array([30, 100, 255]))
    overlayed = cv2.addWeighted(img, 0.7, cv2.cvtColor(mask, cv2.COLOR_GRAY2BGR), 0.3, 0)
    cleaned = cv_ops.morph_remove_noise(mask, cv_ops.rect_kernel(5))
    for contour in cv_ops.outer_contours(cleaned):
        rect = cv2.minAreaRect(contour)
        (center, (w, h), angle) = rect
        if w * h < 500:
            continue
        if 1.0 <= max(w, h) / min(w, h) <= 3.0:
            cv2.drawContours(overlayed, [cv2.boxPoints(rect).astype(np.intp)], 0, (0, 255, 0), 4)
    return overlayed


class PostTap(threading.Thread):
    """Reads the blocks a module posts (module_<name>_post%idx%<post>#<colour space>) through the reference's own
    BlockAccessor, the way the reference's web GUI reader does (core/base.py:266-330)."""

    def __init__(self, rt, module_name, posts):
        super().__init__(daemon=True)
        self.rt, self.stop_flag, self.frames = rt, threading.Event(), {p: [] for p in posts}
        self.names = {p: "module_%s_post%%%d%%%s" % (module_name, i, p) for i, p in enumerate(posts)}

    def run(self):
        acc = {}
        shm_dir = "/dev/shm"
        while not self.stop_flag.is_set():
            for post, name in self.names.items():
                if post not in acc:
                    if not any(name in f for f in os.listdir(shm_dir)):
                        continue                       # not created yet (BlockAccessor.__enter__ would block and retry)
                    a = self.rt.cmf.BlockAccessor(name)
                    a.__enter__()
                    acc[post] = a
                try:
                    status, data, t = acc[post].read_frame()
                except Exception:  # noqa: BLE001  (block deleted while the module shuts down)
                    continue
                if data is not None and status == self.rt.cmf.ReadStatus.SUCCESS:
                    a = np.array(data, copy=True)
                    # the transport hands single-channel planes back as [H,W,1] (core/bindings/...py:334-365)
                    self.frames[post].append(a[..., 0] if a.ndim == 3 and a.shape[2] == 1 else a)
            time.sleep(0.002)
        # no __exit__ here: inside ONE process the reference's library hands every accessor of a direction the same Block
        # (static cmf_heap, lib/camera_message_framework_c.cpp:15,43-60), so a reader's delete_block would erase the
        # writer's block; the module's own ModuleManager.__exit__ deletes it


def run_module_under_reference_runtime(rt, module, direction, images, tmp_path, n_frames, posts):
    """Writes `images` as PNGs, starts the reference's image_directory capture source on `direction`, runs
    `module()` (ModuleBase.__call__) on the main thread until `n_frames` frames went through process(), returns
    (frames seen by process, {post name: [posted arrays]})."""
    d = tmp_path / "frames"
    d.mkdir()
    for i, im in enumerate(images):
        assert cv2.imwrite(str(d / ("%03d.png" % i)), im)
    # the capture source runs as its own process, as in the vehicle (and it has to: inside one process the reference's
    # library shares one Block per direction, so the module's ModuleManager.__exit__ would delete the writer's block):
    #   python capture_sources/image_directory.py <direction> <directory> --fps 40
    code = ("import sys; sys.path.insert(0, %r); import refruntime; refruntime.install(); "
            "sys.argv = ['image_directory.py', %r, %r, '--fps', '40']; "
            "from vision.capture_sources.image_directory import main; main()"
            % (os.path.dirname(os.path.abspath(__file__)), direction, str(d)))
    env = dict(os.environ, PYTHONPATH=refruntime.ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cap = subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    deadline = time.time() + 30
    while not any(direction in f for f in os.listdir("/dev/shm")):     # the block appears with the first frame
        assert cap.poll() is None, cap.stderr.read().decode()[-2000:]
        assert time.time() < deadline, "capture source did not come up"
        time.sleep(0.02)
    seen = []
    inner = module.process

    def counting_process(direction_, image):
        seen.append((direction_, image.copy()))
        out = inner(direction_, image)
        if len(seen) >= n_frames:
            # close the readers of the post blocks before their creator deletes them (the module's ModuleManager.__exit__),
            # then stop the module the way an operator does: SIGINT, whose handler ModuleBase.__call__ installed
            # (core/base.py:660-668) sets the loop's quit flag
            tap.stop_flag.set()
            tap.join(timeout=5)
            os.kill(os.getpid(), signal.SIGINT)
        return out
    module.process = counting_process
    tap = PostTap(rt, module._name, posts)
    tap.start()
    try:
        module()                               # ModuleBase.__call__: with ModuleManager -> thread(_loop) -> join
    finally:
        tap.stop_flag.set()
        tap.join(timeout=5)
        cap.send_signal(signal.SIGINT)          # "Ctrl-C Caught": run_event_loop's own handler (core/capture_source.py:96-99)
        try:
            cap.wait(timeout=10)
        except subprocess.TimeoutExpired:
            cap.kill()
    return seen, tap.frames


@pytest.fixture()
def rt(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["module"])        # ModuleBase.__init__ parses sys.argv (core/base.py:633)
    return refruntime.install()


def images_for_test():
    out = []
    for i in range(4):
        img = synth.gen_c5_frame(600 + i, 240, 320, big_target=bool(i & 1))
        img[30:100, 30:170] = (131, 164, 180)              # a beige box: bins.py accepts and draws its rectangle
        out.append(img)
    return out


def which(frame, images):
    for i, im in enumerate(images):
        if np.array_equal(frame, im):
            return i
    raise AssertionError("process() received a frame the capture source never sent")


def test_bin_detector_under_the_reference_module_base(rt, tmp_path):
    from cuauv_vision_pipeline_b200 import modules
    Bin, _, _ = modules.bind(rt.base.ModuleBase)
    assert issubclass(Bin, rt.base.ModuleBase) and Bin.__name__ == "BinDetectorGPU"
    direction = "b200rt%d" % os.getpid()
    images = images_for_test()
    m = Bin(video_sources=[direction], tuners=[], fps=200, pixels=make_pixels())
    seen, posted = run_module_under_reference_runtime(rt, m, direction, images, tmp_path, 5, ["bins#BGR"])
    assert len(seen) >= 3 and all(d == direction for d, _ in seen)
    assert m.pixels.uploads == len(seen), "one upload per process() call"
    assert len(posted["bins#BGR"]) >= 3
    # every posted image is the reference module's own post for one of the frames that went through process()
    refs = [reference_bins_post(images[which(f, images)]) for _, f in seen]
    for p in posted["bins#BGR"]:
        assert any(np.array_equal(p, r) for r in refs)
    assert sum(1 for r in refs if (r[..., 1] == 255).sum() > 50) >= 1, "a rectangle was drawn (bins.py:71-74)"


def test_buoy_and_color_balance_under_the_reference_module_base(rt, tmp_path):
    from cuauv_vision_pipeline_b200 import modules
    _, Buoy, Bal = modules.bind(rt.base.ModuleBase)
    direction = "b200rtb%d" % os.getpid()
    images = [synth.gen_underwater(240, 320, 700 + i) for i in range(3)]
    tuners = [rt.tuners.IntTuner("thresh_min", 150, 0, 255), rt.tuners.IntTuner("thresh_max", 255, 0, 255)]   # red_buoy.py:10-13
    m = Buoy([direction], tuners, fps=200, pixels=make_pixels())
    seen, posted = run_module_under_reference_runtime(rt, m, direction, images, tmp_path, 4, ["threshed#GRAY", "threshed_cleaned#GRAY"])
    assert len(seen) >= 3
    pairs = [cv_ops.buoy_mask(images[which(f, images)], 150, 255) for _, f in seen]
    assert len(posted["threshed#GRAY"]) >= 2 and len(posted["threshed_cleaned#GRAY"]) >= 2
    for p in posted["threshed#GRAY"]:
        assert any(np.array_equal(p, th) for th, _ in pairs)
    for p in posted["threshed_cleaned#GRAY"]:
        assert any(np.array_equal(p, cl) for _, cl in pairs)
    # normalize() is the reference's own (core/base.py:882-891): ((y - H/2)/W, (x - W/2)/W) of the last frame
    if m.result is not None:
        x, y = m.result["pixel"]
        assert m.result["center_y"] == pytest.approx((y - 120) / 320) and m.result["center_x"] == pytest.approx((x - 160) / 320)

    direction2 = "b200rtc%d" % os.getpid()
    (tmp_path / "second").mkdir()
    mb = Bal([direction2], [], fps=200, pixels=make_pixels())
    seen, posted = run_module_under_reference_runtime(rt, mb, direction2, images, tmp_path / "second", 4, ["orig#BGR", "balanced#BGR"])
    bal = [OraclePixels().balance(images[which(f, images)]) for _, f in seen]
    assert len(posted["balanced#BGR"]) >= 2
    for p in posted["balanced#BGR"]:
        assert any(np.array_equal(p, b) for b in bal)
    for p in posted["orig#BGR"]:
        assert any(np.array_equal(p, im) for im in images)
