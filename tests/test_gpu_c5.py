"""GPU parity for BASELINE.json configs[4] ("C5") at its full size: 3840x2160 frames through
balance() -> BGR2HSV -> inRange([10,20,60],[30,100,255]) -> OPEN 5x5 -> labels + raster moments
(modules/bins.py:13-27 behind modules/preprocessor.py:87-88; blobs per SURVEY.md 8c), the CUDA
stage against the compiled reference process_frame + cv2 + oracle/ccl.py.  Bit-exact: mask, labels,
all ten int64 moments, bounding boxes; the large blob's third-order moments are additionally
recomputed with Python integers (m30 = 3.0e16 > 2^53, where float64 is no longer exact)."""
import cv2
import numpy as np
import pytest

from oracle import ccl, color_balance_np as cb, cv_ops, ref_balance, synth

pytestmark = pytest.mark.gpu
LO, HI = (10, 20, 60), (30, 100, 255)


def oracle_chain(img):
    bal = ref_balance.balance(img) if ref_balance.available() else cb.process_frame_np(img)
    hsv = cv2.cvtColor(bal, cv2.COLOR_BGR2HSV)
    raw = cv2.inRange(hsv, np.array(LO), np.array(HI))
    mask = cv2.morphologyEx(raw, cv2.MORPH_OPEN, cv_ops.rect_kernel(5))
    n, lab, tab = ccl.label_and_moments(mask)
    return mask, n, lab, tab


def python_int_moments(lab, label):
    ys, xs = np.nonzero(lab == label)
    cx = np.bincount(xs)
    cy = np.bincount(ys)
    m30 = sum(int(c) * x ** 3 for x, c in enumerate(cx.tolist()) if c)
    m03 = sum(int(c) * y ** 3 for y, c in enumerate(cy.tolist()) if c)
    pairs = {}
    for x, y in zip(xs.tolist()[::1], ys.tolist()[::1]):
        pairs[y] = pairs.get(y, 0) + x * x          # sum of x^2 per row
    m21 = sum(y * s for y, s in pairs.items())
    return m30, m03, m21


def test_c5_stage_at_3840x2160_vs_oracle(ctx):
    frames = np.stack([synth.gen_c5_frame(9100), synth.gen_c5_frame(9101), synth.gen_c5_frame(9102, big_target=False)])
    desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=LO, hi=HI, morph=[("open", 5, 5, 1)], label=True)
    out = ctx.stage(desc, ctx.upload(frames), want=("mask", "labels", "blobs"), max_blobs=8192)
    mask = ctx.download(out["mask"])
    lab = ctx.download(out["labels"])
    n, tables = ctx.blobs_to_numpy(out["blobs"], out["n_blobs"])
    saw_big = False
    for i in range(frames.shape[0]):
        m_ref, n_ref, lab_ref, tab = oracle_chain(frames[i])
        assert np.array_equal(mask[i], m_ref), "mask of frame %d" % i
        assert int(n[i]) == n_ref and n_ref <= 8192
        assert np.array_equal(lab[i], lab_ref), "labels of frame %d" % i
        for key in ccl.MOMENT_KEYS + ("x0", "y0", "x1", "y1"):
            assert np.array_equal(tables[i][key].astype(np.int64), tab[key]), (i, key)
        big = int(np.argmax(tab["m30"])) if n_ref else -1
        if big >= 0 and int(tab["m30"][big]) > 2 ** 53:
            saw_big = True
            m30, m03, m21 = python_int_moments(lab_ref, big + 1)
            assert (int(tables[i]["m30"][big]), int(tables[i]["m03"][big]), int(tables[i]["m21"][big])) == (m30, m03, m21)
    assert saw_big, "the C5 fixture must contain a blob with m30 > 2^53"


def test_c5_host_entry_sharded_streams_equal_device_stage(ctx):
    """The stream-sharded C5 leg of bench.py goes through stage() per stream; a stream's frames must not
    depend on what else is in the batch (statistics are per frame, color_balance.cpp:396-428)."""
    frames = np.stack([synth.gen_c5_frame(9200 + s, 1080, 1920, big_target=bool(s & 1)) for s in range(4)])
    desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=LO, hi=HI, morph=[("open", 5, 5, 1)], label=True)
    whole = ctx.stage(desc, ctx.upload(frames), want=("mask", "labels"), max_blobs=2048)
    mask_all, lab_all = ctx.download(whole["mask"]), ctx.download(whole["labels"])
    for s in range(4):
        one = ctx.stage(desc, ctx.upload(frames[s:s + 1]), want=("mask", "labels"), max_blobs=2048)
        assert np.array_equal(ctx.download(one["mask"])[0], mask_all[s])
        assert np.array_equal(ctx.download(one["labels"])[0], lab_all[s])
        m_ref, n_ref, lab_ref, _ = oracle_chain(frames[s])
        assert np.array_equal(mask_all[s], m_ref) and np.array_equal(lab_all[s], lab_ref)


@pytest.mark.parametrize("shape,batch", [((1080, 1920), 12), ((1000, 1050), 8)])
def test_unbalanced_stage_chunks_on_side_streams_equal_cv2_and_single_frames(ctx, shape, batch):
    """BASELINE.json configs[2] ("C3": HSV inRange -> OPEN -> labels) as a batch: without colour balance the stage splits
    the batch into chunks of whole frames on the side streams (pixel pass, morphology and labelling of one chunk overlap
    the other chunks').  1920 wide: bits straight from the pixel pass; 1050 wide (width % 16 != 0): through the uint8 mask.
    Every frame against cv2 + oracle/ccl.py, and the same frames one call each."""
    h, w = shape
    frames = np.stack([synth.gen_underwater(h, w, 300 + i) for i in range(batch)])
    desc = ctx.make_stage(cvt="bgr2hsv", lo=LO, hi=HI, morph=[("open", 5, 5, 1)], label=True)
    want = ("converted", "mask", "labels", "blobs")
    out = ctx.stage(desc, ctx.upload(frames), want=want, max_blobs=4096)
    got = {k: ctx.download(out[k]) for k in ("converted", "mask", "labels")}
    n, tables = ctx.blobs_to_numpy(out["blobs"], out["n_blobs"])
    for i in range(batch):
        hsv = cv2.cvtColor(frames[i], cv2.COLOR_BGR2HSV)
        m_ref = cv2.morphologyEx(cv2.inRange(hsv, np.array(LO), np.array(HI)), cv2.MORPH_OPEN, cv_ops.rect_kernel(5))
        n_ref, lab_ref, tab = ccl.label_and_moments(m_ref)
        assert np.array_equal(got["converted"][i], hsv), i
        assert np.array_equal(got["mask"][i], m_ref), i
        assert int(n[i]) == n_ref and np.array_equal(got["labels"][i], lab_ref), i
        for key in ccl.MOMENT_KEYS + ("x0", "y0", "x1", "y1"):
            assert np.array_equal(tables[i][key].astype(np.int64), tab[key]), (i, key)
    for i in (0, batch - 1):     # one frame per call: no chunking at all
        one = ctx.stage(desc, ctx.upload(frames[i]), want=("mask", "labels"), max_blobs=4096)
        assert np.array_equal(ctx.download(one["mask"]), got["mask"][i])
        assert np.array_equal(ctx.download(one["labels"]), got["labels"][i])
    # the host entry point (uploads chunked, too) delivers the same
    host = ctx.stage_host(desc, frames, want=("mask", "labels"), max_blobs=4096)
    assert np.array_equal(host["mask"], got["mask"]) and np.array_equal(host["labels"], got["labels"])
