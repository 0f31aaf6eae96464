"""Drop-in through the reference's UNCHANGED transport: a capture-source thread writes frames into
the reference's own camera message framework (compiled unmodified, oracle/_ref), a module loop reads
them the way ModuleBase._loop does (core/base.py:711-844) and calls process(direction, frame) on the
GPU modules.  Also the "next" row f2 in small: the reader's library-owned buffer is pinned once
with bv_host_register and handed to the C ABI without the extra host copy."""
import os
import threading
import time

import numpy as np
import pytest

from oracle import cmf, cv_ops, synth

needs_cmf = pytest.mark.skipif(not cmf.available(), reason="oracle/_ref/libcamera_message_framework.so not built")


@needs_cmf
def test_reference_transport_round_trip():
    direction = "b200test_rt_%d" % os.getpid()
    img = synth.gen_underwater(120, 160, 1)
    w = cmf.Writer(direction, img.nbytes)
    try:
        r = cmf.Reader(direction)
        assert r.read()[0] == cmf.lib().NO_NEW_FRAME
        assert w.write(1234, img) == cmf.lib().SUCCESS
        st, t, view = r.read()
        assert st == cmf.lib().SUCCESS and t == 1234 and np.array_equal(view, img)
        assert r.read()[0] == cmf.lib().NO_NEW_FRAME
        r.close()
    finally:
        w.close()


@needs_cmf
@pytest.mark.gpu
def test_modules_fed_through_the_reference_transport(ctx):
    import cuauv_vision_pipeline_b200 as bv
    from cuauv_vision_pipeline_b200.modules import BinDetectorGPU, ColorBalanceGPU
    from cuauv_vision_pipeline_b200.runtime import ffi, lib, check
    from oracle import ref_balance, color_balance_np

    direction = "b200test_fwd_%d" % os.getpid()
    frames = [synth.gen_underwater(480, 640, 500 + i) for i in range(6)]
    writer = cmf.Writer(direction, frames[0].nbytes)
    stop = threading.Event()

    def capture_source():  # role of capture_sources/image_directory.py:30-36 + core/capture_source.py:183-234
        i = 0
        while not stop.is_set() and i < 400:                  # keeps cycling: a cold first call may take a while
            writer.write(1000 + i, frames[i % len(frames)])
            i += 1
            time.sleep(0.05)

    try:
        reader = cmf.Reader(direction)
        bins = BinDetectorGPU(video_sources=["forward"], tuners=[])
        bins_desc = ctx.make_stage(cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)])
        bal = ColorBalanceGPU(video_sources=["forward"])
        th = threading.Thread(target=capture_source)
        th.start()
        seen = {}
        pinned_ptr = None
        deadline = time.time() + 20
        while len(seen) < 3 and time.time() < deadline:      # the module loop (core/base.py:739-812)
            st, t, view = reader.read()
            if st != cmf.lib().SUCCESS:
                time.sleep(0.005)
                continue
            idx = (t - 1000) % len(frames)
            # zero-copy ingest: pin the reader's buffer once, hand the pointer straight to the C ABI
            if pinned_ptr != reader.data_pointer():
                if pinned_ptr is not None:
                    check(lib.bv_host_unregister(ffi.cast("void *", pinned_ptr)))
                pinned_ptr = reader.data_pointer()
                check(lib.bv_host_register(ffi.cast("void *", pinned_ptr), view.nbytes))
            direct = ctx.stage_host(bins_desc, view, want=("mask",))["mask"]
            frame = np.array(view)                            # the writable copy of core/base.py:765-768
            bins.process("forward", frame)
            balanced = bal.process("forward", frame)
            seen[idx] = (direct, bins.posted["bins"].copy(), balanced.copy(), len(bins.blobs))
        stop.set()
        th.join()
        if pinned_ptr is not None:
            check(lib.bv_host_unregister(ffi.cast("void *", pinned_ptr)))
        assert len(seen) >= 3, "module loop saw too few frames"
        for idx, (direct, posted, balanced, _) in seen.items():
            _, cleaned = cv_ops.bins_mask(frames[idx])
            assert np.array_equal(direct, cleaned)
            from test_gpu_balance_stage import reference_bins_post
            assert np.array_equal(posted, reference_bins_post(frames[idx])[0])      # modules/bins.py:81
            want = ref_balance.balance(frames[idx]) if ref_balance.available() else color_balance_np.process_frame_np(frames[idx])
            assert np.array_equal(balanced, want)
        reader.close()
    finally:
        stop.set()
        writer.close()
