"""Seqlock-aware ingest from the reference's shared-memory transport.

The writer is the reference's own library compiled unmodified (oracle/_ref/libcamera_message_framework.so through
oracle/cmf.py), i.e. `Block::write_frame` (lib/camera_message_framework.cpp:261-372).  The reader is
cuauv_vision_pipeline_b200.cmf_ingest.CmfRing + bv_ingest_seqlock: one DMA out of the mapped file, validated after the
copy."""
import os
import threading
import time

import numpy as np
import pytest

from oracle import cmf, synth

needs_cmf = pytest.mark.skipif(not cmf.available(), reason="oracle/_ref/libcamera_message_framework.so not built")


@needs_cmf
def test_layout_mirror_matches_the_compiled_transport():
    """CPU: the ctypes mirror of `struct Buffer` reads back exactly what the reference's writer stored."""
    from cuauv_vision_pipeline_b200 import cmf_ingest as ci
    direction = "b200lay_%d" % os.getpid()
    img = synth.gen_underwater(60, 80, 3)
    w = cmf.Writer(direction, img.nbytes)
    try:
        ring = ci.CmfRing(direction, pin=False)
        assert ring.slot_bytes == img.nbytes and ring.published() == 0 and not ring.deleted
        for k in range(1, 6):
            frame = np.roll(img, k, axis=1)
            assert w.write(1000 + k, frame) == cmf.lib().SUCCESS
            assert ring.published() == k
            seq, t_acq, planes = ring._meta(k)
            m = ring.header.metadata[k % ci.BUFFER_CNT]
            assert int(m.v_a) == int(m.v_b) == seq and t_acq == 1000 + k
            assert planes == [(80, 60, 3, 1, 0, "")]
            off = ci.DATA_OFFSET + (k % ci.BUFFER_CNT) * ring.slot_bytes
            assert np.array_equal(ring._view[off:off + img.nbytes].reshape(img.shape), frame)
        ring.close()
    finally:
        w.close()


@needs_cmf
@pytest.mark.gpu
def test_ingest_with_a_concurrent_writer_never_tears(ctx):
    from cuauv_vision_pipeline_b200.cmf_ingest import CmfRing
    direction = "b200ing_%d" % os.getpid()
    base = synth.gen_underwater(480, 640, 11)
    # frame k: the base image rolled by k columns with its index written into the first pixels: every byte identifies k
    frames = [np.ascontiguousarray(np.roll(base, 7 * k, axis=1)) for k in range(16)]
    for k, f in enumerate(frames):
        f[0, :8] = k
    writer = cmf.Writer(direction, frames[0].nbytes)
    stop = threading.Event()
    written = [0]

    def capture_source():
        k = 0
        while not stop.is_set():
            writer.write(5000 + k, frames[k % len(frames)])
            written[0] = k = k + 1
            if k % 64 == 0:
                time.sleep(0.0005)                       # otherwise a tight writer: it laps a 0.9 MB DMA all the time
    writer.write(4999, frames[0])
    th = threading.Thread(target=capture_source)
    ring = CmfRing(direction)
    try:
        th.start()
        got, t0 = 0, time.time()
        while got < 300 and time.time() - t0 < 30:
            r = ring.ingest(ctx, max_retries=64)
            if r is None:
                continue
            dev, t_acq, uid = r
            host = ctx.download(dev)
            k = int(host[0, 0, 0])
            assert k < len(frames) and np.array_equal(host, frames[k]), "torn frame"
            assert t_acq == 4999 or (t_acq - 5000) % len(frames) == k
            got += 1
        assert got >= 300
    finally:
        stop.set()
        th.join()
        ring.close()
        writer.close()
    # the writer lapped the reader at least once in 300 frames of a free-running loop, and the lap check caught it
    assert written[0] > 300
    print("ingested %d frames, writer wrote %d, retries %d" % (got, written[0], ring.retries))


@needs_cmf
@pytest.mark.gpu
def test_ingest_rgba_drops_alpha_on_the_device(ctx):
    """capture_sources/zed.py:49-50 / zed.cpp:54-71 (RGBA -> RGB) as part of the upload; swap_rb gives BGR."""
    from cuauv_vision_pipeline_b200.cmf_ingest import CmfRing
    direction = "b200rgba_%d" % os.getpid()
    rgba = np.random.default_rng(5).integers(0, 256, (243, 325, 4), dtype=np.uint8)
    writer = cmf.Writer(direction, rgba.nbytes)
    try:
        writer.write(1, rgba)
        with CmfRing(direction) as ring:
            dev, t_acq, uid = ring.ingest(ctx)
            assert np.array_equal(ctx.download(dev), rgba[..., :3]) and t_acq == 1 and uid == 1
            assert ring.ingest(ctx) is None                     # NO_NEW_FRAME
            writer.write(2, rgba)
            dev, _, _ = ring.ingest(ctx, swap_rb=True)
            assert np.array_equal(ctx.download(dev), rgba[..., 2::-1])
    finally:
        writer.close()
