#!/usr/bin/env python3
"""End-to-end cost of getting ONE camera frame from the reference's shared-memory transport through the bins stage
(balance -> BGR2HSV -> inRange -> OPEN 5x5 -> labels + moments, blob table back), two ways:

  reference-style   Block::read_frame (memcpy to the heap, lib/camera_message_framework.cpp:445-449) -> np.array copy
                    (core/base.py:765-768) -> bv_stage_host (pageable H2D)
  ingest            CmfRing.ingest (one DMA from the pinned mapping, seqlock-validated) -> bv_stage on the device -> blob table D2H

The writer is the reference's transport compiled unmodified (oracle/_ref, TEST INFRASTRUCTURE used as the frame source).
    python tools/ingest_bench.py > gpurun_out/r02_ingest_e2e.log"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from cuauv_vision_pipeline_b200.cmf_ingest import CmfRing  # noqa: E402
from oracle import cmf, synth  # noqa: E402


def main():
    ctx = bv.Context(0)
    for (h, w) in [(1242, 2208), (1080, 1920), (480, 640)]:
        direction = "b200bench_%d_%d" % (os.getpid(), h)
        frames = [synth.gen_underwater(h, w, 900 + i) for i in range(4)]
        writer = cmf.Writer(direction, frames[0].nbytes)
        writer.write(1, frames[0])
        reader = cmf.Reader(direction)
        ring = CmfRing(direction)
        desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
        res = {}
        for name in ("reference-style", "ingest"):
            times, out, dev_out = [], {}, {}
            for i in range(40):
                writer.write(10 + i, frames[i % 4])
                t0 = time.perf_counter()
                if name == "reference-style":
                    st, t_acq, view = reader.read()
                    frame = np.array(view, copy=True)
                    out = ctx.stage_host(desc, frame, want=("blobs",), max_blobs=1024, out=out)
                    n = int(out["n_blobs"][0])
                else:
                    dev, t_acq, uid = ring.ingest(ctx)
                    dev_out = ctx.stage(desc, dev, want=("blobs",), max_blobs=1024, out=dev_out)
                    nb, tables = ctx.blobs_to_numpy(dev_out["blobs"], dev_out["n_blobs"])
                    n = int(nb[0])
                dt = time.perf_counter() - t0
                if i >= 8:
                    times.append(dt)
            res[name] = (np.median(times) * 1e3, n)
        a, b = res["reference-style"], res["ingest"]
        print("%dx%d  reference-style read+copy+stage_host: %.3f ms/frame (%.0f frames/s, %d blobs) | seqlock ingest + device stage: "
              "%.3f ms/frame (%.0f frames/s, %d blobs) | x%.2f, ingest retries %d"
              % (w, h, a[0], 1e3 / a[0], a[1], b[0], 1e3 / b[0], b[1], a[0] / b[0], ring.retries), flush=True)
        ring.close()
        reader.close()
        writer.close()
    ctx.close()


if __name__ == "__main__":
    main()
