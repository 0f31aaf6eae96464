#!/usr/bin/env python3
"""Runs the test sequence around the intermittent HSI failure in one process, many times, and diagnoses a failure."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
import test_gpu_balance_stage as T  # noqa: E402
from oracle import synth  # noqa: E402

ctx = bv.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cases = [("underwater", (480, 640), 7, {}), ("random", (480, 640), 4, dict(hsv_contrast_correct=False)),
         ("underwater", (1242, 2208), 6, dict(equalize_rgb=False, rgb_extrema_clipping=False)),
         ("underwater", (479, 641), 9, dict(rgb_contrast_correct=True)),
         ("underwater", (480, 640), 11, dict(horizontal_blocks=4, vertical_blocks=2))]
img = synth.gen_underwater(1242, 2208, 6)
want = T.oracle_balance(img, hsi_contrast_correct=True, equalize_rgb=False, rgb_extrema_clipping=False)
fails = 0
for it in range(N):
    try:
        T.test_stage_mask_only_many_bounds_reuse_and_evict_tables(ctx)
        for k, (kind, shape, seed, flags) in enumerate(cases):
            if k == 2:   # the failing case, with the oracle result computed once
                got = ctx.download(ctx.color_balance(ctx.upload(img), hsi_contrast_correct=True, **flags))
                diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
                if int(diff.max()) > 1:
                    fails += 1
                    d = np.argwhere(np.any(diff > 1, axis=2))
                    print("iteration %d: %d pixels off by > 1 (max %d); rows %d..%d cols %d..%d; first %s" %
                          (it, len(d), int(diff.max()), d[:, 0].min(), d[:, 0].max(), d[:, 1].min(), d[:, 1].max(), d[0]), flush=True)
                    bal = ctx.download(ctx.color_balance(ctx.upload(img), **flags))
                    again = ctx.download(ctx.color_balance(ctx.upload(img), hsi_contrast_correct=True, **flags))
                    print("   repeated at once: max diff %d; balance-only output equals oracle balance: %s" %
                          (int(np.abs(again.astype(np.int16) - want.astype(np.int16)).max()),
                           np.array_equal(bal, T.oracle_balance(img, **flags))), flush=True)
                    rows = np.unique(d[:, 0])
                    print("   distinct rows %d, distinct cols %d, got sample %s want %s" % (len(rows), len(np.unique(d[:, 1])), got[d[0][0], d[0][1]], want[d[0][0], d[0][1]]), flush=True)
            else:
                T.test_balance_hsi_branch(ctx, kind, shape, seed, flags)
    except Exception:  # noqa: BLE001
        fails += 1
        traceback.print_exc()
print("done: %d iterations, %d failures" % (N, fails))
ctx.close()
