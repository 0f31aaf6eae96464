#!/usr/bin/env python3
"""Repeats the 2208x1242 HSI-branch case and its balance-only half, comparing every run with the first one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402

ctx = bv.Context(0)
flags = dict(equalize_rgb=False, rgb_extrema_clipping=False)
img = synth.gen_underwater(1242, 2208, 6)
small = synth.gen_underwater(480, 640, 7)
ref_bal = ref_hsi = None
bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for it in range(N):
    if it % 7 == 0:   # what the neighbouring tests do: other sizes in between
        ctx.download(ctx.color_balance(ctx.upload(small), hsi_contrast_correct=True))
    dev = ctx.upload(img)
    bal = ctx.download(ctx.color_balance(dev, **flags))
    hsi = ctx.download(ctx.color_balance(ctx.upload(img), hsi_contrast_correct=True, **flags))
    if ref_bal is None:
        ref_bal, ref_hsi = bal, hsi
        continue
    for name, a, b in (("balance", bal, ref_bal), ("hsi", hsi, ref_hsi)):
        if not np.array_equal(a, b):
            d = np.argwhere(np.any(a != b, axis=2))
            bad += 1
            print("iteration %d: %s differs at %d pixels, rows %d..%d, cols %d..%d, first %s, flat index of first %d"
                  % (it, name, len(d), d[:, 0].min(), d[:, 0].max(), d[:, 1].min(), d[:, 1].max(), d[0],
                     d[0][0] * 2208 + d[0][1]), flush=True)
print("done: %d iterations, %d mismatches" % (N, bad))
ctx.close()
