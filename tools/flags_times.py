#!/usr/bin/env python3
"""Device-resident time per 2208x1242 frame of bv_color_balance for the flag combinations of process_frame
(color_balance.cpp:343-780): CUDA events around 10 calls on 16 frames."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuauv_vision_pipeline_b200 as bv  # noqa: E402
from oracle import synth  # noqa: E402  (input generator only)

ctx = bv.Context(0)
H, W, N = 1242, 2208, 16
base = np.stack([synth.gen_underwater(H, W, 7000 + i) for i in range(4)])
frames = ctx.upload(np.concatenate([np.roll(base, 7 * k, axis=2) for k in range(N // 4)]))
CASES = [
    ("default flags", {}),
    ("rgb_contrast_correct off", dict(rgb_contrast_correct=0)),
    ("hsv_contrast_correct off", dict(hsv_contrast_correct=0)),
    ("equalize_rgb off", dict(equalize_rgb=0)),
    ("rgb_extrema_clipping on", dict(rgb_extrema_clipping=1)),
    ("adaptive_cast_correction on", dict(adaptive_cast_correction=1)),
    ("tiles 2 x 2", dict(horizontal_blocks=2, vertical_blocks=2)),
    ("tiles 8 x 6", dict(horizontal_blocks=8, vertical_blocks=6)),
    ("tiles 8 x 6 + adaptive cast", dict(horizontal_blocks=8, vertical_blocks=6, adaptive_cast_correction=1)),
    ("everything off", dict(equalize_rgb=0, rgb_contrast_correct=0, hsv_contrast_correct=0)),
    ("hsi_contrast_correct on", dict(hsi_contrast_correct=1)),
]
for name, flags in CASES:
    try:
        for _ in range(3):
            ctx.color_balance(frames, **flags)
        ctx.sync()
        reps = 3 if flags.get("hsi_contrast_correct") else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            ctx.color_balance(frames, **flags)
        e1.record(ctx.torch_stream)
        ctx.sync()
        print("%-34s %8.1f us/frame" % (name, e0.elapsed_time(e1) * 1e3 / reps / N), flush=True)
    except Exception as exc:  # noqa: BLE001
        print("%-34s %s" % (name, exc))
