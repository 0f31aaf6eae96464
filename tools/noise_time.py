"""Times the Gaussian-noise step (csrc/noise.cu) on one frame next to numpy's own randn on the host."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cuauv_vision_pipeline_b200.runtime import Context  # noqa: E402

ctx = Context()
for shape in ((480, 640, 3), (1242, 2208, 3), (2160, 3840, 3)):
    img = np.random.default_rng(1).integers(0, 256, shape, dtype=np.uint8)
    d = ctx.upload(img)
    np.random.seed(1)
    ctx.add_gaussian_noise(d, 5.0)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.add_gaussian_noise(d, 5.0)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    ref = np.clip(img + np.random.randn(*shape) * 5.0, 0., 255.).astype(np.uint8)
    t_ref = time.perf_counter() - t0
    print(f"{shape}: device {t_dev * 1e3:8.2f} ms   numpy {t_ref * 1e3:8.2f} ms")

# white_balance_bgr_blur (utils/color.py:381-391), device-resident, next to the reference expressions on the host
import cv2  # noqa: E402  (host comparison only)
from cuauv_vision_pipeline_b200 import color  # noqa: E402

for shape, k in (((1242, 2208, 3), 15), ((1242, 2208, 3), 101)):
    img = np.random.default_rng(2).integers(0, 256, shape, dtype=np.uint8)
    d = ctx.upload(img)
    color.white_balance_bgr_blur(d, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        color.white_balance_bgr_blur(d, k)
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB).astype(np.float32)
    l_, a_, b_ = cv2.split(lab)
    a_ -= cv2.blur(a_, (k, k), 0, borderType=cv2.BORDER_REPLICATE) - 128
    b_ -= cv2.blur(b_, (k, k), 0, borderType=cv2.BORDER_REPLICATE) - 128
    cv2.cvtColor(cv2.merge((l_, a_, b_)).astype(np.uint8), cv2.COLOR_LAB2BGR)
    t_ref = time.perf_counter() - t0
    print(f"white_balance_bgr_blur {shape} k={k}: device {t_dev * 1e3:8.3f} ms   cv2/numpy {t_ref * 1e3:8.2f} ms")
