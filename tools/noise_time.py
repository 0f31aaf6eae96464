"""Times the Gaussian-noise step (csrc/noise.cu) on one frame next to numpy's own randn on the host."""
import time
import numpy as np
import torch
from cuauv_vision_pipeline_b200.runtime import Context

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
