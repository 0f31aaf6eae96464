"""TEST INFRASTRUCTURE ONLY -- seeded synthetic frames and masks.

The reference ships no images or fixtures; these generators give the parity tests and the bench
the same deterministic inputs everywhere (the GPU box has no /root/reference and no network).
"""
import numpy as np
import cv2


def gen_underwater(height, width, seed, targets=True):
    """Blue-green cast frame with optional red-buoy / beige-bin targets (uint8 BGR, HWC)."""
    rng = np.random.default_rng(seed)
    # low-res noise, blurred and upsampled: cheap stand-in for sigma=3 blur at full resolution
    sh, sw = max(8, height // 4), max(8, width // 4)
    base = rng.integers(0, 256, (sh, sw, 3)).astype(np.float32)
    base = cv2.GaussianBlur(base, (0, 0), 3)
    base = cv2.resize(base, (width, height), interpolation=cv2.INTER_LINEAR)
    lo, hi = float(base.min()), float(base.max())
    base = (base - lo) * (255.0 / max(hi - lo, 1e-6))
    img = base * np.array([0.9, 0.7, 0.35], np.float32) + np.array([30, 20, 5], np.float32)
    if targets:
        k = int(rng.integers(3, 41))
        for _ in range(k):
            cx, cy = int(rng.integers(0, width)), int(rng.integers(0, height))
            rad = int(np.exp(rng.uniform(np.log(4), np.log(max(5, height / 6)))))
            if rng.random() < 0.5:   # red buoy
                col = np.array([40, 40, 200], np.float32) + rng.uniform(-10, 10, 3).astype(np.float32)
                cv2.circle(img, (cx, cy), rad, tuple(float(c) for c in col), -1)
            else:                    # beige bin: HSV about (20, 60, 180) -> BGR
                hsv = np.uint8([[[int(rng.integers(12, 29)), int(rng.integers(30, 95)), int(rng.integers(90, 250))]]])
                col = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)[0, 0].astype(np.float32)
                ang = float(rng.uniform(0, 180))
                box = cv2.boxPoints(((cx, cy), (2.0 * rad, 1.0 * rad), ang)).astype(np.int32)
                cv2.fillPoly(img, [box], tuple(float(c) for c in col))
    noise = rng.normal(0, 6, (height, width, 1)).astype(np.float32)
    img = img + noise
    return np.clip(img, 0, 255).astype(np.uint8)


def gen_random_bgr(height, width, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (height, width, 3), dtype=np.uint8)


def all_colors_image():
    """4096x4096x3 image enumerating every (c0, c1, c2) triple exactly once (the 2^24 sweep)."""
    idx = np.arange(1 << 24, dtype=np.uint32).reshape(4096, 4096)
    img = np.empty((4096, 4096, 3), np.uint8)
    img[..., 0] = idx & 0xFF
    img[..., 1] = (idx >> 8) & 0xFF
    img[..., 2] = (idx >> 16) & 0xFF
    return img


def mask_blobs(height, width, seed, sigma=6.0, pct=70):
    """(i) threshold of blurred noise, opened 5x5: a few hundred irregular blobs."""
    rng = np.random.default_rng(seed)
    f = rng.random((height, width)).astype(np.float32)
    f = cv2.GaussianBlur(f, (0, 0), sigma)
    m = (f > np.percentile(f, pct)).astype(np.uint8) * 255
    return cv2.morphologyEx(m, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))


def mask_lattice(height, width):
    """(ii) isolated pixels on a 2-px lattice: ceil(W/2)*ceil(H/2) single-pixel blobs."""
    m = np.zeros((height, width), np.uint8)
    m[::2, ::2] = 255
    return m


def mask_serpentine(height, width):
    """(iii) one 1-px corridor snaking over the whole frame: a single blob with very long
    union chains."""
    m = np.zeros((height, width), np.uint8)
    m[::2, :] = 255
    for i, y in enumerate(range(1, height, 2)):
        m[y, width - 1 if i % 2 == 0 else 0] = 255
    return m


def mask_rings(height, width, step=6):
    """(v) nested square rings around the centre (blobs inside other blobs' holes)."""
    m = np.zeros((height, width), np.uint8)
    cy, cx = height // 2, width // 2
    for k in range(2, min(cy, cx) - 1, step):
        cv2.rectangle(m, (cx - k, cy - k), (cx + k, cy + k), 255, 1)
    return m


def mask_random(height, width, seed, density=0.5):
    rng = np.random.default_rng(seed)
    return (rng.random((height, width)) < density).astype(np.uint8) * 255


def mask_diagonals(height, width):
    """8-connectivity stress: anti-diagonal and diagonal 1-px lines that 4-connectivity would split."""
    m = np.zeros((height, width), np.uint8)
    yy, xx = np.mgrid[0:height, 0:width]
    m[((yy + xx) % 7 == 0) | ((yy - xx) % 11 == 0)] = 255
    return m


def gen_c5_frame(seed, height=2160, width=3840, big_target=True):
    """C5 frame (BASELINE.json configs[4]): an underwater frame plus, optionally, one large flat bin-coloured
    target with two holes in the right half.  After balance() -> BGR2HSV -> inRange([10,20,60],[30,100,255]) ->
    OPEN 5x5 (modules/bins.py:13-27) seed 9100 at 3840x2160 labels it as ONE blob of ~0.94 Mpx whose third-order
    moment m30 = 3.0e16 exceeds 2^53, so int64 accumulation is exercised."""
    img = gen_underwater(height, width, seed)
    if big_target:
        y0, y1 = height * 1100 // 2160, height * 2000 // 2160
        x0, x1 = width * 2600 // 3840, width * 3700 // 3840
        img[y0:y1, x0:x1] = (122, 113, 53)
        cv2.circle(img, (width * 3000 // 3840, height * 1500 // 2160), max(2, height * 120 // 2160), (30, 20, 5), -1)
        cv2.circle(img, (width * 3400 // 3840, height * 1300 // 2160), max(2, height * 40 // 2160), (30, 20, 5), -1)
    return img
