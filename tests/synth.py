"""Seeded synthetic images shared by the tests and bench.py:
u8 noise blended 50/50 with a smooth gradient, so both the filter interior and
the clamp on Lanczos overshoot are exercised.  RGBA / LA inputs: alpha = 255 for
7 of 8 images, random alpha for the eighth (exercises the f32 blend of overlay).
"""
import numpy as np


def synth_image(seed: int, h: int, w: int, c: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (h, w, c), dtype=np.uint16)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.empty((h, w, c), np.uint16)
    for k in range(c):
        g = (xx * (k + 1) * 255 // max(w - 1, 1) + yy * (3 - k % 3) * 255 // max(h - 1, 1)) % 511
        g = np.where(g > 255, 510 - g, g)
        img[..., k] = (noise[..., k] + g + 1) // 2
    img = img.astype(np.uint8)
    if c in (2, 4):
        if seed % 8 == 7:
            img[..., c - 1] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            img[..., c - 1] = 255
    return img
