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


def synth_deep(seed: int, h: int, w: int, c: int, dtype) -> np.ndarray:
    """The same picture with 16-bit or f32 subpixels (ImageLuma16 .. ImageRgba16, ImageRgb32F / ImageRgba32F): u16 =
    257 x the u8 value plus low-order noise (so that the samples are not multiples of 257), f32 = the u8 value scaled to
    [-0.1, 1.2] (HDR-like: resize and blur clamp to [0, 1], the colour ops do not).  Alpha as in synth_image."""
    a = synth_image(seed, h, w, c)
    dt = np.dtype(dtype)
    if dt == np.uint8:
        return a
    rng = np.random.default_rng(seed + 7919)
    if dt == np.uint16:
        b = (a.astype(np.int32) * 257 + rng.integers(-120, 121, a.shape)).clip(0, 65535).astype(np.uint16)
        if c in (2, 4) and seed % 8 != 7:
            b[..., c - 1] = 65535
        return b
    assert dt == np.float32
    f = (a.astype(np.float32) / np.float32(255)) * np.float32(1.3) - np.float32(0.1)
    if c == 4:
        f[..., 3] = np.float32(1.0) if seed % 8 != 7 else a[..., 3].astype(np.float32) / np.float32(255)
    return np.ascontiguousarray(f, dtype=np.float32)
