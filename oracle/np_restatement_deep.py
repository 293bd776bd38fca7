"""Second, independent restatement (numpy) of the image-0.25.6 arithmetic of fanlin-rs's pixel-transform path for the
16-bit and f32 DynamicImage variants.

TEST INFRASTRUCTURE ONLY (same rule as fanlin_oracle_deep.c, whose header lists what is restated and from where).  It is
vectorised over pixels, keeps the crate's per-tap sequential f32 multiply-then-add, and shares only the tap tables,
resize_dimensions, the Rgba<u8> overlay and the orientation with the u8 numpy restatement -- the subpixel-generic parts
(sampling store, luma, invert, conversions to u8, sequencing) are written here a second time.
"""
from __future__ import annotations

import numpy as np

from . import np_restatement as u8r

f32 = np.float32
MAXV = {np.dtype(np.uint8): 255, np.dtype(np.uint16): 65535, np.dtype(np.float32): 1.0}


def _round_half_away_nonneg(t32: np.ndarray) -> np.ndarray:
    # f32::round on t >= 0: floor(t + 0.5) with the sum taken exactly (f64 holds any f32 + 0.5)
    return np.floor(t32.astype(np.float64) + 0.5)


def _two_pass(img: np.ndarray, nw: int, nh: int, kind: str, sigma: float = 0.0) -> np.ndarray:
    h, w, c = img.shape
    dt = img.dtype
    src = img.astype(f32)  # `sample as f32`: exact for u8 / u16
    tmp = np.empty((nh, w, c), f32)
    for oy, (left, ws) in enumerate(u8r.taps(kind, h, nh, sigma)):
        t = np.zeros((w, c), f32)
        for i, wt in enumerate(ws):
            t = t + src[left + i] * wt
        tmp[oy] = t
    out = np.empty((nh, nw, c), dt)
    mx = f32(MAXV[dt])
    for ox, (left, ws) in enumerate(u8r.taps(kind, w, nw, sigma)):
        t = np.zeros((nh, c), f32)
        for i, wt in enumerate(ws):
            t = t + tmp[:, left + i, :] * wt
        t = np.where(t < f32(0), f32(0), np.where(t > mx, mx, t)).astype(f32)  # clamp(t, MIN, MAX)
        out[:, ox, :] = t if dt == np.float32 else _round_half_away_nonneg(t).astype(dt)
    return out


def resize(img, nw, nh, kind):
    h, w, _ = img.shape
    return img.copy() if (nw, nh) == (w, h) else _two_pass(img, nw, nh, kind)


def blur(img, sigma):
    h, w, _ = img.shape
    return _two_pass(img, w, h, "gaussian", 1.0 if sigma <= 0 else sigma)


def grayscale(img):
    c = img.shape[2]
    if c <= 2:
        return img.copy()
    if img.dtype == np.float32:  # Rgb32F / Rgba32F keep their type: luma in f64, replicated
        a = img.astype(np.float64)
        l = (((2126.0 * a[..., 0] + 7152.0 * a[..., 1]) + 722.0 * a[..., 2]) / 10000.0).astype(f32)
        out = img.copy()
        out[..., 0] = out[..., 1] = out[..., 2] = l
        return out
    a = img.astype(np.uint64)
    l = ((2126 * a[..., 0] + 7152 * a[..., 1] + 722 * a[..., 2]) // 10000).astype(img.dtype)
    return l[..., None] if c == 3 else np.stack([l, img[..., 3]], axis=-1)


def invert(img):
    out = img.copy()
    c = img.shape[2]
    ncol = c - 1 if c in (2, 4) else c
    mx = img.dtype.type(MAXV[img.dtype])
    out[..., :ncol] = mx - out[..., :ncol]
    return out


def sub_to_u8(a: np.ndarray) -> np.ndarray:
    """FromPrimitive<S> for u8."""
    if a.dtype == np.uint8:
        return a
    if a.dtype == np.uint16:
        return ((a.astype(np.uint32) + 128) // 257).astype(np.uint8)
    v = np.clip(a, f32(0), f32(1)).astype(f32) * f32(255.0)
    return _round_half_away_nonneg(v.astype(f32)).astype(np.uint8)


def to_rgba8(img):
    """pixel.to_rgba().into_color() per pixel == DynamicImage::to_rgba8."""
    return u8r.to_rgba8(sub_to_u8(img))


def process(img, *, w=None, h=None, rgb=(32, 32, 32), crop=False, blur_sigma=0.0, gray=False, inverse=False,
            orientation=1, to_rgb8=False, to_rgba8_out=False):
    if img.ndim == 2:
        img = img[:, :, None]
    img = np.ascontiguousarray(u8r.apply_orientation(img, orientation))
    if gray:
        img = grayscale(img)
    elif inverse:
        img = invert(img)
    if w is not None and h is not None:
        ih, iw, _ = img.shape
        if (w, h) != (iw, ih):
            if crop:
                w2, h2 = u8r.resize_dimensions(iw, ih, w, h, True)
                mid = resize(img, w2, h2, "lanczos3")
                if w * h2 > w2 * h:
                    y0 = (h2 - h) // 2
                    img = mid[y0:y0 + h, :w]
                else:
                    x0 = (w2 - w) // 2
                    img = mid[:h, x0:x0 + w]
            else:
                w2, h2 = u8r.resize_dimensions(iw, ih, w, h, False)
                img = resize(img, w2, h2, "lanczos3")
        ih, iw, _ = img.shape
        if w > iw or h > ih:
            bg = np.empty((h, w, 4), np.uint8)
            bg[...] = (rgb[0], rgb[1], rgb[2], 255)
            img = u8r.overlay(bg, to_rgba8(img), abs(w - iw) // 2, abs(h - ih) // 2)
    img = np.ascontiguousarray(img)
    if blur_sigma > 0:
        img = blur(img, blur_sigma)
    if to_rgba8_out:
        img = to_rgba8(img)
    elif to_rgb8 and not (img.dtype == np.uint8 and img.shape[2] == 3):
        img = to_rgba8(img)[:, :, :3]
    return np.ascontiguousarray(img)
