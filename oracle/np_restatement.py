"""Second, independent restatement (numpy) of the image-0.25.6 arithmetic on
fanlin-rs's pixel-transform path.

TEST INFRASTRUCTURE ONLY (same rule as fanlin_oracle.c).  Its one job is to be a
differently-structured implementation of SURVEY.md Appendix A, so that a
transcription error in either restatement shows up as a bit difference between
them (tests/test_oracle.py).  It is vectorised across pixels but keeps the
crate's per-tap sequential f32 multiply-then-add order; sinf/expf come from the
same glibc the C oracle (and Rust's f32::sin/exp on linux-gnu) uses.

Follows: imageops/sample.rs (resize, blur, vertical_sample, horizontal_sample),
math/utils.rs (resize_dimensions), color.rs (rgb_to_luma, Blend, Invert),
imageops/mod.rs (overlay, crop), dynimage.rs (resize, resize_to_fill), and the
sequencing of reference src/handler.rs:224-255 / :329-355.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]
_libm.expf.restype = ctypes.c_float
PI = f32(math.pi)


def _round_half_away(v: float) -> int:
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def resize_dimensions(width, height, nwidth, nheight, fill):
    wratio = nwidth / width
    hratio = nheight / height
    ratio = max(wratio, hratio) if fill else min(wratio, hratio)
    nw = max(_round_half_away(width * ratio), 1)
    nh = max(_round_half_away(height * ratio), 1)
    u32max = 2**32 - 1
    if nw > u32max:
        r = u32max / width
        return u32max, max(_round_half_away(height * r), 1)
    if nh > u32max:
        r = u32max / height
        return max(_round_half_away(width * r), 1), u32max
    return nw, nh


def _sinc(t):
    a = f32(t * PI)
    if t == 0:
        return f32(1.0)
    return f32(f32(_libm.sinf(a)) / a)


def _lanczos3(x):
    if abs(x) < 3.0:
        return f32(_sinc(x) * _sinc(f32(x / f32(3.0))))
    return f32(0.0)


def _gaussian(x, r):
    lhs = f32(f32(1.0) / f32(np.sqrt(f32(f32(2.0) * PI)) * r))
    e = f32(f32(-f32(x * x)) / f32(f32(2.0) * f32(r * r)))
    return f32(lhs * f32(_libm.expf(e)))


def taps(kind, n_in, n_out, sigma=0.0):
    """[(left, weights f32[n])] for every output index of one axis."""
    if kind == "nearest":
        support, k = f32(0.0), (lambda x: f32(1.0))
    elif kind == "lanczos3":
        support, k = f32(3.0), _lanczos3
    elif kind == "gaussian":
        sg = f32(sigma)
        support, k = f32(f32(2.0) * sg), (lambda x: _gaussian(x, sg))
    else:
        raise ValueError(kind)
    ratio = f32(f32(n_in) / f32(n_out))
    sratio = f32(1.0) if ratio < 1 else ratio
    src_support = f32(support * sratio)
    out = []
    for o in range(n_out):
        c = f32(f32(f32(o) + f32(0.5)) * ratio)
        left = int(math.floor(f32(c - src_support)))
        left = min(max(left, 0), n_in - 1)
        right = int(math.ceil(f32(c + src_support)))
        right = min(max(right, left + 1), n_in)
        c = f32(c - f32(0.5))
        ws = np.empty(right - left, f32)
        s = f32(0.0)
        for j, i in enumerate(range(left, right)):
            w = k(f32(f32(f32(i) - c) / sratio))
            ws[j] = w
            s = f32(s + w)
        ws = (ws / s).astype(f32)
        out.append((left, ws))
    return out


def vertical_sample(img_u8, new_h, kind, sigma=0.0):
    h, w, c = img_u8.shape
    src = img_u8.astype(f32)
    out = np.empty((new_h, w, c), f32)
    for oy, (left, ws) in enumerate(taps(kind, h, new_h, sigma)):
        t = np.zeros((w, c), f32)
        for i, wt in enumerate(ws):
            t = t + src[left + i] * wt  # f32*f32 rounded, then f32+f32 rounded
        out[oy] = t
    return out


def horizontal_sample(tmp, new_w, kind, sigma=0.0):
    h, w, c = tmp.shape
    out = np.empty((h, new_w, c), np.uint8)
    for ox, (left, ws) in enumerate(taps(kind, w, new_w, sigma)):
        t = np.zeros((h, c), f32)
        for i, wt in enumerate(ws):
            t = t + tmp[:, left + i, :] * wt
        t = np.clip(t, f32(0), f32(255))
        # f32::round = half away from zero; t >= 0 here, and the +0.5 is exact in f64
        out[:, ox, :] = np.floor(t.astype(np.float64) + 0.5).astype(np.uint8)
    return out


def resize(img, nw, nh, kind):
    h, w, _ = img.shape
    if (nw, nh) == (w, h):
        return img.copy()
    return horizontal_sample(vertical_sample(img, nh, kind), nw, kind)


def blur(img, sigma):
    h, w, _ = img.shape
    sigma = 1.0 if sigma <= 0 else sigma
    return horizontal_sample(vertical_sample(img, h, "gaussian", sigma), w, "gaussian", sigma)


def grayscale(img):
    c = img.shape[2]
    if c <= 2:
        return img.copy()
    a = img.astype(np.uint32)
    l = ((2126 * a[..., 0] + 7152 * a[..., 1] + 722 * a[..., 2]) // 10000).astype(np.uint8)
    if c == 3:
        return l[..., None]
    return np.stack([l, img[..., 3]], axis=-1)


def invert(img):
    out = img.copy()
    c = img.shape[2]
    ncol = c - 1 if c in (2, 4) else c
    out[..., :ncol] = 255 - out[..., :ncol]
    return out


def to_rgba8(img):
    h, w, c = img.shape
    out = np.empty((h, w, 4), np.uint8)
    if c <= 2:
        out[..., 0] = out[..., 1] = out[..., 2] = img[..., 0]
        out[..., 3] = 255 if c == 1 else img[..., 1]
    else:
        out[..., :3] = img[..., :3]
        out[..., 3] = 255 if c == 3 else img[..., 3]
    return out


def overlay(bottom, top, x, y):
    """bottom RGBA8 (modified copy returned); x, y >= 0 as in handler.rs:241-246."""
    out = bottom.copy()
    bh, bw, _ = out.shape
    fg = to_rgba8(top)
    th, tw, _ = fg.shape
    rw, rh = min(tw, bw - x), min(th, bh - y)
    if rw <= 0 or rh <= 0:
        return out
    fg = fg[:rh, :rw]
    bgv = out[y:y + rh, x:x + rw]
    m = f32(255.0)
    bgf = bgv.astype(f32) / m
    fgf = fg.astype(f32) / m
    ba, fa = bgf[..., 3], fgf[..., 3]
    a_final = (ba + fa) - ba * fa
    res = bgv.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(3):
            o = (fgf[..., k] * fa + (bgf[..., k] * ba) * (f32(1.0) - fa)) / a_final
            res[..., k] = np.trunc(m * o).astype(np.uint8)
        res[..., 3] = np.trunc(m * a_final).astype(np.uint8)
    opaque = fg[..., 3] == 255
    clear = (fg[..., 3] == 0) | (a_final == 0)
    res[opaque] = fg[opaque]
    res[clear] = bgv[clear]
    out[y:y + rh, x:x + rw] = res
    return out


def apply_orientation(img, exif):
    """Orientation::from_exif + apply_orientation, written with whole-array operations
    (independent of the per-pixel put_pixel loops of the C restatement): rotate90 is clockwise."""
    if exif in (0, 1):
        return img
    if exif == 2:
        return img[:, ::-1]
    if exif == 3:
        return img[::-1, ::-1]
    if exif == 4:
        return img[::-1]
    if exif == 5:  # Rotate90FlipH = transpose
        return np.rot90(img, k=-1)[:, ::-1]
    if exif == 6:
        return np.rot90(img, k=-1)
    if exif == 7:  # Rotate270FlipH = transverse
        return np.rot90(img, k=1)[:, ::-1]
    if exif == 8:
        return np.rot90(img, k=1)
    raise ValueError("orientation")


def process(img, *, w=None, h=None, rgb=(32, 32, 32), crop=False, blur_sigma=0.0, gray=False,
            inverse=False, gif=False, orientation=1, to_rgb8=False):
    if img.ndim == 2:
        img = img[:, :, None]
    kind = "nearest" if gif else "lanczos3"
    if not gif:
        img = np.ascontiguousarray(apply_orientation(img, orientation))
    if gray:
        img = grayscale(img)
    elif inverse:
        img = invert(img)
    if w is not None and h is not None:
        ih, iw, _ = img.shape
        if (w, h) != (iw, ih):
            if crop:
                w2, h2 = resize_dimensions(iw, ih, w, h, True)
                mid = resize(img, w2, h2, kind)
                if w * h2 > w2 * h:
                    y0 = (h2 - h) // 2
                    img = mid[y0:y0 + h, :w]
                else:
                    x0 = (w2 - w) // 2
                    img = mid[:h, x0:x0 + w]
            else:
                w2, h2 = resize_dimensions(iw, ih, w, h, False)
                img = resize(img, w2, h2, kind)
        ih, iw, _ = img.shape
        if w > iw or h > ih:
            bg = np.empty((h, w, 4), np.uint8)
            bg[...] = (rgb[0], rgb[1], rgb[2], 255)
            img = overlay(bg, img, abs(w - iw) // 2, abs(h - ih) // 2)
    if not gif and blur_sigma > 0:
        img = blur(img, blur_sigma)
    if gif:
        img = to_rgba8(img)
    elif to_rgb8 and img.shape[2] != 3:  # DynamicImage::to_rgb8: alpha dropped, luma replicated
        img = np.repeat(img[:, :, :1], 3, axis=2) if img.shape[2] <= 2 else img[:, :, :3]
    return np.ascontiguousarray(img)
