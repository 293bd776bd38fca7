"""Host-side mirror of the reference's two callers of the stage:
process_image's pixel section (src/handler.rs:224-255) and process_gif's
per-frame closure (src/handler.rs:321-357).  Decode and encode stay outside; these
take and return decoded pixels, like the DynamicImage the Rust code holds."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .device import SAMPLE_DTYPES, Device, Job, lib, plan_job
from .query import Query


def _as_img(a) -> np.ndarray:
    """u8 (ImageLuma8 .. ImageRgba8), u16 (ImageLuma16 .. ImageRgba16) or f32 (ImageRgb32F / ImageRgba32F) pixels;
    anything else is taken as u8, as before."""
    a = np.asarray(a)
    a = np.ascontiguousarray(a, dtype=a.dtype if a.dtype in (np.uint8, np.uint16, np.float32) else np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3 or not 1 <= a.shape[2] <= 4:
        raise ValueError("image must be (H, W) or (H, W, C<=4) u8 / u16 / f32")
    return a


TO_RGB8 = 1 << 5  # enum fanlin_flags FANLIN_TO_RGB8
TO_RGBA8 = 1 << 4
TO_YCBCR = 1 << 6  # FANLIN_TO_YCBCR: the JPEG encoder's planes of the result, (3, H, W)


def make_job(img: np.ndarray, params: Query, *, gif: bool = False, orientation: int = 1, to_rgb8: bool = False,
             to_rgba8: bool = False, to_ycbcr: bool = False) -> Job:
    """orientation: the EXIF value decoder.orientation() reported for a still (src/handler.rs:206);
    the device turns the image instead of img.apply_orientation(o) on the host (:221-223).
    to_rgb8: the caller will encode JPEG (:274-278) and wants the RGB8 the encoder works on;
    to_rgba8: the caller will encode WebP (:287 into_rgba8)."""
    a = _as_img(img)
    j = Job()
    lib().fanlin_job_from_query(C.byref(params._q), int(gif), C.byref(j))
    j.orientation = int(orientation)
    if to_rgb8:
        j.flags |= TO_RGB8
    if to_rgba8:
        j.flags |= TO_RGBA8
    if to_ycbcr:
        j.flags |= TO_YCBCR
    j.src = a.ctypes.data
    j.src_h, j.src_w, j.src_channels = a.shape
    j.src_sample = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[a.dtype]
    j._keep = a
    return j


def _run(dev: Device, jobs):
    outs = []
    for j in jobs:
        p = plan_job(j)
        planar = bool(p.stages & 64)  # FANLIN_TO_YCBCR: (3, H, W)
        o = np.empty((3, p.out_h, p.out_w) if planar else (p.out_h, p.out_w, p.out_channels), SAMPLE_DTYPES[p.out_sample])
        j.dst, j.dst_capacity = o.ctypes.data, o.nbytes
        outs.append(o)
    arr = (Job * len(jobs))()
    for i, j in enumerate(jobs):
        C.memmove(C.byref(arr, i * C.sizeof(Job)), C.byref(j), C.sizeof(Job))
    dev.run(arr)
    return outs


def process_image(dev: Device, img: np.ndarray, params: Query, *, orientation: int = 1, to_rgb8: bool = False,
                  to_rgba8: bool = False, to_ycbcr: bool = False) -> np.ndarray:
    """Pixel section of State::process_image: decoded pixels (as stored, with their EXIF
    orientation) in, transformed pixels out (RGB8 when the JPEG encoder follows, RGBA8 for WebP).
    The array's dtype is the DynamicImage variant's subpixel type (u8, u16, f32)."""
    return _run(dev, [make_job(img, params, orientation=orientation, to_rgb8=to_rgb8, to_rgba8=to_rgba8, to_ycbcr=to_ycbcr)])[0]


def process_images(dev: Device, imgs, params: Query):
    """Several decoded images with the same request in one ragged launch (what the batcher
    does with concurrent requests); same-shaped images share CTA pairs on the device."""
    return _run(dev, [make_job(im, params) for im in imgs])


def process_gif_frames(dev: Device, frames, params: Query):
    """All composited RGBA8 frames of a GIF in one ragged launch; RGBA8 frames out."""
    return _run(dev, [make_job(f, params, gif=True) for f in frames])
