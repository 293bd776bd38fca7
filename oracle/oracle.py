"""ctypes front end of the CPU oracle (oracle/fanlin_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from fanlin-rs_b200/.
PARITY UNPINNED: see the header of fanlin_oracle.c.

The functions mirror the reference call sites they restate:
  process()        src/handler.rs:224-255 (still) / :329-355 (GIF frame)
  resize_dimensions, resize, blur, grayscale, invert, overlay, to_rgba8
                   the image-0.25.6 functions those lines call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfanlin_oracle.so")

NEAREST, LANCZOS3, GAUSSIAN_BLUR = 0, 1, 100
GRAYSCALE, INVERSE, HAS_DIMS, CROP, GIF_FRAME, TO_RGB8 = 1, 2, 4, 8, 16, 32


class Job(C.Structure):
    _fields_ = [
        ("src", C.c_void_p),
        ("src_w", C.c_uint32), ("src_h", C.c_uint32), ("src_c", C.c_uint32),
        ("flags", C.c_uint32),
        ("req_w", C.c_uint32), ("req_h", C.c_uint32),
        ("fill", C.c_uint8 * 3), ("orientation", C.c_uint8),
        ("blur_sigma", C.c_float),
        ("dst", C.c_void_p),
        ("dst_cap", C.c_uint64),
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("out_c", C.c_uint32),
    ]


class DeepJob(C.Structure):
    """fod_job of fanlin_oracle_deep.c: the request on a DynamicImage variant of any subpixel type."""
    _fields_ = [
        ("src", C.c_void_p),
        ("src_w", C.c_uint32), ("src_h", C.c_uint32), ("src_c", C.c_uint32), ("sample", C.c_uint32),
        ("flags", C.c_uint32),
        ("req_w", C.c_uint32), ("req_h", C.c_uint32),
        ("fill", C.c_uint8 * 3), ("orientation", C.c_uint8),
        ("blur_sigma", C.c_float),
        ("dst", C.c_void_p),
        ("dst_cap", C.c_uint64),
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("out_c", C.c_uint32), ("out_sample", C.c_uint32),
    ]


TO_RGBA8 = 64  # fod_process only: DynamicImage::into_rgba8 of a still (the WebP branch, handler.rs:287)
TO_YCBCR = 128  # fod_process only: the JPEG encoder's Y, Cb, Cr planes of the result ([3][h][w])
SAMPLE_DTYPES = {0: np.uint8, 1: np.uint16, 2: np.float32}


def sample_kind(dtype) -> int:
    return {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[np.dtype(dtype)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("fanlin_oracle.c", "fanlin_oracle_deep.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        assert L.fo_job_size() == C.sizeof(Job), "fo_job layout mismatch"
        L.fo_process.argtypes = [C.POINTER(Job)]
        L.fo_process_batch.argtypes = [C.POINTER(Job), C.c_uint32, C.c_uint32]
        L.fo_resize_dimensions.argtypes = [C.c_uint32] * 4 + [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.fo_resize_dimensions.restype = None
        L.fo_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.fo_blur.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
        L.fo_grayscale.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.fo_grayscale.restype = C.c_uint32
        L.fo_invert.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.fo_invert.restype = None
        L.fo_to_rgba8.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.fo_to_rgba8.restype = None
        L.fo_overlay.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64]
        L.fo_overlay.restype = None
        L.fo_weight_table.argtypes = [C.c_int, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.fo_weight_table.restype = C.c_uint32
        L.fo_ycck_to_cmyk.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_ycck_to_cmyk.restype = None
        L.fo_apply_orientation.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        assert L.fod_job_size() == C.sizeof(DeepJob), "fod_job layout mismatch"
        L.fod_process.argtypes = [C.POINTER(DeepJob)]
        _lib = L
    return _lib


def _img(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    assert a.ndim == 3 and 1 <= a.shape[2] <= 4
    return a


def resize_dimensions(w, h, nw, nh, fill):
    ow, oh = C.c_uint32(), C.c_uint32()
    lib().fo_resize_dimensions(w, h, nw, nh, int(bool(fill)), C.byref(ow), C.byref(oh))
    return ow.value, oh.value


def resize(img, nw, nh, kind=LANCZOS3):
    a = _img(img)
    h, w, c = a.shape
    out = np.empty((nh, nw, c), np.uint8)
    rc = lib().fo_resize(a.ctypes.data, w, h, c, nw, nh, kind, out.ctypes.data)
    if rc:
        raise ValueError(f"fo_resize rc={rc}")
    return out


def blur(img, sigma):
    a = _img(img)
    h, w, c = a.shape
    out = np.empty_like(a)
    rc = lib().fo_blur(a.ctypes.data, w, h, c, float(sigma), out.ctypes.data)
    if rc:
        raise ValueError(f"fo_blur rc={rc}")
    return out


def grayscale(img):
    a = _img(img)
    h, w, c = a.shape
    out = np.empty((h, w, c), np.uint8)
    oc = lib().fo_grayscale(a.ctypes.data, h * w, c, out.ctypes.data)
    return np.ascontiguousarray(out.reshape(-1)[: h * w * oc].reshape(h, w, oc))


def invert(img):
    a = _img(img).copy()
    h, w, c = a.shape
    lib().fo_invert(a.ctypes.data, h * w, c)
    return a


def to_rgba8(img):
    a = _img(img)
    h, w, c = a.shape
    out = np.empty((h, w, 4), np.uint8)
    lib().fo_to_rgba8(a.ctypes.data, h * w, c, out.ctypes.data)
    return out


def overlay(bottom_rgba, top, x, y):
    b = _img(bottom_rgba).copy()
    t = _img(top)
    assert b.shape[2] == 4
    lib().fo_overlay(b.ctypes.data, b.shape[1], b.shape[0], t.ctypes.data, t.shape[1], t.shape[0], t.shape[2], x, y)
    return b


def weight_table(kind, n_in, n_out, sigma=0.0):
    """(lefts[n_out], counts[n_out], weights[n_out, max_taps]) of one axis."""
    mx = lib().fo_weight_table(kind, sigma, n_in, n_out, None, None, None, 0)
    lefts = np.zeros(n_out, np.uint32)
    counts = np.zeros(n_out, np.uint32)
    ws = np.zeros((n_out, mx), np.float32)
    lib().fo_weight_table(kind, sigma, n_in, n_out, lefts.ctypes.data, counts.ctypes.data, ws.ctypes.data, mx)
    return lefts, counts, ws


def make_job(img, *, w=None, h=None, rgb=(32, 32, 32), crop=False, blur=0.0, grayscale=False,
             inverse=False, gif=False, orientation=1, to_rgb8=False):
    """Job from the accessor values of query::Query (src/query.rs:28-70)."""
    a = _img(img)
    j = Job()
    j.src = a.ctypes.data
    j.src_h, j.src_w, j.src_c = a.shape
    fl = 0
    if grayscale:
        fl |= GRAYSCALE
    if inverse:
        fl |= INVERSE
    if w is not None and h is not None:
        fl |= HAS_DIMS
        j.req_w, j.req_h = int(w), int(h)
    if crop:
        fl |= CROP
    if gif:
        fl |= GIF_FRAME
    if to_rgb8:
        fl |= TO_RGB8
    j.flags = fl
    j.fill[0], j.fill[1], j.fill[2] = rgb
    j.blur_sigma = float(blur)
    j.orientation = int(orientation)
    j._keep = a
    return j


def apply_orientation(img, exif: int) -> np.ndarray:
    """DynamicImage::apply_orientation for an EXIF orientation value (handler.rs:221-223)."""
    a = _img(img)
    h, w, c = a.shape
    out = np.empty(a.size, np.uint8)
    ow, oh = C.c_uint32(), C.c_uint32()
    rc = lib().fo_apply_orientation(C.c_void_p(a.ctypes.data), w, h, c, int(exif), C.c_void_p(out.ctypes.data), C.byref(ow), C.byref(oh))
    if rc:
        raise ValueError(f"fo_apply_orientation rc={rc}")
    return out.reshape(oh.value, ow.value, c)


def ycck_to_cmyk(raw) -> np.ndarray:
    """The YCCK -> CMYK loop of convert_jpeg_color_if_needed (src/handler.rs:420-439) on a flat u8 buffer of 4-byte pixels."""
    a = np.ascontiguousarray(raw, dtype=np.uint8).copy()
    lib().fo_ycck_to_cmyk(a.ctypes.data, a.size)
    return a


def out_capacity(j: Job) -> int:
    """Upper bound of the output size of a job (canvas or source, RGBA)."""
    if j.flags & HAS_DIMS:
        return max(j.req_w * j.req_h, 1) * 4
    return j.src_w * j.src_h * 4


def process(img, **kw) -> np.ndarray:
    """One image through the stage; returns (H, W, C) u8."""
    j = make_job(img, **kw)
    cap = max(out_capacity(j), j.src_w * j.src_h * 4)
    buf = np.empty(cap, np.uint8)
    j.dst = buf.ctypes.data
    j.dst_cap = cap
    rc = lib().fo_process(C.byref(j))
    if rc:
        raise ValueError(f"fo_process rc={rc}")
    n = j.out_w * j.out_h * j.out_c
    return buf[:n].reshape(j.out_h, j.out_w, j.out_c).copy()


def process_batch(imgs, n_threads=1, **kw):
    """Many images, same parameters, one image per thread. Returns list of arrays."""
    n = len(imgs)
    jobs = (Job * n)()
    keep, bufs = [], []
    for i, im in enumerate(imgs):
        j = make_job(im, **kw)
        keep.append(j._keep)
        cap = max(out_capacity(j), j.src_w * j.src_h * 4)
        b = np.empty(cap, np.uint8)
        bufs.append(b)
        j.dst = b.ctypes.data
        j.dst_cap = cap
        C.memmove(C.byref(jobs, i * C.sizeof(Job)), C.byref(j), C.sizeof(Job))
    rc = lib().fo_process_batch(jobs, n, n_threads)
    if rc:
        raise ValueError(f"fo_process_batch rc={rc}")
    outs = []
    for i in range(n):
        j = jobs[i]
        outs.append(bufs[i][: j.out_w * j.out_h * j.out_c].reshape(j.out_h, j.out_w, j.out_c))
    return outs


def process_deep(img, *, w=None, h=None, rgb=(32, 32, 32), crop=False, blur=0.0, grayscale=False, inverse=False, gif=False,
                 orientation=1, to_rgb8=False, to_rgba8=False, to_ycbcr=False) -> np.ndarray:
    """One image of any subpixel type (u8 / u16 / f32 array, (H, W) or (H, W, C)) through the stage as restated in
    fanlin_oracle_deep.c; returns (H, W, C) in the subpixel type the reference would hold (u8 behind a letterbox or
    to_rgb8 / to_rgba8)."""
    a = np.ascontiguousarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    j = DeepJob()
    j.src = a.ctypes.data
    j.src_h, j.src_w, j.src_c = a.shape
    j.sample = sample_kind(a.dtype)
    fl = (GRAYSCALE if grayscale else 0) | (INVERSE if inverse else 0) | (CROP if crop else 0) | (GIF_FRAME if gif else 0)
    fl |= (TO_RGB8 if to_rgb8 else 0) | (TO_RGBA8 if to_rgba8 else 0) | (TO_YCBCR if to_ycbcr else 0)
    if w is not None and h is not None:
        fl |= HAS_DIMS
        j.req_w, j.req_h = int(w), int(h)
    j.flags = fl
    j.fill[0], j.fill[1], j.fill[2] = rgb
    j.blur_sigma = float(blur)
    j.orientation = int(orientation)
    cap = max(j.req_w * j.req_h if fl & HAS_DIMS else 0, j.src_w * j.src_h, 1) * 4 * 4
    buf = np.empty(cap, np.uint8)
    j.dst, j.dst_cap = buf.ctypes.data, cap
    rc = lib().fod_process(C.byref(j))
    if rc:
        raise ValueError(f"fod_process rc={rc}")
    dt = SAMPLE_DTYPES[j.out_sample]
    n = j.out_w * j.out_h * j.out_c
    if to_ycbcr:  # planar: (3, H, W)
        return buf[:n].reshape(3, j.out_h, j.out_w).copy()
    return buf[: n * np.dtype(dt).itemsize].view(dt).reshape(j.out_h, j.out_w, j.out_c).copy()
