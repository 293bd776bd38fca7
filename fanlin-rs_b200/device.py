"""ctypes binding of include/fanlin_device.h -- the same symbols the Rust FFI
crate of INTEGRATION.md binds."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FILTER_NEAREST, FILTER_LANCZOS3 = 0, 1
GRAYSCALE, INVERSE, HAS_DIMS, CROP, TO_RGBA8 = 1, 2, 4, 8, 16
SAMPLE_U8, SAMPLE_U16, SAMPLE_F32 = 0, 1, 2  # enum fanlin_sample
SAMPLE_DTYPES = {SAMPLE_U8: np.uint8, SAMPLE_U16: np.uint16, SAMPLE_F32: np.float32}


class FanlinError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"fanlin status {status}: {msg}")
        self.status = status


class Job(C.Structure):
    _fields_ = [
        ("src", C.c_void_p),
        ("src_w", C.c_uint32), ("src_h", C.c_uint32), ("src_channels", C.c_uint32), ("src_pitch", C.c_uint32),
        ("flags", C.c_uint32), ("filter", C.c_uint32),
        ("req_w", C.c_uint32), ("req_h", C.c_uint32),
        ("fill_rgb", C.c_uint8 * 3), ("orientation", C.c_uint8),
        ("blur_sigma", C.c_float),
        ("dst", C.c_void_p),
        ("dst_capacity", C.c_uint64),
        ("src_sample", C.c_uint32), ("reserved", C.c_uint32),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("out_channels", C.c_uint32),
        ("resized_w", C.c_uint32), ("resized_h", C.c_uint32),
        ("crop_x", C.c_uint32), ("crop_y", C.c_uint32),
        ("overlay_x", C.c_uint32), ("overlay_y", C.c_uint32),
        ("src_x0", C.c_uint32), ("src_y0", C.c_uint32), ("src_x1", C.c_uint32), ("src_y1", C.c_uint32),
        ("stages", C.c_uint32),
        ("out_bytes", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
        ("out_sample", C.c_uint32), ("reserved", C.c_uint32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("exact", C.c_uint32),
        ("device_scratch_bytes", C.c_uint64), ("pinned_bytes", C.c_uint64),
        ("batch_window_us", C.c_uint32), ("max_batch_jobs", C.c_uint32),
        ("vertical_path", C.c_uint32), ("blur_path", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("jobs", C.c_uint64), ("batches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("table_bytes", C.c_uint64)]


class QueryStruct(C.Structure):
    _fields_ = [("has_w", C.c_uint32), ("w", C.c_uint32), ("has_h", C.c_uint32), ("h", C.c_uint32),
                ("has_rgb", C.c_uint32), ("rgb", C.c_char * 64),
                ("has_quality", C.c_uint32), ("quality", C.c_uint32),
                ("has_crop", C.c_uint32), ("crop", C.c_uint32),
                ("has_blur", C.c_uint32), ("blur", C.c_uint32),
                ("has_grayscale", C.c_uint32), ("grayscale", C.c_uint32),
                ("has_inverse", C.c_uint32), ("inverse", C.c_uint32),
                ("has_avif", C.c_uint32), ("avif", C.c_uint32),
                ("has_webp", C.c_uint32), ("webp", C.c_uint32)]


def lib_path() -> str:
    return os.path.join(_HERE, "libfanlin_device.so")


_lib = None


def lib():
    """The C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise FanlinError(-1, f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback for this path)")
        L = C.CDLL(p)
        vp, u32, P = C.c_void_p, C.c_uint32, C.POINTER
        L.fanlin_abi_version.restype = C.c_int
        L.fanlin_plan_job.argtypes = [P(Job), P(Plan)]
        L.fanlin_init.argtypes = [P(C.c_int), C.c_int, P(Config), P(vp)]
        L.fanlin_shutdown.argtypes = [vp]
        L.fanlin_shutdown.restype = None
        L.fanlin_device_count.argtypes = [vp]
        L.fanlin_run.argtypes = [vp, P(Job), u32, P(Plan)]
        L.fanlin_batch_prepare.argtypes = [vp, C.c_int, P(Job), u32, P(Plan), P(vp)]
        L.fanlin_batch_launch.argtypes = [vp, vp]
        L.fanlin_batch_launch_count.argtypes = [vp]
        L.fanlin_batch_free.argtypes = [vp]
        L.fanlin_batch_free.restype = None
        L.fanlin_batch_set_timing.argtypes = [vp, C.c_int]
        L.fanlin_batch_kernel_times.argtypes = [vp, P(C.c_char_p), P(C.c_float), C.c_int]
        L.fanlin_host_alloc.argtypes = [vp, C.c_size_t]
        L.fanlin_host_alloc.restype = vp
        L.fanlin_host_free.argtypes = [vp, vp]
        L.fanlin_host_free.restype = None
        L.fanlin_get_stats.argtypes = [vp, P(Stats)]
        L.fanlin_shard_range.argtypes = [u32, u32, u32, P(u32), P(u32)]
        L.fanlin_shard_range.restype = None
        L.fanlin_ycck_to_cmyk.argtypes = [vp, vp, vp, C.c_uint64]
        L.fanlin_ycck_to_cmyk_device.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, vp]
        L.fanlin_last_error.restype = C.c_char_p
        L.fanlin_query_parse.argtypes = [C.c_char_p, P(QueryStruct)]
        L.fanlin_query_dimensions.argtypes = [P(QueryStruct), P(u32), P(u32)]
        L.fanlin_query_fill_color.argtypes = [P(QueryStruct), C.c_uint8 * 3]
        L.fanlin_query_fill_color.restype = None
        L.fanlin_query_blur.argtypes = [P(QueryStruct)]
        L.fanlin_query_blur.restype = C.c_float
        L.fanlin_query_as_is.argtypes = [P(QueryStruct)]
        L.fanlin_query_unsupported_scale_size.argtypes = [P(QueryStruct)]
        L.fanlin_job_from_query.argtypes = [P(QueryStruct), C.c_int, P(Job)]
        L.fanlin_job_from_query.restype = None
        assert L.fanlin_abi_version() == 2
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise FanlinError(rc, (lib().fanlin_last_error() or b"").decode())


def shard_range(n_jobs: int, n_shards: int, shard: int):
    """[lo, hi) of the images shard `shard` owns (fanlin_shard_range)."""
    lo, hi = C.c_uint32(), C.c_uint32()
    lib().fanlin_shard_range(n_jobs, n_shards, shard, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def plan_job(job: Job) -> Plan:
    p = Plan()
    check(lib().fanlin_plan_job(C.byref(job), C.byref(p)))
    return p


class DeviceBatch:
    """A prepared device-resident batch (fanlin_batch_prepare / _launch / _free)."""

    def __init__(self, dev: "Device", handle, plans, n_jobs):
        self._dev, self._h, self.plans, self.n_jobs = dev, handle, plans, n_jobs

    def launch(self, cuda_stream: int | None = None):
        check(lib().fanlin_batch_launch(self._h, C.c_void_p(cuda_stream or 0)))

    @property
    def launches_per_run(self) -> int:
        return lib().fanlin_batch_launch_count(self._h)

    def set_timing(self, enable: bool = True):
        check(lib().fanlin_batch_set_timing(self._h, int(enable)))

    def kernel_times(self):
        """[(kernel name, ms)] of the launches since the last call; synchronise their stream first."""
        cap = 8192
        names, ms = (C.c_char_p * cap)(), (C.c_float * cap)()
        n = lib().fanlin_batch_kernel_times(self._h, names, ms, cap)
        return [(names[i].decode(), float(ms[i])) for i in range(min(n, cap))]

    def free(self):
        if self._h:
            lib().fanlin_batch_free(self._h)
            self._h = None

    def __del__(self):
        self.free()


class Device:
    """fanlin_ctx: created once at start-up, shared by all request threads."""

    def __init__(self, device_ids=None, *, exact=False, device_scratch_bytes=0, pinned_bytes=0,
                 batch_window_us=0, max_batch_jobs=0, tensor_cores=True, vertical_path=None, blur_path=0):
        """vertical_path: None = from tensor_cores (0 / 1); 2 = tensor cores for the vertical pass only
        (the horizontal stage stays on the CUDA cores); 3 = both passes on the tensor cores whatever the
        batch size (0 takes them from 256 jobs per batch on)."""
        cfg = Config(C.sizeof(Config), int(exact), device_scratch_bytes, pinned_bytes, batch_window_us, max_batch_jobs,
                     (0 if tensor_cores else 1) if vertical_path is None else int(vertical_path), int(blur_path))
        h = C.c_void_p()
        if device_ids:
            arr = (C.c_int * len(device_ids))(*device_ids)
            rc = lib().fanlin_init(arr, len(device_ids), C.byref(cfg), C.byref(h))
        else:
            rc = lib().fanlin_init(None, 0, C.byref(cfg), C.byref(h))
        check(rc)
        self._h = h

    def close(self):
        if self._h:
            lib().fanlin_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return lib().fanlin_device_count(self._h)

    def stats(self) -> dict:
        s = Stats()
        check(lib().fanlin_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def run(self, jobs):
        """Blocking host-buffer path (fanlin_run).  jobs: ctypes array or list of Job."""
        n = len(jobs)
        arr = jobs if isinstance(jobs, C.Array) else (Job * n)(*jobs)
        plans = (Plan * n)()
        check(lib().fanlin_run(self._h, arr, n, plans))
        return plans

    def prepare(self, jobs, device_index=0) -> DeviceBatch:
        n = len(jobs)
        arr = jobs if isinstance(jobs, C.Array) else (Job * n)(*jobs)
        plans = (Plan * n)()
        h = C.c_void_p()
        check(lib().fanlin_batch_prepare(self._h, device_index, arr, n, plans, C.byref(h)))
        return DeviceBatch(self, h, plans, n)

    def ycck_to_cmyk(self, raw: np.ndarray) -> np.ndarray:
        """The YCCK -> CMYK loop of convert_jpeg_color_if_needed (src/handler.rs:420-439) on the device:
        flat u8 buffer of 4-byte pixels in, converted copy out."""
        a = np.ascontiguousarray(raw, dtype=np.uint8)
        out = np.empty_like(a)
        check(lib().fanlin_ycck_to_cmyk(self._h, C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), a.size // 4))
        if a.size % 4:
            out.reshape(-1)[a.size - a.size % 4:] = a.reshape(-1)[a.size - a.size % 4:]
        return out

    def ycck_to_cmyk_device(self, src_ptr: int, dst_ptr: int, n_pixels: int, device_index: int = 0, cuda_stream: int | None = None):
        check(lib().fanlin_ycck_to_cmyk_device(self._h, device_index, C.c_void_p(src_ptr), C.c_void_p(dst_ptr), n_pixels, C.c_void_p(cuda_stream or 0)))

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer from the context's pool as a u8 array (free with host_free)."""
        p = lib().fanlin_host_alloc(self._h, nbytes)
        if not p:
            raise FanlinError(2, "pinned allocation failed")
        buf = (C.c_uint8 * nbytes).from_address(p)
        a = np.frombuffer(buf, dtype=np.uint8)
        a.flags.writeable = True
        return a

    def host_free(self, arr: np.ndarray):
        lib().fanlin_host_free(self._h, C.c_void_p(arr.ctypes.data))
