"""Python face of the query::Query mirror (reference src/query.rs:3-94); all
semantics live in csrc/query.cpp behind the C ABI."""
from __future__ import annotations

import ctypes as C

from .device import FanlinError, QueryStruct, check, lib


class Query:
    def __init__(self, query_string: str = ""):
        self._q = QueryStruct()
        check(lib().fanlin_query_parse(query_string.encode(), C.byref(self._q)))

    @classmethod
    def try_from_uri(cls, uri: str):
        """Ok(Query) / Err like axum::extract::Query::try_from_uri (src/query.rs:383-403)."""
        try:
            return cls(uri), None
        except FanlinError as e:
            return None, e

    def dimensions(self):
        w, h = C.c_uint32(), C.c_uint32()
        if lib().fanlin_query_dimensions(C.byref(self._q), C.byref(w), C.byref(h)):
            return (w.value, h.value)
        return None

    def fill_color(self):
        rgb = (C.c_uint8 * 3)()
        lib().fanlin_query_fill_color(C.byref(self._q), rgb)
        return (rgb[0], rgb[1], rgb[2])

    def quality(self) -> int:
        return self._q.quality if self._q.has_quality else 75

    def cropping(self) -> bool:
        return bool(self._q.has_crop and self._q.crop)

    def blur(self) -> float:
        return float(lib().fanlin_query_blur(C.byref(self._q)))

    def grayscale(self) -> bool:
        return bool(self._q.has_grayscale and self._q.grayscale)

    def inverse(self) -> bool:
        return bool(self._q.has_inverse and self._q.inverse)

    def use_avif(self) -> bool:
        return bool(self._q.has_avif and self._q.avif)

    def use_webp(self) -> bool:
        return bool(self._q.has_webp and self._q.webp)

    def as_is(self) -> bool:
        return bool(lib().fanlin_query_as_is(C.byref(self._q)))

    def unsupported_scale_size(self) -> bool:
        return bool(lib().fanlin_query_unsupported_scale_size(C.byref(self._q)))

    def fields(self) -> dict:
        """Some(..) fields only, for comparison against the reference's `want` structs."""
        q, out = self._q, {}
        for k in ("w", "h", "quality", "blur"):
            if getattr(q, "has_" + k):
                out[k] = getattr(q, k)
        for k in ("crop", "grayscale", "inverse", "avif", "webp"):
            if getattr(q, "has_" + k):
                out[k] = bool(getattr(q, k))
        if q.has_rgb:
            out["rgb"] = q.rgb.decode()
        return out
