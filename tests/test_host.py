"""CPU-side tests of the product's host logic: the C-ABI library loads and exports
every symbol include/fanlin_device.h declares, the query::Query mirror follows the
reference's own test table (src/query.rs:100-405), and the planner reproduces the
oracle's output geometry.  No compute calls: there is no GPU here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(fanlin):
    hdr = open(os.path.join(ROOT, "include", "fanlin_device.h")).read()
    names = re.findall(r"FANLIN_API [^;(]*?\b(fanlin_\w+)\(", hdr)
    assert len(names) >= 20
    L = C.CDLL(fanlin.lib_path())
    for n in names:
        assert hasattr(L, n), n
    assert L.fanlin_abi_version() == 2


def test_no_device_is_an_error_not_a_fallback(fanlin):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fanlin.FanlinError) as ei:
        fanlin.Device([0])
    assert ei.value.status == 5  # FANLIN_ENODEVICE


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fanlin-rs_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dp, fn)


# ---- query::Query mirror: the reference's table, src/query.rs:108-381 -------------

U = "http://127.0.0.1:3000"
QUERY_CASES = [
    (U, False, {}, dict(dimensions=None, fill_color=(32, 32, 32), quality=75, cropping=False, blur=0.0, grayscale=False,
                        inverse=False, use_avif=False, use_webp=False, as_is=True, unsupported_scale_size=False)),
    (U + "?w=", True, {}, {}),
    (U + "?unknown=1", False, {}, {}),
    (U + "?w=2000&h=1000", False, dict(w=2000, h=1000), dict(dimensions=(2000, 1000), as_is=False, unsupported_scale_size=False)),
    (U + "?w=1618", False, dict(w=1618), dict(dimensions=None, as_is=True, unsupported_scale_size=False)),
    (U + "?w=2001&h=1001", False, dict(w=2001, h=1001), dict(dimensions=(2001, 1001), as_is=False, unsupported_scale_size=True)),
    (U + "?w=foo&h=bar", True, {}, {}),
    (U + "?rgb=255,255,255", False, dict(rgb="255,255,255"), dict(fill_color=(255, 255, 255), as_is=True)),
    (U + "?rgb=255,255,255,255", False, dict(rgb="255,255,255,255"), dict(fill_color=(255, 255, 255), as_is=True)),
    (U + "?rgb=255,255", False, dict(rgb="255,255"), dict(fill_color=(32, 32, 32), as_is=True)),
    (U + "?rgb=foo,bar,baz", False, dict(rgb="foo,bar,baz"), dict(fill_color=(32, 32, 32), as_is=True)),
    (U + "?quality=50", False, dict(quality=50), dict(quality=50, as_is=True)),
    (U + "?quality=foo", True, {}, {}),
    (U + "?crop=true", False, dict(crop=True), dict(cropping=True, as_is=True)),
    (U + "?crop=foo", True, {}, {}),
    (U + "?blur=10", False, dict(blur=10), dict(blur=10.0, as_is=False)),
    (U + "?blur=foo", True, {}, {}),
    (U + "?grayscale=true", False, dict(grayscale=True), dict(grayscale=True, as_is=False)),
    (U + "?grayscale=foo", True, {}, {}),
    (U + "?inverse=true", False, dict(inverse=True), dict(inverse=True, as_is=False)),
    (U + "?inverse=foo", True, {}, {}),
    (U + "?avif=true", False, dict(avif=True), dict(use_avif=True, as_is=False)),
    (U + "?avif=foo", True, {}, {}),
    (U + "?webp=true", False, dict(webp=True), dict(use_webp=True, as_is=False)),
    (U + "?webp=foo", True, {}, {}),
]


@pytest.mark.parametrize("uri,error,want,asserts", QUERY_CASES, ids=[c[0][len(U):] or "none" for c in QUERY_CASES])
def test_query(fanlin, uri, error, want, asserts):
    got, err = fanlin.Query.try_from_uri(uri)
    if error:
        assert got is None and err is not None, uri
        return
    assert err is None, uri
    assert got.fields() == want
    for name, val in asserts.items():
        assert getattr(got, name)() == val, (uri, name)


def test_query_blur_clamps_and_rgb_quirks(fanlin):
    Q = fanlin.Query
    assert Q("blur=1").blur() == 10.0 and Q("blur=15").blur() == 15.0 and Q("blur=200").blur() == 20.0  # query.rs:59-62
    assert Q.try_from_uri("?blur=256")[0] is None  # u8 overflow is a parse error
    assert Q("rgb=1,foo,3").fill_color() == (1, 32, 3)
    assert Q("rgb=256,0,0").fill_color() == (32, 0, 0)
    assert Q("rgb=10%2C20%2C30").fill_color() == (10, 20, 30)


def test_query_deserialisation_edge_cases(fanlin):
    """What axum's Query<T> extractor (serde_urlencoded over form_urlencoded; Rust's FromStr for u32 / u8 / bool) does
    with inputs the reference's table (src/query.rs:108-381) does not hold: '+' is a space, only '&' separates pairs, a
    repeated field is an error, unknown and empty keys are ignored, bools are exactly "true" / "false", integers are
    plain decimal digits within the type's range."""
    Q = fanlin.Query
    bad = ["w=+5&h=7", "w=-1", "w=%205", "crop=TRUE", "crop=1", "w=5&w=6", "blur=300", "quality=256", "w=4294967296", "w", "w&h=5",
           "w=5;h=6", "rgb=1,2,3&rgb=4,5,6", "crop=true&crop=false", "w=1e3", "w=0x10", "blur=+12", "blur=12.0", "crop=true+", "w="]
    for qs in bad:
        assert Q.try_from_uri("?" + qs)[0] is None, qs
    ok = {
        "w=05&h=7": dict(dimensions=(5, 7)),
        "w=%35&h=1": dict(dimensions=(5, 1)),
        "w=4294967295&h=1": dict(dimensions=(4294967295, 1)),
        "&&w=5&&h=6&": dict(dimensions=(5, 6)),
        "=5": dict(dimensions=None, as_is=True),
        "foo=bar&w=3&h=4": dict(dimensions=(3, 4)),
        "rgb=1,2": dict(fill_color=(32, 32, 32)),      # fewer than three parts: the default (query.rs:44-46)
        "rgb=1,2,3,4": dict(fill_color=(1, 2, 3)),     # take(3)
        "rgb=+1,2,3": dict(fill_color=(32, 2, 3)),     # " 1" is not a u8
        "rgb=": dict(fill_color=(32, 32, 32)),
        "rgb=1,,3": dict(fill_color=(1, 32, 3)),
        "rgb=%31,2,3": dict(fill_color=(1, 2, 3)),
        "quality=0": dict(quality=0, as_is=True),
        "blur=0": dict(blur=10.0, as_is=False),        # Some(0) clamps to 10: blur=0 is NOT "off" (query.rs:59-62)
        "grayscale=false&inverse=false": dict(grayscale=False, inverse=False, as_is=True),
        "avif=true": dict(use_avif=True, as_is=False),
    }
    for qs, want in ok.items():
        q = Q(qs)
        for name, val in want.items():
            assert getattr(q, name)() == val, (qs, name)


# ---- planner vs oracle geometry -------------------------------------------------------

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]


def _qs(p):
    parts = []
    for k in ("w", "h"):
        if k in p:
            parts.append(f"{k}={p[k]}")
    if "rgb" in p:
        parts.append("rgb=" + ",".join(map(str, p["rgb"])))
    if p.get("crop"):
        parts.append("crop=true")
    if p.get("blur"):
        parts.append(f"blur={int(p['blur'])}")
    if p.get("grayscale"):
        parts.append("grayscale=true")
    if p.get("inverse"):
        parts.append("inverse=true")
    return "&".join(parts)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_plan_matches_oracle_geometry(fanlin, case):
    if case["input"] == "lenna":
        h, w, c = 512, 512, 3
    else:
        _, h, w, c = case["input"]
    img = np.zeros((h, w, c), np.uint8)
    j = fanlin.make_job(img, fanlin.Query(_qs(case["params"])), gif=case["params"].get("gif", False))
    p = fanlin.plan_job(j)
    assert (p.out_h, p.out_w, p.out_channels) == (case["out_h"], case["out_w"], case["out_c"])
    assert p.out_bytes == case["out_h"] * case["out_w"] * case["out_c"]


@pytest.mark.parametrize("src,c,qs,alg", [
    ((1920, 1080), 3, "w=300&h=200", 6460800),            # C2
    ((512, 512), 3, "w=300&h=200&rgb=32,32,32", 1026432),  # C1
    ((3840, 2160), 4, "w=1618&h=1000&crop=true&blur=10", 36763840),  # C3: cols 167..3672
    ((4000, 3000), 3, "w=1618&h=1000&crop=true&grayscale=true&blur=10", 31426000),  # C5 crop
    ((4000, 3000), 3, "w=1618&h=1000&grayscale=true&blur=10", 42472000),  # C5 fit
])
def test_algorithmic_bytes_match_survey(fanlin, src, c, qs, alg):
    j = fanlin.Job()
    C.memset(C.byref(j), 0, C.sizeof(j))
    q = fanlin.Query(qs)
    fanlin.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(j))
    j.src_w, j.src_h, j.src_channels = src[0], src[1], c
    assert fanlin.plan_job(j).algorithmic_bytes == alg


def test_plan_rejects_bad_jobs(fanlin):
    j = fanlin.Job()
    j.src_w, j.src_h, j.src_channels = 0, 10, 3
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)
    j.src_w, j.src_channels = 10, 5
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)


def _job(fanlin, w, h, c, qs, gif=False):
    j = fanlin.Job()
    q = fanlin.Query(qs)
    fanlin.lib().fanlin_job_from_query(C.byref(q._q), int(gif), C.byref(j))
    j.src_w, j.src_h, j.src_channels = w, h, c
    return j


@pytest.mark.parametrize("exif", range(0, 9))
def test_plan_sees_the_oriented_image(fanlin, exif):
    """EXIF 5..8 swap width and height before anything else (apply_orientation, handler.rs:221-223):
    the plan of a stored 400x300 image with orientation e equals the plan of the turned image."""
    j = _job(fanlin, 400, 300, 3, "w=100&h=100&rgb=1,2,3")
    j.orientation = exif
    p = fanlin.plan_job(j)
    t = _job(fanlin, 300 if exif >= 5 else 400, 400 if exif >= 5 else 300, 3, "w=100&h=100&rgb=1,2,3")
    q = fanlin.plan_job(t)
    for f in ("out_w", "out_h", "out_channels", "resized_w", "resized_h", "overlay_x", "overlay_y", "out_bytes", "algorithmic_bytes"):
        assert getattr(p, f) == getattr(q, f), (exif, f)
    assert (p.resized_w, p.resized_h) == ((75, 100) if exif >= 5 else (100, 75))
    j.orientation = 9
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)


def test_plan_output_layouts(fanlin):
    """FANLIN_TO_RGB8 (JPEG branch, handler.rs:274-278) and FANLIN_TO_RGBA8 (WebP branch :287, GIF frames :355)."""
    TO_RGBA8, TO_RGB8 = 1 << 4, 1 << 5
    for c, qs, native in [(1, "w=50&h=50&crop=true", 1), (4, "w=50&h=50&crop=true", 4), (3, "w=50&h=40", 4), (3, "w=50&h=50&crop=true", 3)]:
        j = _job(fanlin, 200, 200, c, qs)
        assert fanlin.plan_job(j).out_channels == native
        j.flags |= TO_RGB8
        p = fanlin.plan_job(j)
        assert p.out_channels == 3 and p.out_bytes == p.out_w * p.out_h * 3
        j.flags |= TO_RGBA8
        with pytest.raises(fanlin.FanlinError):
            fanlin.plan_job(j)
        j.flags &= ~TO_RGB8
        assert fanlin.plan_job(j).out_channels == 4


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (no GPU needed): exactly one line on stdout, valid JSON, with the keys the
    driver reads; whatever libraries print goes to stderr."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_output_layout_flags_in_the_plan(fanlin):
    """fanlin_plan_job (pure host): FANLIN_TO_RGB8 / _TO_RGBA8 / _TO_YCBCR describe the layout the caller gets and exclude
    each other; the source window of a crop request is what fanlin_run copies."""
    img = synth_image(5, 60, 90, 4)
    q = fanlin.Query("w=40&h=40")
    plain = fanlin.plan_job(fanlin.make_job(img, q))
    assert (plain.out_w, plain.out_h, plain.out_channels, plain.out_sample) == (40, 40, 4, 0) and plain.out_bytes == 40 * 40 * 4
    rgb = fanlin.plan_job(fanlin.make_job(img, q, to_rgb8=True))
    assert rgb.out_channels == 3 and rgb.out_bytes == 40 * 40 * 3 and rgb.stages & 32
    ycc = fanlin.plan_job(fanlin.make_job(img, q, to_ycbcr=True))
    assert ycc.out_channels == 3 and ycc.out_bytes == 40 * 40 * 3 and ycc.stages & 64 and not ycc.stages & 32
    gray = fanlin.plan_job(fanlin.make_job(synth_image(6, 60, 90, 3), fanlin.Query("w=40&h=30&crop=true&grayscale=true"), to_rgba8=True))
    assert gray.out_channels == 4 and gray.stages & 16
    for a, b in (("to_rgb8", "to_rgba8"), ("to_rgb8", "to_ycbcr"), ("to_rgba8", "to_ycbcr")):
        with pytest.raises(fanlin.FanlinError):
            fanlin.plan_job(fanlin.make_job(img, q, **{a: True, b: True}))
    tall = fanlin.plan_job(fanlin.make_job(synth_image(7, 900, 300, 3), fanlin.Query("w=200&h=100&crop=true")))
    assert 0 < tall.src_y0 < tall.src_y1 < 900 and (tall.src_x0, tall.src_x1) == (0, 300)
    assert tall.algorithmic_bytes == (tall.src_y1 - tall.src_y0) * 300 * 3 + tall.out_bytes


# ---- the Rust binding on paper must agree with the header and the Makefile (nobody can compile it here) ----

def _c_struct_fields(hdr, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        first, *rest = decl.split(",")
        out.append(re.search(r"(\w+)(\[\d+\])?$", first.strip()).group(1))
        out += [re.search(r"(\w+)", r).group(1) for r in rest]
    return out


def _rust_struct_fields(md, name):
    body = re.search(r"pub struct %s \{(.*?)\}\n" % name, md, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    return re.findall(r"pub (\w+)\s*:", body)


def test_integration_md_agrees_with_header_and_makefile():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    hdr = open(os.path.join(ROOT, "include", "fanlin_device.h")).read()
    mk = open(os.path.join(ROOT, "fanlin-rs_b200", "csrc", "Makefile")).read()
    # build.rs compiles exactly the Makefile's sources
    srcs = re.search(r"^SRCS := (.*)$", mk, re.M).group(1).split()
    rs = re.findall(r'"(\w+\.(?:cpp|cu))"', re.search(r"let src = \[(.*?)\];", md, re.S).group(1))
    assert sorted(rs) == sorted(srcs)
    # every struct crosses the boundary with the header's fields in the header's order
    for c_name, rust_name in (("fanlin_job", "FanlinJob"), ("fanlin_plan", "FanlinPlan"), ("fanlin_config", "FanlinConfig")):
        assert _rust_struct_fields(md, rust_name) == _c_struct_fields(hdr, c_name), rust_name
    # every extern "C" fn of the binding is declared by the header
    declared = set(re.findall(r"FANLIN_API [^;(]*?\b(fanlin_\w+)\(", hdr))
    bound = re.findall(r"^\s*fn (fanlin_\w+)\(", md, re.M)
    assert len(bound) >= 8 and set(bound) <= declared
    # the flag / enum constants carry the header's values
    enums = dict((k, int(v, 0)) for k, v in re.findall(r"FANLIN_(\w+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", hdr))
    enums.update((k, 1 << int(v)) for k, v in re.findall(r"FANLIN_(\w+)\s*=\s*1u?\s*<<\s*(\d+)", hdr))
    for k, v in re.findall(r"pub const (\w+): u32 = (\d+);", md):
        assert enums.get(k, enums.get("FILTER_" + k)) == int(v), (k, v)


def test_planner_fuzz_against_the_oracle(fanlin):
    """400 random requests (every flag, EXIF value, channel count, up- and downscales, requests equal to the image in one
    or both axes): the plan's output dims / channels are the oracle's, and the plan is consistent with itself -- the crop
    rectangle lies inside the resized image, the overlay inside the canvas, the source window inside the oriented image,
    out_bytes and algorithmic_bytes follow from the rest.  Host code only: the CUDA library plans without a device."""
    rng = np.random.default_rng(2024)
    TO_RGBA8, TO_RGB8 = 1 << 4, 1 << 5
    for it in range(400):
        c = int(rng.integers(1, 5))
        w, h = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        gif = bool(rng.random() < 0.15)
        if gif:
            c = 4  # process_gif composites every frame to RGBA8 first (handler.rs:325-328)
        kw = {}
        parts = []
        r = rng.random()
        if r < 0.85:
            rw = w if rng.random() < 0.1 else int(rng.integers(1, 80))
            rh = h if rng.random() < 0.1 else int(rng.integers(1, 80))
            kw.update(w=rw, h=rh)
            parts += [f"w={rw}", f"h={rh}"]
        if rng.random() < 0.4:
            kw["crop"] = True
            parts.append("crop=true")
        if rng.random() < 0.3:
            kw["grayscale"] = True
            parts.append("grayscale=true")
        if rng.random() < 0.3:
            kw["inverse"] = True
            parts.append("inverse=true")
        if rng.random() < 0.3:
            rgb = tuple(int(x) for x in rng.integers(0, 256, 3))
            kw["rgb"] = rgb
            parts.append("rgb=%d,%d,%d" % rgb)
        if not gif and rng.random() < 0.25:
            kw["blur"] = float(rng.integers(10, 21))
            parts.append("blur=%d" % kw["blur"])
        exif = 0 if gif else int(rng.integers(0, 9))
        to_rgb8 = bool(not gif and rng.random() < 0.2)
        j = _job(fanlin, w, h, c, "&".join(parts), gif=gif)
        j.orientation = exif
        if to_rgb8:
            j.flags |= TO_RGB8
        p = fanlin.plan_job(j)
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        want = O.process(img, gif=gif, orientation=max(exif, 1), to_rgb8=to_rgb8, **kw)
        tag = (it, w, h, c, parts, exif, gif, to_rgb8)
        assert (p.out_h, p.out_w, p.out_channels) == want.shape, tag
        assert p.out_bytes == want.size and p.out_sample == 0, tag
        ow, oh = (h, w) if exif >= 5 else (w, h)  # the oriented image
        assert p.src_x0 <= p.src_x1 <= ow and p.src_y0 <= p.src_y1 <= oh, tag
        assert p.algorithmic_bytes == (p.src_x1 - p.src_x0) * (p.src_y1 - p.src_y0) * c + p.out_bytes, tag
        rw, rh = p.resized_w or ow, p.resized_h or oh  # the image behind the resample (resized_* = 0: none happened)
        if p.stages & 4:  # letterbox: resize() fits inside the request, the result sits centred on the canvas (handler.rs:244-245)
            assert rw <= p.out_w and rh <= p.out_h and (rw < p.out_w or rh < p.out_h), tag
            assert (p.overlay_x, p.overlay_y) == ((p.out_w - rw) // 2, (p.out_h - rh) // 2) and (p.crop_x, p.crop_y) == (0, 0), tag
        else:  # the output is the (cropped) image: the crop rectangle lies inside it
            assert p.crop_x + p.out_w <= rw and p.crop_y + p.out_h <= rh, tag
            if not kw.get("crop"):
                assert (p.crop_x, p.crop_y) == (0, 0) and (p.out_w, p.out_h) == (rw, rh), tag
        assert bool(p.stages & 8) == bool(kw.get("blur")), tag
        assert (not (p.stages & 32) or to_rgb8) and (not to_rgb8 or p.out_channels == 3), tag  # bit 5 only where a conversion is needed


def test_header_is_plain_c_and_the_ctypes_mirror_has_its_layout(fanlin, tmp_path):
    """include/fanlin_device.h is the boundary a cgo / Rust / ctypes binding reads: it must compile as C99 (and as C++)
    on its own, and the ctypes structures of fanlin-rs_b200/device.py -- what every test and the bench call through -- must
    have the size and the field offsets the C compiler gives the header's structs."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    pairs = [("fanlin_job", fanlin.device.Job), ("fanlin_plan", fanlin.device.Plan), ("fanlin_config", fanlin.device.Config),
             ("fanlin_stats", fanlin.device.Stats), ("fanlin_query", fanlin.device.QueryStruct)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fanlin_device.h"', "int main(void) {"]
    for c_name, cls in pairs:
        lines.append(f'  printf("{c_name} %zu\\n", sizeof({c_name}));')
        for f, _ in cls._fields_:
            lines.append(f'  printf("{c_name}.{f} %zu\\n", offsetof({c_name}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for c_name, cls in pairs:
        assert int(got[c_name]) == C.sizeof(cls), c_name
        for f, _ in cls._fields_:
            assert int(got[f"{c_name}.{f}"]) == getattr(cls, f).offset, (c_name, f)
    # and as C++ (the header wraps its declarations in extern "C")
    gxx = shutil.which("g++")
    if gxx:
        cpp = tmp_path / "hdr.cpp"
        cpp.write_text('#include "fanlin_device.h"\nint main() { return sizeof(fanlin_job) ? 0 : 1; }\n')
        subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(cpp)], check=True)
