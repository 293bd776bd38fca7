"""Pins the CPU oracle (oracle/fanlin_oracle.c).  The reference holds no pixel
golden vectors for this path (src/main.rs:457-468 asserts status and
Content-Type only) and cannot be built here, so the oracle is "parity
unpinned"; what pins it instead is listed in SURVEY.md 8c and checked here."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import np_restatement as N
from oracle import oracle as O
from synth import synth_image

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def _input(spec, lenna):
    if spec == "lenna":
        return lenna
    seed, h, w, c = spec
    return synth_image(seed, h, w, c)


def _okw(params):
    return {k: (tuple(v) if k == "rgb" else v) for k, v in params.items()}


# image-0.25.6 math/utils.rs upstream unit tests (recalled; the crate is not vendored):
# resize_handles_fill, resize_never_rounds_to_zero, resize_handles_overflow, resize_rounds
@pytest.mark.parametrize("args,want", [
    ((100, 200, 200, 500, True), (250, 500)),
    ((200, 100, 500, 200, True), (500, 250)),
    ((1, 150, 128, 128, False), (1, 128)),
    ((150, 1, 128, 128, False), (128, 1)),
    ((100, 2**32 - 1, 200, 2**32 - 1, True), (100, 2**32 - 1)),
    ((2**32 - 1, 100, 2**32 - 1, 200, True), (2**32 - 1, 100)),
    ((4264, 2476, 3840, 2160, True), (3840, 2230)),
    ((2476, 4264, 2160, 3840, False), (2160, 3720)),
])
def test_resize_dimensions_upstream_kats(args, want):
    assert O.resize_dimensions(*args) == want
    assert N.resize_dimensions(*args) == want


# SURVEY.md section 8 shape table (derived with the A.1 formula)
@pytest.mark.parametrize("src,req,fill,want", [
    ((512, 512), (300, 200), False, (200, 200)),
    ((1920, 1080), (300, 200), False, (300, 169)),
    ((3840, 2160), (1618, 1000), True, (1778, 1000)),
    ((4000, 3000), (1618, 1000), True, (1618, 1214)),
    ((4000, 3000), (1618, 1000), False, (1333, 1000)),
    ((480, 270), (200, 113), False, (200, 113)),
])
def test_resize_dimensions_config_shapes(src, req, fill, want):
    assert O.resize_dimensions(*src, *req, fill) == want


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_golden_table(case, lenna):
    out = O.process(_input(case["input"], lenna), **_okw(case["params"]))
    assert out.shape == (case["out_h"], case["out_w"], case["out_c"])
    assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"]


SMALL = [c for c in GOLDEN["cases"] if c["input"] == "lenna" and c["params"].get("w", 0) <= 600
         or c["input"] != "lenna" and c["input"][1] * c["input"][2] <= 40000]


@pytest.mark.parametrize("case", SMALL, ids=[c["name"] for c in SMALL])
def test_two_restatements_agree_bit_for_bit(case, lenna):
    img = _input(case["input"], lenna)
    p = dict(case["params"])
    kw = dict(w=p.get("w"), h=p.get("h"), rgb=tuple(p.get("rgb", (32, 32, 32))), crop=p.get("crop", False),
              blur_sigma=p.get("blur", 0.0), gray=p.get("grayscale", False), inverse=p.get("inverse", False),
              gif=p.get("gif", False), orientation=p.get("orientation", 1))
    a = O.process(img, **_okw(case["params"]))
    b = N.process(img, **kw)
    assert a.shape == b.shape
    assert np.array_equal(a, b)


def test_weight_tables_agree_and_are_normalised():
    for kind, nm, n_in, n_out, sg in [(O.LANCZOS3, "lanczos3", 1080, 169, 0), (O.LANCZOS3, "lanczos3", 512, 2000, 0),
                                      (O.GAUSSIAN_BLUR, "gaussian", 300, 300, 10.0), (O.NEAREST, "nearest", 480, 200, 0)]:
        lefts, counts, ws = O.weight_table(kind, n_in, n_out, sg)
        tp = N.taps(nm, n_in, n_out, sg)
        for o in range(n_out):
            assert lefts[o] == tp[o][0] and counts[o] == len(tp[o][1])
            assert np.array_equal(ws[o, :counts[o]], tp[o][1])
            assert abs(float(ws[o, :counts[o]].astype(np.float64).sum()) - 1.0) < 1e-5
    # blur sigma=10: 41 taps in the interior, truncated at the borders
    lefts, counts, _ = O.weight_table(O.GAUSSIAN_BLUR, 300, 300, 10.0)
    assert counts[150] == 41 and lefts[150] == 130 and counts[0] == 21
    _, counts, _ = O.weight_table(O.GAUSSIAN_BLUR, 300, 300, 20.0)
    assert counts[150] == 81


def test_pillow_lanczos_structural_check(lenna):
    from PIL import Image

    ours = O.resize(lenna, 200, 200, O.LANCZOS3)
    pil = np.asarray(Image.fromarray(lenna).resize((200, 200), Image.LANCZOS))
    d = np.abs(ours.astype(int) - pil.astype(int))
    assert d.max() <= 2 and (d == 0).mean() > 0.8  # same filter; Pillow is fixed-point with u8 between passes


def test_nearest_is_a_pure_gather():
    img = synth_image(99, 270, 480, 4)
    out = O.resize(img, 200, 113, O.NEAREST)
    ry = np.float32(270) / np.float32(113)
    rx = np.float32(480) / np.float32(200)
    ys = np.minimum(np.floor((np.arange(113, dtype=np.float32) + np.float32(0.5)) * ry).astype(int), 269)
    xs = np.minimum(np.floor((np.arange(200, dtype=np.float32) + np.float32(0.5)) * rx).astype(int), 479)
    assert np.array_equal(out, img[ys][:, xs])


def test_constant_image_stays_constant():
    img = np.full((60, 80, 3), 77, np.uint8)
    assert (O.resize(img, 33, 21, O.LANCZOS3) == 77).all()
    assert (O.blur(img, 10.0) == 77).all()


def test_colour_ops_exact():
    img = synth_image(7, 9, 11, 4)
    g = O.grayscale(img)
    i32 = img.astype(np.uint32)
    l = (2126 * i32[..., 0] + 7152 * i32[..., 1] + 722 * i32[..., 2]) // 10000
    assert g.shape[2] == 2 and np.array_equal(g[..., 0], l) and np.array_equal(g[..., 1], img[..., 3])
    inv = O.invert(img)
    assert np.array_equal(inv[..., :3], 255 - img[..., :3]) and np.array_equal(inv[..., 3], img[..., 3])
    assert np.array_equal(O.invert(O.invert(img)), img)
    assert O.grayscale(img[..., :3]).shape[2] == 1
    assert np.array_equal(O.to_rgba8(g)[..., 0], g[..., 0]) and np.array_equal(O.to_rgba8(g)[..., 3], g[..., 1])


def test_overlay_opaque_is_copy_transparent_keeps_bg_and_never_overflows():
    bg = np.empty((40, 50, 4), np.uint8)
    bg[...] = (32, 32, 32, 255)
    top = synth_image(15, 20, 30, 4)  # seed%8==7: random alpha
    top[0, 0, 3], top[0, 1, 3] = 0, 255
    out = O.overlay(bg, top, 10, 10)
    assert np.array_equal(out[10, 10], bg[10, 10]) and np.array_equal(out[10, 11], top[0, 1])
    assert np.array_equal(out[:10], bg[:10]) and np.array_equal(out[:, :10], bg[:, :10])
    assert np.array_equal(out, N.overlay(bg, top, 10, 10))
    # all 256 alphas x a few colours: the truncating casts stay inside u8 (upstream would panic otherwise)
    a = np.arange(256, dtype=np.uint8)
    for col in (0, 1, 127, 254, 255):
        t = np.stack([np.full(256, col, np.uint8)] * 3 + [a], -1)[None]
        o = O.overlay(np.full((1, 256, 4), 255, np.uint8), t, 0, 0)
        assert o[0, :, 3].min() >= 254


def test_sequencing_semantics(lenna):
    # w alone -> no resize (query.rs:28-33); the oracle sees HAS_DIMS unset
    assert np.array_equal(O.process(lenna, w=300), lenna)
    # grayscale beats inverse (handler.rs:224-228)
    assert np.array_equal(O.process(lenna, grayscale=True, inverse=True), O.process(lenna, grayscale=True))
    # crop never letterboxes; fit letterboxes to RGBA (handler.rs:238)
    assert O.process(lenna, w=300, h=200, crop=True).shape == (200, 300, 3)
    out = O.process(lenna, w=300, h=200, rgb=(1, 2, 3))
    assert out.shape == (200, 300, 4)
    assert (out[:, :50] == (1, 2, 3, 255)).all() and (out[:, 250:] == (1, 2, 3, 255)).all()
    # fill colour is not grayscaled (canvas is built after the colour op)
    out = O.process(lenna, w=300, h=200, rgb=(1, 2, 3), grayscale=True)
    assert (out[:, :50] == (1, 2, 3, 255)).all()
    # GIF frames: nearest, no blur, RGBA out
    fr = synth_image(3, 27, 48, 4)
    assert O.process(fr, gif=True, blur=10.0, grayscale=True).shape == (27, 48, 4)
    assert np.array_equal(O.process(fr, gif=True, blur=10.0), fr)


def test_batch_driver_matches_single():
    imgs = [synth_image(100 + i, 40 + i, 64, 3) for i in range(6)]
    outs = O.process_batch(imgs, n_threads=3, w=30, h=30)
    for im, o in zip(imgs, outs):
        assert np.array_equal(o, O.process(im, w=30, h=30))


# ---- EXIF orientation (handler.rs:206,221-223) ------------------------

@pytest.mark.parametrize("exif", range(0, 9))
@pytest.mark.parametrize("c", [1, 3, 4])
def test_orientation_c_and_numpy_restatements_agree(exif, c):
    img = synth_image(40 + exif, 7, 11, c)
    got = O.apply_orientation(img, exif)
    want = N.apply_orientation(img, exif)
    assert got.shape == want.shape and np.array_equal(got, want)
    if exif >= 5:
        assert got.shape[:2] == (11, 7)


def test_orientation_known_answers():
    """2x3 image with pixel value 10*y + x: the eight EXIF cases spelled out by hand."""
    img = np.array([[0, 1, 2], [10, 11, 12]], np.uint8)[:, :, None]
    want = {
        1: [[0, 1, 2], [10, 11, 12]],
        2: [[2, 1, 0], [12, 11, 10]],          # mirror left-right
        3: [[12, 11, 10], [2, 1, 0]],          # rotate 180
        4: [[10, 11, 12], [0, 1, 2]],          # mirror top-bottom
        5: [[0, 10], [1, 11], [2, 12]],        # transpose
        6: [[10, 0], [11, 1], [12, 2]],        # rotate 90 clockwise
        7: [[12, 2], [11, 1], [10, 0]],        # transverse
        8: [[2, 12], [1, 11], [0, 10]],        # rotate 90 counter-clockwise
    }
    for exif, w in want.items():
        assert O.apply_orientation(img, exif)[:, :, 0].tolist() == w, exif


@pytest.mark.parametrize("exif", [2, 3, 6, 7])
def test_orientation_then_pipeline(exif):
    """The stage sees the oriented image: process(orientation=e) == process(apply_orientation(img, e))."""
    img = synth_image(77, 60, 90, 3)
    kw = dict(w=40, h=30, rgb=(1, 2, 3), grayscale=True)
    a = O.process(img, orientation=exif, **kw)
    b = O.process(np.ascontiguousarray(O.apply_orientation(img, exif)), **kw)
    assert np.array_equal(a, b)
    n = N.process(img, w=40, h=30, rgb=(1, 2, 3), gray=True, orientation=exif)
    assert np.array_equal(a, n)


# ---- encoder layout: RGB8 for the JPEG branch (handler.rs:274-278) ------------

@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_to_rgb8_known_answers_and_restatements(c):
    img = synth_image(60 + c, 9, 13, c)
    a = O.process(img, to_rgb8=True)
    assert a.shape == (9, 13, 3)
    if c <= 2:
        assert all(np.array_equal(a[:, :, k], img[:, :, 0]) for k in range(3))  # luma replicated, alpha dropped
    else:
        assert np.array_equal(a, img[:, :, :3])
    b = O.process(img, w=20, h=20, rgb=(7, 8, 9), to_rgb8=True)  # letterboxed RGBA canvas -> RGB
    full = O.process(img, w=20, h=20, rgb=(7, 8, 9))
    assert np.array_equal(b, full[:, :, :3])
    assert np.array_equal(b, N.process(img, w=20, h=20, rgb=(7, 8, 9), to_rgb8=True))


def test_oracle_matches_reference_golden(lenna):
    """Runs whenever tests/golden/ref_golden.json exists -- the file oracle/pin (image = 0.25.6 through the reference's
    own call sequence, see oracle/pin/README.md) writes on a machine with cargo.  Until then parity stays unpinned
    and this test is skipped, loudly."""
    import hashlib
    import json

    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "tests", "golden", "ref_golden.json")
    if not os.path.exists(path):
        pytest.skip("PARITY UNPINNED: tests/golden/ref_golden.json absent (no Rust toolchain here; recipe in oracle/pin/README.md)")
    ref = json.load(open(path))
    assert ref.get("pinned_to_reference") is True
    ours = {c["name"]: c for c in json.load(open(os.path.join(root, "tests", "golden", "golden.json")))["cases"]}
    from synth import synth_image

    bad = []
    for rc in ref["cases"]:
        c = ours[rc["name"]]
        img = lenna if c["input"] == "lenna" else synth_image(*c["input"])
        kw = {k: (tuple(v) if k == "rgb" else v) for k, v in c["params"].items()}
        got = O.process(img, **kw)
        if got.shape != (rc["out_h"], rc["out_w"], rc["out_c"]) or hashlib.sha256(got.tobytes()).hexdigest() != rc["sha256"]:
            bad.append(rc["name"])
    assert not bad, f"oracle differs from image 0.25.6 on {bad}"
