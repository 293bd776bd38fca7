"""16-bit and f32 DynamicImage variants (reference src/handler.rs:219 yields them for 16-bit PNG /
TIFF and HDR / EXR, and :224-255 runs them through the same generic image-crate code as the u8 variants).

CPU part: the restatement (oracle/fanlin_oracle_deep.c) is held to the u8 oracle (sample kind 0 must reproduce it bit
for bit), to an independent numpy restatement, to analytic invariants and to its committed hashes; the planner's output
description is held to the oracle's.  GPU part (-m gpu): the CUDA path through the C ABI is BIT-EXACT against the
oracle -- these subpixel types take kernels in the crate's operation order (kernels_deep.cu); only a blur behind a
letterbox (an Rgba<u8> canvas) runs the u8 tensor-core blur, within 1 LSB.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import np_restatement_deep as ND
from oracle import oracle as O
from synth import synth_deep, synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_deep.json")))["cases"]

REQUESTS = [
    dict(w=30, h=20), dict(w=30, h=20, crop=True), dict(w=40, h=40, grayscale=True), dict(w=25, h=31, inverse=True, blur=1.5),
    dict(grayscale=True), dict(blur=2.0), dict(w=30, h=20, orientation=6), dict(w=30, h=20, to_rgb8=True),
    dict(w=17, h=40, crop=True, to_rgba8=True), dict(w=64, h=48), dict(w=64, h=10, crop=True, blur=1.2, orientation=3),
    dict(w=70, h=48, rgb=(200, 100, 0)), dict(w=100, h=90, crop=True, inverse=True),
]


def _okw(params):
    return {k: (tuple(v) if k == "rgb" else v) for k, v in params.items()}


def _nkw(kw):
    kn = dict(kw)
    for a, b in (("blur", "blur_sigma"), ("grayscale", "gray"), ("to_rgba8", "to_rgba8_out")):
        if a in kn:
            kn[b] = kn.pop(a)
    return kn


def _dtypes(c):
    return (np.uint16, np.float32) if c >= 3 else (np.uint16,)


# ---- CPU: the oracle ----------------------------------------------------------------------------------

@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_deep_oracle_with_u8_samples_is_the_u8_oracle(c):
    for seed in range(4):
        img = synth_image(seed + 4 * c, 48, 64, c)  # seeds 7, 15: random alpha
        for kw in REQUESTS:
            if "to_rgba8" in kw:
                continue
            a, b = O.process(img, **kw), O.process_deep(img, **kw)
            assert a.shape == b.shape and a.dtype == b.dtype and (a == b).all(), (seed, c, kw)


@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_deep_oracle_matches_numpy_restatement(c):
    for seed in (5, 7):
        for dt in _dtypes(c):
            img = synth_deep(seed, 48, 64, c, dt)
            for kw in REQUESTS:
                a, b = O.process_deep(img, **kw), ND.process(img, **_nkw(kw))
                assert a.shape == b.shape and a.dtype == b.dtype, (c, dt, kw)
                assert a.tobytes() == b.tobytes(), (seed, c, dt, kw)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_deep_oracle_golden_hashes(case):
    seed, h, w, c, dt = case["input"]
    res = O.process_deep(synth_deep(seed, h, w, c, np.dtype(dt)), **_okw(case["params"]))
    assert res.shape == (case["out_h"], case["out_w"], case["out_c"]) and str(res.dtype) == case["out_dtype"]
    assert hashlib.sha256(res.tobytes()).hexdigest() == case["sha256"]


def test_deep_oracle_matches_reference_golden():
    """Runs whenever tests/golden/ref_golden.json exists (oracle/pin: image = 0.25.6 through the reference's own call sequence,
    on a machine with cargo) and holds the 16-bit / f32 restatement to it.  Until then parity stays unpinned."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.json")
    if not os.path.exists(path):
        pytest.skip("PARITY UNPINNED: tests/golden/ref_golden.json absent (no Rust toolchain here; recipe in oracle/pin/README.md)")
    ref = {c["name"]: c for c in json.load(open(path))["cases"]}
    bad = []
    for case in GOLDEN:
        rc = ref.get(case["name"])
        if rc is None:
            bad.append(case["name"] + " (missing)")
            continue
        seed, h, w, c, dt = case["input"]
        got = O.process_deep(synth_deep(seed, h, w, c, np.dtype(dt)), **_okw(case["params"]))
        if got.shape != (rc["out_h"], rc["out_w"], rc["out_c"]) or str(got.dtype) != rc.get("out_dtype", str(got.dtype)) or \
                hashlib.sha256(got.tobytes()).hexdigest() != rc["sha256"]:
            bad.append(case["name"])
    assert not bad, f"the 16-bit / f32 oracle disagrees with image 0.25.6 on: {bad}"


def test_deep_invariants():
    # u16 samples that are 257 x a u8 image: integer colour ops and the u8 view commute with the scaling
    for c in (1, 2, 3, 4):
        a8 = synth_image(40 + c, 21, 33, c)
        a16 = a8.astype(np.uint16) * 257
        assert (O.process_deep(a16, inverse=True) == O.process(a8, inverse=True).astype(np.uint16) * 257).all()
        assert (O.process_deep(a16, to_rgba8=True) == O.process_deep(a8, to_rgba8=True)).all()
    # a constant image stays constant through resize and blur; f32 resize clamps to [0, 1]
    k16 = np.full((30, 40, 3), 51234, np.uint16)
    assert (O.process_deep(k16, w=17, h=11, crop=True, blur=10.0) == 51234).all()
    hdr = np.full((30, 40, 3), 2.5, np.float32)
    assert (O.process_deep(hdr, w=17, h=11, crop=True) == 1.0).all()
    assert (O.process_deep(hdr, inverse=True) == -1.5).all()  # the colour ops do not clamp
    # grayscale keeps the f32 pixel type (luma replicated), narrows the 16-bit ones
    assert O.process_deep(synth_deep(1, 8, 9, 3, np.float32), grayscale=True).shape == (8, 9, 3)
    assert O.process_deep(synth_deep(1, 8, 9, 4, np.uint16), grayscale=True).shape == (8, 9, 2)
    g = O.process_deep(synth_deep(1, 8, 9, 4, np.float32), grayscale=True)
    assert (g[..., 0] == g[..., 1]).all() and (g[..., 1] == g[..., 2]).all()
    # a letterbox makes an Rgba<u8> canvas whatever the image
    o = O.process_deep(synth_deep(2, 20, 50, 3, np.uint16), w=60, h=60)
    assert o.dtype == np.uint8 and o.shape == (60, 60, 4)
    # FromPrimitive<u16> for u8 == round(c * 255 / 65535) on all 65536 values
    allv = np.arange(65536, dtype=np.uint16).reshape(256, 256, 1)
    ref = np.floor(allv.astype(np.float64) * 255.0 / 65535.0 + 0.5).astype(np.uint8)
    assert (O.process_deep(allv, to_rgba8=True)[..., 0:1] == ref).all()


def _job(fanlin, img, qs, **kw):
    return fanlin.make_job(img, fanlin.Query(qs), **kw)


def _qs(p):
    parts = [f"{k}={p[k]}" for k in ("w", "h") if k in p]
    if "rgb" in p:
        parts.append("rgb=" + ",".join(map(str, p["rgb"])))
    for k in ("crop", "grayscale", "inverse"):
        if p.get(k):
            parts.append(f"{k}=true")
    if p.get("blur"):
        parts.append(f"blur={int(p['blur'])}")
    return "&".join(parts)


def test_planner_describes_the_oracle_output(fanlin):
    """fanlin_plan_job (pure host): output size, channels, subpixel type and bytes for 16-bit / f32 requests."""
    for case in GOLDEN:
        seed, h, w, c, dt = case["input"]
        p = case["params"]
        img = synth_deep(seed, h, w, c, np.dtype(dt))
        j = _job(fanlin, img, _qs(p), orientation=p.get("orientation", 1), to_rgb8=p.get("to_rgb8", False), to_rgba8=p.get("to_rgba8", False))
        pl = fanlin.plan_job(j)
        assert (pl.out_h, pl.out_w, pl.out_channels) == (case["out_h"], case["out_w"], case["out_c"]), case["name"]
        assert str(np.dtype(fanlin.device.SAMPLE_DTYPES[pl.out_sample])) == case["out_dtype"], case["name"]
        assert pl.out_bytes == case["out_h"] * case["out_w"] * case["out_c"] * np.dtype(case["out_dtype"]).itemsize
    # rejected: f32 without colour channels, unknown subpixel type, odd pitch for u16
    j = _job(fanlin, synth_deep(1, 8, 9, 3, np.uint16), "w=4&h=4")
    j.src_sample = 3
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)
    j.src_sample, j.src_channels = 2, 2
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)
    j.src_sample, j.src_channels, j.src_pitch = 1, 3, 9 * 3 * 2 + 1
    with pytest.raises(fanlin.FanlinError):
        fanlin.plan_job(j)


# ---- GPU: parity through the C ABI -----------------------------------------------------------------

@pytest.fixture(scope="module")
def dev(fanlin):
    d = fanlin.Device([0])
    yield d
    d.close()


@pytest.fixture(scope="module")
def dev_exact(fanlin):
    d = fanlin.Device([0], exact=True)
    yield d
    d.close()


def _run(fanlin, dev, img, p):
    return fanlin.process_image(dev, img, fanlin.Query(_qs(p)), orientation=p.get("orientation", 1), to_rgb8=p.get("to_rgb8", False),
                                to_rgba8=p.get("to_rgba8", False))


def _u8_blur_behind_letterbox(p, img):
    """True when the request ends in a blur of an Rgba<u8> canvas: the only part of a 16-bit / f32 request that takes a
    u8 fast path (1-LSB bar) on a default context."""
    return p.get("blur", 0) > 0 and O.process_deep(img, **{k: v for k, v in _okw(p).items() if k not in ("blur", "to_rgb8", "to_rgba8")}).dtype == np.uint8 and img.dtype != np.uint8


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_deep_golden_on_device(fanlin, dev, dev_exact, case):
    seed, h, w, c, dt = case["input"]
    p = case["params"]
    img = synth_deep(seed, h, w, c, np.dtype(dt))
    want = O.process_deep(img, **_okw(p))
    assert hashlib.sha256(want.tobytes()).hexdigest() == case["sha256"]  # the box's own build of the oracle
    for d in (dev_exact, dev):
        got = _run(fanlin, d, img, p)
        assert got.shape == want.shape and got.dtype == want.dtype, case["name"]
        if d is dev and _u8_blur_behind_letterbox(p, img):
            assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1, case["name"]
        else:
            assert got.tobytes() == want.tobytes(), (case["name"], "exact" if d is dev_exact else "default")


@pytest.mark.gpu
@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_deep_requests_bit_exact(fanlin, dev, c):
    reqs = [r for r in REQUESTS] + [dict(w=50, h=50, blur=10), dict(w=33, h=21, crop=True, blur=10, grayscale=True, orientation=8),
                                    dict(w=64, h=48, orientation=5, inverse=True), dict(w=20, h=20, orientation=2, to_rgb8=True),
                                    dict(blur=15, orientation=4), dict(orientation=7, grayscale=True)]
    for seed in (3, 7):
        for dt in _dtypes(c):
            img = synth_deep(seed, 48, 64, c, dt)
            for p in reqs:
                p = {k: (int(v) if k == "blur" else v) for k, v in p.items() if not (k == "blur" and v < 10)}
                want = O.process_deep(img, **_okw(p))
                got = _run(fanlin, dev, img, p)
                assert got.shape == want.shape and got.dtype == want.dtype, (c, dt, p)
                if _u8_blur_behind_letterbox(p, img):
                    assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1, (c, dt, p)
                else:
                    assert got.tobytes() == want.tobytes(), (seed, c, dt, p)


@pytest.mark.gpu
def test_deep_and_u8_jobs_share_a_batch(fanlin, dev):
    """One fanlin_run call with u8, u16 and f32 jobs of different geometry (what the batcher merges)."""
    specs = [(synth_image(1, 90, 120, 3), dict(w=50, h=40)), (synth_deep(2, 70, 50, 4, np.uint16), dict(w=40, h=40, crop=True)),
             (synth_deep(3, 64, 64, 3, np.float32), dict(w=30, h=30, grayscale=True)), (synth_image(4, 200, 300, 4), dict(w=100, h=100, crop=True, blur=10)),
             (synth_deep(5, 33, 47, 1, np.uint16), dict(w=60, h=60)), (synth_deep(15, 80, 80, 2, np.uint16), dict(inverse=True))] * 7  # 42 jobs: past the batcher, one ragged batch
    jobs = [_job(fanlin, im, _qs(p)) for im, p in specs]
    outs = fanlin.stage._run(dev, jobs)
    for (im, p), got in zip(specs, outs):
        want = O.process_deep(im, **_okw(p))
        assert got.shape == want.shape and got.dtype == want.dtype
        if im.dtype == np.uint8:
            assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
        else:
            assert got.tobytes() == want.tobytes(), p


@pytest.mark.gpu
def test_deep_device_batch(fanlin, dev):
    """fanlin_batch_prepare / _launch on device-resident 16-bit images (pitched rows), against the oracle."""
    import torch

    imgs = [synth_deep(50 + i, 61, 83, 3, np.uint16) for i in range(5)]
    pitch = (83 * 3 * 2 + 15) // 16 * 16 + 16  # padded rows
    src = torch.zeros((5, 61, pitch), dtype=torch.uint8, device="cuda")
    for i, im in enumerate(imgs):
        src[i, :, : 83 * 6] = torch.from_numpy(im.reshape(61, 83 * 3).view(np.uint8)).cuda()
    q = fanlin.Query("w=40&h=30&crop=true")
    jobs = (fanlin.Job * 5)()
    dst = torch.zeros((5, 30, 40, 3), dtype=torch.int16, device="cuda")
    for i in range(5):
        j = fanlin.make_job(imgs[i], q)
        j.src, j.src_pitch = src[i].data_ptr(), pitch
        j.dst, j.dst_capacity = dst[i].data_ptr(), 30 * 40 * 3 * 2
        C.memmove(C.byref(jobs, i * C.sizeof(fanlin.Job)), C.byref(j), C.sizeof(fanlin.Job))
    torch.cuda.synchronize()
    b = dev.prepare(jobs)
    b.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dst.cpu().numpy().view(np.uint16)
    for i in range(5):
        assert (out[i] == O.process_deep(imgs[i], w=40, h=30, crop=True)).all()
    b.free()


@pytest.mark.gpu
def test_deep_fuzz(fanlin, dev):
    rng = np.random.default_rng(20261018)
    for it in range(60):
        c = int(rng.integers(1, 5))
        dt = np.float32 if (c >= 3 and rng.random() < 0.4) else np.uint16
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 120))
        img = synth_deep(int(rng.integers(0, 1000)), h, w, c, dt)
        p = {}
        if rng.random() < 0.85:
            p["w"], p["h"] = int(rng.integers(1, 140)), int(rng.integers(1, 110))
            if rng.random() < 0.5:
                p["crop"] = True
            if rng.random() < 0.5:
                p["rgb"] = tuple(int(v) for v in rng.integers(0, 256, 3))
        r = rng.random()
        if r < 0.3:
            p["grayscale"] = True
        elif r < 0.5:
            p["inverse"] = True
        if rng.random() < 0.3:
            p["blur"] = int(rng.integers(10, 21))
        if rng.random() < 0.4:
            p["orientation"] = int(rng.integers(1, 9))
        r = rng.random()
        if r < 0.2:
            p["to_rgb8"] = True
        elif r < 0.4:
            p["to_rgba8"] = True
        want = O.process_deep(img, **_okw(p))
        got = _run(fanlin, dev, img, p)
        assert got.shape == want.shape and got.dtype == want.dtype, (it, c, dt, h, w, p)
        if _u8_blur_behind_letterbox(p, img):
            assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1, (it, p)
        else:
            assert got.tobytes() == want.tobytes(), (it, c, dt, h, w, p)


# ---- FANLIN_TO_YCBCR: the JPEG encoder's planes of the result ---------------------------------

def _ycbcr_np(rgb):
    """codecs/jpeg/encoder.rs rgb_to_ycbcr, vectorised in f32 (a second restatement): (..., 3) u8 -> three u8 arrays."""
    f = np.float32
    r, g, b = (rgb[..., k].astype(f) for k in range(3))
    m = f(255.0)
    y = (f(76.245) / m) * r + (f(149.685) / m) * g + (f(29.07) / m) * b
    cb = (f(-43.0185) / m) * r - (f(84.4815) / m) * g + (f(127.5) / m) * b + f(128.0)
    cr = (f(127.5) / m) * r - (f(106.7685) / m) * g - (f(20.7315) / m) * b + f(128.0)
    return tuple(np.clip(np.trunc(v), 0, 255).astype(np.uint8) for v in (y, cb, cr))


def _all_rgb():
    v = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([v & 255, (v >> 8) & 255, v >> 16], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)


def test_ycbcr_oracle_all_triples():
    img = _all_rgb()
    got = O.process_deep(img, to_ycbcr=True)
    assert got.shape == (3, 4096, 4096)
    for plane, want in zip(got, _ycbcr_np(img)):
        assert (plane == want).all()
    # known answers: black, white (f32: 76.245/255*255 + ... rounds just below 255), the primaries of JFIF
    px = np.array([[[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]]], np.uint8)
    y, cb, cr = O.process_deep(px, to_ycbcr=True)
    assert y[0].tolist()[:1] == [0] and cb[0, 0] == 128 and cr[0, 0] == 128
    assert y[0, 2] == 76 and y[0, 3] == 149 and y[0, 4] == 29 and cb[0, 4] == 255 and cr[0, 2] == 255
    # the request runs first, the planes come from to_rgb8() of its result
    a = synth_image(9, 40, 60, 4)
    rgb = O.process_deep(a, w=50, h=50, to_rgb8=True)
    assert (np.stack(_ycbcr_np(rgb)) == O.process_deep(a, w=50, h=50, to_ycbcr=True)).all()


@pytest.mark.gpu
def test_ycbcr_planes_on_device(fanlin, dev, dev_exact):
    img = _all_rgb()  # every (R, G, B): bit-exact
    got = fanlin.process_image(dev, img, fanlin.Query(""), to_ycbcr=True)
    assert got.shape == (3, 4096, 4096) and (got == O.process_deep(img, to_ycbcr=True)).all()
    for (seed, h, w, c, dt), p in [((1, 90, 120, 3, np.uint8), dict(w=64, h=64, rgb=(3, 4, 5))), ((7, 90, 120, 4, np.uint8), dict(w=64, h=40, crop=True)),
                                   ((2, 90, 120, 1, np.uint8), dict(w=50, h=50, crop=True, blur=10)), ((3, 60, 80, 2, np.uint8), dict(inverse=True)),
                                   ((4, 70, 90, 3, np.uint16), dict(w=40, h=30, crop=True, orientation=6)), ((5, 70, 90, 4, np.float32), dict(w=60, h=60))]:
        img = synth_deep(seed, h, w, c, dt)
        want = O.process_deep(img, to_ycbcr=True, **_okw(p))
        e = fanlin.process_image(dev_exact, img, fanlin.Query(_qs(p)), orientation=p.get("orientation", 1), to_ycbcr=True)
        assert e.shape == want.shape and (e == want).all(), p
        g = fanlin.process_image(dev, img, fanlin.Query(_qs(p)), orientation=p.get("orientation", 1), to_ycbcr=True)
        assert g.shape == want.shape and np.abs(g.astype(np.int16) - want.astype(np.int16)).max() <= 2, p  # a 1-LSB RGB difference moves a plane by <= 1 (+ truncation)
