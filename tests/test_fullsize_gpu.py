"""BASELINE.json configs at their STATED sizes through the C ABI against the CPU oracle.

C1 lenna 512x512 RGB -> w=300&h=200&rgb=32,32,32; C2 1920x1080 RGB -> 300x200 fit + fill;
C3 3840x2160 RGBA -> w=1618&h=1000&crop=true&blur=10 (opaque and random-alpha images);
C4 480x270 RGBA GIF frames (in test_parity_gpu.py at full size already); C5 4000x3000 RGB ->
w=1618&h=1000 with grayscale + blur=10, crop and fit + fill sub-configs (SURVEY.md section 8:
crop and fill exclude each other, so "resize+crop+fill+blur+grayscale" is two requests); one
sigma = 20 (81 taps) case at full size.  Each on a default context (what a single request
takes) and on a `vertical_path = 3` context (what the BASELINE batches take: both Lanczos3
passes on the tensor cores where the geometry allows).  Bars: Lanczos3 / blur within 1 LSB,
histogram asserted (>= 2 must be empty); a batch through the device-resident entry point must
give the same bytes as the single request.  Reference call sites: src/handler.rs:224-255.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from synth import synth_image

pytestmark = pytest.mark.gpu


def hist(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return {0: int((d == 0).sum()), 1: int((d == 1).sum()), ">=2": int((d >= 2).sum())}


@pytest.fixture(scope="module", params=[0, 3], ids=["default_ctx", "vertical_path3"])
def dev(fanlin, request):
    d = fanlin.Device([0], vertical_path=request.param)
    yield d
    d.close()


#        name                 seed  h     w     c  query                                                    oracle kwargs
FULL = [
    ("c2_fit_fill",           2001, 1080, 1920, 3, "w=300&h=200",                                           dict(w=300, h=200)),
    ("c3_crop_blur10",          16, 2160, 3840, 4, "w=1618&h=1000&crop=true&blur=10",                       dict(w=1618, h=1000, crop=True, blur=10.0)),
    ("c3_crop_blur10_alpha",    23, 2160, 3840, 4, "w=1618&h=1000&crop=true&blur=10",                       dict(w=1618, h=1000, crop=True, blur=10.0)),
    ("c3_crop_noblur_alpha",    31, 2160, 3840, 4, "w=1618&h=1000&crop=true",                               dict(w=1618, h=1000, crop=True)),
    ("c3_crop_blur20",          17, 2160, 3840, 4, "w=1618&h=1000&crop=true&blur=20",                       dict(w=1618, h=1000, crop=True, blur=20.0)),
    ("c5_crop_gray_blur10",   5001, 3000, 4000, 3, "w=1618&h=1000&crop=true&grayscale=true&blur=10",        dict(w=1618, h=1000, crop=True, grayscale=True, blur=10.0)),
    ("c5_fit_fill_gray_blur", 5002, 3000, 4000, 3, "w=1618&h=1000&rgb=200,16,99&grayscale=true&blur=10",    dict(w=1618, h=1000, rgb=(200, 16, 99), grayscale=True, blur=10.0)),
    ("c5_crop_inverse_blur",  5003, 3000, 4000, 3, "w=1618&h=1000&crop=true&inverse=true&blur=15",          dict(w=1618, h=1000, crop=True, inverse=True, blur=15.0)),
]


@pytest.mark.parametrize("case", FULL, ids=[c[0] for c in FULL])
def test_baseline_config_full_size(fanlin, dev, case):
    name, seed, h, w, c, qs, kw = case
    img = synth_image(seed, h, w, c)
    if c == 4 and "alpha" in name:
        assert (img[..., 3] != 255).any()  # seed % 8 == 7: the f32 blend / alpha filtering is exercised
    want = O.process(img, **kw)
    got = fanlin.process_image(dev, img, fanlin.Query(qs))
    assert got.shape == want.shape, (got.shape, want.shape)
    hh = hist(got, want)
    print(name, "mismatch histogram", hh)
    assert hh[">=2"] == 0, hh
    assert hh[1] <= got.size // 50, hh  # the fast paths differ by one on a tiny share of values, not on a percent


def test_c1_lenna_full_request(fanlin, dev, lenna):
    want = O.process(lenna, w=300, h=200, rgb=(32, 32, 32))
    got = fanlin.process_image(dev, lenna, fanlin.Query("w=300&h=200&rgb=32,32,32"))
    assert got.shape == (200, 300, 4)
    hh = hist(got, want)
    assert hh[">=2"] == 0, hh
    # fill is bit-exact: the letterbox bars are the fill colour, opaque
    assert (got[:, :50] == np.array([32, 32, 32, 255], np.uint8)).all() and (got[:, 250:] == np.array([32, 32, 32, 255], np.uint8)).all()


@pytest.mark.parametrize("sigma", [0.8, 1.3, 10.3, 12.4, 17.76, 20.0])
def test_blur_fractional_sigma(fanlin, dev, sigma):
    """The C ABI takes any float sigma (Query::blur() only yields integers): the window is
    ceil(2 sigma - 0.5) taps either side (10.3 -> 43 taps, 0.8 -> 5), not floor(2 sigma)."""
    img = synth_image(77, 97, 131, 4 if sigma > 5 else 3)
    j = fanlin.make_job(img, fanlin.Query(""))
    j.blur_sigma = sigma
    p = fanlin.plan_job(j)
    out = np.zeros((p.out_h, p.out_w, p.out_channels), np.uint8)
    j.dst, j.dst_capacity = out.ctypes.data, out.nbytes
    dev.run([j])
    want = O.process(img, blur=float(np.float32(sigma)))
    hh = hist(out, want)
    assert out.shape == want.shape and hh[">=2"] == 0, (sigma, hh)


def test_device_batch_equals_single_requests_full_size(fanlin, dev):
    """8 C3-shaped images (one with random alpha) as ONE device-resident batch: the bytes a merged
    launch produces are the bytes each request produces alone."""
    import torch

    qs = "w=1618&h=1000&crop=true&blur=10"
    seeds = [40, 41, 42, 47]  # 47 % 8 == 7: random alpha
    imgs = [synth_image(s, 2160, 3840, 4) for s in seeds]
    singles = [fanlin.process_image(dev, im, fanlin.Query(qs)) for im in imgs]
    device = torch.device("cuda", 0)
    src = torch.stack([torch.from_numpy(im) for im in imgs]).to(device)
    dst = torch.zeros((len(imgs), 1000, 1618, 4), dtype=torch.uint8, device=device)
    torch.cuda.synchronize()
    q = fanlin.Query(qs)
    jobs = (fanlin.Job * len(imgs))()
    for i in range(len(imgs)):
        fanlin.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(jobs[i]))
        jobs[i].src = src.data_ptr() + i * imgs[0].nbytes
        jobs[i].src_w, jobs[i].src_h, jobs[i].src_channels = 3840, 2160, 4
        jobs[i].dst = dst.data_ptr() + i * 1000 * 1618 * 4
        jobs[i].dst_capacity = 1000 * 1618 * 4
    b = dev.prepare(jobs, 0)
    b.launch(None)
    b.launch(None)  # relaunching a prepared batch is idempotent
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    b.free()
    for i, s in enumerate(singles):
        assert np.array_equal(got[i], s), (seeds[i], hist(got[i], s))


def test_tables_are_cached_per_geometry(fanlin):
    """The context keeps the filter tables of every geometry it has seen on the device (TableGen, csrc/runtime.h): a
    second request -- or batch -- of the same geometry uploads no table bytes, which is what lets a single request take
    the both-passes tensor-core kernels (~1 MB of weight tiles per geometry)."""
    d = fanlin.Device([0])
    try:
        a = synth_image(2001, 1080, 1920, 3)
        b = synth_image(2002, 1080, 1920, 3)
        q = fanlin.Query("w=300&h=200")
        out_a = fanlin.process_image(d, a, q)
        t1 = d.stats()["table_bytes"]
        assert t1 > 100_000  # the per-chunk weight tiles of the 1080p -> 300x169 geometry
        out_b = fanlin.process_image(d, b, q)
        outs = fanlin.process_images(d, [a, b, a], q)
        assert d.stats()["table_bytes"] == t1, "a geometry the context has seen uploaded table bytes again"
        assert np.array_equal(outs[0], out_a) and np.array_equal(outs[1], out_b) and np.array_equal(outs[2], out_a)
        fanlin.process_image(d, a, fanlin.Query("w=320&h=200"))  # another geometry: new tables
        assert d.stats()["table_bytes"] > t1
        want = O.process(a, w=300, h=200)
        assert hist(out_a, want)[">=2"] == 0
    finally:
        d.close()


def test_concurrent_prepares_share_the_table_generation(fanlin):
    """Several threads preparing batches of old and new geometries on one device at once (the batcher thread and users of the
    device-batch API do): results identical to the sequential ones."""
    import threading

    d = fanlin.Device([0])
    try:
        imgs = [synth_image(300 + i, 400 + 16 * i, 640, 3) for i in range(6)]
        qs = ["w=200&h=150", "w=160&h=120&crop=true", "w=100&h=100&blur=10"]
        want = {(i, k): fanlin.process_image(d, imgs[i], fanlin.Query(qs[k])) for i in range(6) for k in range(3)}
        d2 = fanlin.Device([0])
        got, errs = {}, []

        def work(i):
            try:
                for k in range(3):
                    got[(i, k)] = fanlin.process_image(d2, imgs[i], fanlin.Query(qs[k]))
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
        [t.start() for t in th]
        [t.join() for t in th]
        d2.close()
        assert not errs, errs
        for key, w in want.items():
            assert np.array_equal(got[key], w), key
    finally:
        d.close()


@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_inverse_rides_on_the_vertical_pass(fanlin, c):
    """inverse=true on the tensor-core resample: sum q (255 - x) = 255 sum q - sum q x in the exact integer contraction, so
    no pass over the source inverts its bytes (the launch list holds no colour pass), alpha is left alone, and the
    result is what inverting first gives (handler.rs:226-228 inverts BEFORE the resize)."""
    import torch

    d = fanlin.Device([0])
    try:
        h, w = 1080, 1920
        img = synth_image(70 + c + (7 - (70 + c) % 8 if c in (2, 4) else 0), h, w, c)  # LA / RGBA: a seed with random alpha
        if c in (2, 4):
            assert (img[..., c - 1] != 255).any()
        want = O.process(img, w=300, h=200, inverse=True, rgb=(9, 8, 7))
        got = fanlin.process_image(d, img, fanlin.Query("w=300&h=200&inverse=true&rgb=9,8,7"))
        hh = hist(got, want)
        assert got.shape == want.shape and hh[">=2"] == 0, hh
        src = torch.from_numpy(img).cuda()
        dst = torch.zeros(want.shape, dtype=torch.uint8, device="cuda")
        j = fanlin.make_job(img, fanlin.Query("w=300&h=200&inverse=true&rgb=9,8,7"))
        j.src, j.dst, j.dst_capacity = src.data_ptr(), dst.data_ptr(), want.size
        torch.cuda.synchronize()
        b = d.prepare([j], 0)
        b.set_timing(True)
        b.launch(None)
        torch.cuda.synchronize()
        names = [k for k, _ in b.kernel_times()]
        b.free()
        assert names and all("color_pass" not in k for k in names), names
        assert np.array_equal(dst.cpu().numpy(), got)
    finally:
        d.close()


@pytest.mark.parametrize("exif,h,w,c,qs,kw", [
    (6, 1920, 1080, 3, "w=300&h=200", dict(w=300, h=200)),                                                  # C2 stored rotated (a phone's portrait frame)
    (3, 1080, 1920, 3, "w=300&h=200&rgb=1,2,3", dict(w=300, h=200, rgb=(1, 2, 3))),
    (8, 2160, 3840, 4, "w=1000&h=1618&crop=true&blur=10", dict(w=1000, h=1618, crop=True, blur=10.0)),       # odd crop difference, mirrored axis
    (5, 1500, 2001, 3, "w=333&h=251&crop=true&grayscale=true", dict(w=333, h=251, crop=True, grayscale=True)),
    (7, 1001, 1333, 4, "w=400&h=300&inverse=true", dict(w=400, h=300, inverse=True)),
    (2, 1080, 1920, 1, "w=301&h=170&crop=true", dict(w=301, h=170, crop=True)),
    (4, 777, 1234, 2, "w=200&h=200", dict(w=200, h=200)),
])
def test_orientation_behind_the_resample_full_size(fanlin, dev, exif, h, w, c, qs, kw):
    """EXIF orientation 2..8 at BASELINE sizes: on the fast paths the Lanczos3 stage runs on the image as stored and the small
    output is oriented (plan.h stored_axes_stage); the oracle orients first (handler.rs:221-223)."""
    img = synth_image(600 + exif + (7 - (600 + exif) % 8 if c in (2, 4) else 0), h, w, c)
    want = O.process(img, orientation=exif, **kw)
    got = fanlin.process_image(dev, img, fanlin.Query(qs), orientation=exif)
    hh = hist(got, want)
    assert got.shape == want.shape and hh[">=2"] == 0, (exif, hh)
    assert hh[1] <= max(64, got.size // 100), hh


def test_table_generation_rolls_over(fanlin):
    """The per-device table generation is REPLACED, never edited, once it has grown past its bound (768 MB; 4 MB here through
    FANLIN_TABLE_GEN_LIMIT_MB, read once per process -- so this runs in a process of its own): geometries keep working
    across the roll-over, results stay what a fresh context gives, and batches prepared before it still launch from the
    tables they were built with."""
    import os
    import subprocess
    import sys

    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import __graft_entry__ as G
from synth import synth_image
import torch
pkg = G.load_package()
d = pkg.Device([0])
ref = pkg.Device([0])
img = synth_image(5, 1080, 1920, 3)
src = torch.from_numpy(img).cuda(); dst = torch.zeros((200, 300, 4), dtype=torch.uint8, device='cuda')
j = pkg.make_job(img, pkg.Query('w=300&h=200')); j.src, j.dst, j.dst_capacity = src.data_ptr(), dst.data_ptr(), 240000
torch.cuda.synchronize()
old_batch = d.prepare([j], 0)   # holds the first generation's device tables
t_prev, rolled = 0, 0
for k in range(12):             # ~1 MB of tiles per geometry: the 4 MB bound is passed several times
    q = pkg.Query(f'w={300 + 4 * k}&h=200')
    a = pkg.process_image(d, img, q)
    t = d.stats()['table_bytes']
    assert t > t_prev, 'a new geometry must upload tables'
    t_prev = t
    assert np.array_equal(a, pkg.process_image(ref, img, q)), k
old_batch.launch(None); torch.cuda.synchronize()
assert np.array_equal(dst.cpu().numpy(), pkg.process_image(ref, img, pkg.Query('w=300&h=200')))
old_batch.free(); d.close(); ref.close()
print('ROLLOVER_OK')
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FANLIN_TABLE_GEN_LIMIT_MB="4")
    r = subprocess.run([sys.executable, "-c", code % (root, os.path.join(root, "tests"))], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ROLLOVER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
