"""The arithmetic of the tensor-core resample, modelled in numpy and held against the oracle on the CPU.

The shipped Lanczos3 path (fanlin-rs_b200/csrc/fused_tc.cpp + kernels_fused_tc.cu / kernels_fused_tc3.cu)
does not compute in the crate's f32 recipe:

  vertical   q = lround(w * 2^sh) as three signed base-128 digits (s8), u8 x s8 -> s32 on the tensor cores: EXACT integers;
             recombined per element as  v = f32(mid * 128 + lo) * 2^-sh  (+)  f32(hi) * 2^(14 - sh)   (two FMAs);
  horizontal v split into f16 halves hi = f16(v), lo = f16(v - hi); weights x 16 split the same way on the host;
             D = T_hi.W_hi + T_lo.W_hi + T_hi.W_lo in f32 (products of f16 halves are exact in f32; T_lo.W_lo is dropped);
  rounding   cvt.rzi.sat.u8(add.rz(D / 16, 0.5)).

This file restates exactly that (the accumulation order of the tensor core is the one thing it cannot know: it sums in
f64 and rounds once, and the bound below leaves room for it) and checks, against the oracle (image 0.25.6's
vertical_sample -> horizontal_sample, src/handler.rs:230-237 of the reference):

  * the value before rounding stays within 0.002 LSB of the exact (real-arithmetic) separable filter -- so a result can
    differ from the oracle's only where the oracle's own f32 value lies that close to a rounding boundary;
  * after rounding no value is off by two, and off-by-one values are as rare as on the device
    (tools/check_hmma.py on a B200: 2-24 per 240 k-6.5 M values).

It is a model of the kernels' arithmetic, not the kernels: the `-m gpu` parity tests are the proof for those.  What it
guards on a machine without a GPU is the scheme itself and the constants the host builder uses (the shift rule, the
digit range, the x16 weight scale).
"""
import numpy as np
import pytest

from oracle import oracle as O
from synth import synth_image

TC2_WSCALE = 16.0  # fused_tc.h


def _weight_shift(ws):
    """fused_tc.cpp weight_shift: the largest sh <= 30 with max |w| * 2^sh <= 2 080 000 (three base-128 digits)."""
    maxw = float(np.abs(ws).max())
    sh = 30
    while sh > 0 and maxw * 2.0 ** sh > 2080000.0:
        sh -= 1
    return sh


def _lround(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)


def _digits(q):
    lo = ((q + 64) & 127) - 64
    q1 = (q - lo) // 128
    mid = ((q1 + 64) & 127) - 64
    hi = (q1 - mid) // 128
    return hi, mid, lo


def _f16_split(v32):
    hi = v32.astype(np.float16)
    lo = (v32 - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def _dense(lefts, counts, ws, n_in, dtype):
    m = np.zeros((len(lefts), n_in), dtype)
    for o in range(len(lefts)):
        m[o, lefts[o]:lefts[o] + counts[o]] = ws[o, :counts[o]]
    return m


def model_resize(img, nw, nh, kind=O.LANCZOS3, sigma=0.0):
    """(u8 result of the modelled arithmetic, its value before rounding as f64, the exact separable filter as f64)."""
    h, w, c = img.shape
    vl, vc, vw = O.weight_table(kind, h, nh, sigma)
    hl, hc, hw = O.weight_table(kind, w, nw, sigma)
    # ---- vertical: exact integer contraction of base-128 digits ----
    sh = _weight_shift(vw)
    assert sh >= 20, sh  # 2^-21 steps for a downscale; an upscale's centre weight can reach 1.0 and costs one bit
    q = _dense(vl, vc, _lround(vw.astype(np.float64) * 2.0 ** sh), h, np.int64)
    dh, dm, dl = _digits(q)
    assert dh.min() >= -128 and dh.max() <= 127 and np.array_equal((dh * 128 + dm) * 128 + dl, q)
    src = img.reshape(h, w * c).astype(np.int64)
    s_hi, s_mid, s_lo = dh @ src, dm @ src, dl @ src  # s32 accumulators in TMEM
    assert max(np.abs(s_hi).max(), np.abs(s_mid).max(), np.abs(s_lo).max()) < 2 ** 31
    scale, scale_hi = np.float32(2.0 ** -sh), np.float32(2.0 ** (14 - sh))
    low = (s_mid * 128 + s_lo).astype(np.float32)  # float(int(mid) * 128 + int(lo)): one rounding for |x| >= 2^24
    v = (low.astype(np.float64) * np.float64(scale)).astype(np.float32)                                  # ffma2(r, fl, scale)
    v = (s_hi.astype(np.float32).astype(np.float64) * np.float64(scale_hi) + v.astype(np.float64)).astype(np.float32)  # ffma2(r, fh, scale_hi)
    # ---- horizontal: f16 hi / lo halves, f32 accumulators ----
    t_hi, t_lo = _f16_split(v)
    w16 = (_dense(hl, hc, hw, w, np.float32) * np.float32(TC2_WSCALE)).astype(np.float32)
    w_hi, w_lo = _f16_split(w16)
    assert np.all(np.isfinite(t_hi.astype(np.float32))) and np.all(np.isfinite(w_hi.astype(np.float32)))
    T_hi = t_hi.astype(np.float64).reshape(nh, w, c)
    T_lo = t_lo.astype(np.float64).reshape(nh, w, c)
    W_hi, W_lo = w_hi.astype(np.float64), w_lo.astype(np.float64)
    d = np.einsum("ok,rkc->roc", W_hi, T_hi) + np.einsum("ok,rkc->roc", W_hi, T_lo) + np.einsum("ok,rkc->roc", W_lo, T_hi)
    d32 = d.astype(np.float32)
    pre = d32 * np.float32(1.0 / TC2_WSCALE)
    # round_u8: add.rz(t, 0.5) then cvt.rzi.sat.u8 -- in f64 the sum is exact, truncating it is what both .rz steps give
    out = np.clip(np.trunc(pre.astype(np.float64) + 0.5), 0, 255).astype(np.uint8)
    # ---- the exact separable filter on the same f32 weights ----
    exact = np.einsum("ok,rkc->roc", _dense(hl, hc, hw, w, np.float64),
                      (_dense(vl, vc, vw, h, np.float64) @ src.astype(np.float64)).reshape(nh, w, c))
    return out, pre.astype(np.float64), exact


CASES = [  # (seed, h, w, c, nw, nh): the BASELINE ratios at sizes the model finishes in seconds
    (1, 270, 480, 3, 75, 42),     # C2's ratio (6.4): 1080p -> 300x169 at a quarter of the size
    (2, 256, 256, 3, 150, 150),   # C1 (512 -> 200 on the short side)
    (3, 216, 384, 4, 162, 91),    # C3's ratio (2.37), RGBA with random alpha
    (4, 300, 400, 1, 162, 121),   # C5 after grayscale: one plane
    (5, 97, 131, 2, 211, 160),    # upscale, LA
    (6, 64, 64, 3, 7, 5),         # a strong downscale: the widest windows
]


@pytest.mark.parametrize("seed,h,w,c,nw,nh", CASES)
def test_modelled_tensor_core_arithmetic_stays_within_one_lsb_of_the_oracle(seed, h, w, c, nw, nh):
    img = synth_image(seed, h, w, c)
    got, pre, exact = model_resize(img, nw, nh)
    want = O.resize(img, nw, nh, O.LANCZOS3)
    assert got.shape == want.shape
    # before rounding: within 0.002 LSB (measured: <= 5.2e-4) of the real-arithmetic filter wherever that is not clamped away
    inside = (exact > -0.5) & (exact < 255.5)
    assert np.abs(pre - exact)[inside].max() < 2e-3
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1
    assert (diff == 1).mean() < 2e-4, (diff == 1).sum()
    # every off-by-one value is one whose exact value lies next to a rounding boundary: a tie the oracle's own f32
    # accumulation decides one way and this arithmetic the other
    if (diff == 1).any():
        frac = np.abs((exact[diff == 1] + 0.5) - np.round(exact[diff == 1] + 0.5))
        assert frac.max() < 2e-3


def test_noise_image_worst_case_for_the_f16_split():
    # full-range noise: the largest |v| (ringing beyond 0..255) and the least cancellation in the dropped lo x lo term
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    img[::7] = 255
    img[3::11] = 0
    got, pre, exact = model_resize(img, 61, 47)
    want = O.resize(img, 61, 47, O.LANCZOS3)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    inside = (exact > -0.5) & (exact < 255.5)
    assert np.abs(pre - exact)[inside].max() < 2e-3


def test_gaussian_weights_through_the_same_arithmetic():
    # the blur's contraction (blur_tc.cpp: same digit rule and x16 f16 halves) with the table's normalised weights in
    # both passes: digits + f16 halves on 41- and 81-tap Gaussians (the kernel's own horizontal form -- interior weights
    # times a per-column border factor -- is model_blur_tc2 below)
    img = synth_image(9, 90, 110, 3)
    for sigma in (10.0, 20.0):
        got, pre, exact = model_resize(img, 110, 90, O.GAUSSIAN_BLUR, sigma)
        want = O.blur(img, sigma)
        assert np.abs(pre - exact).max() < 2e-3
        diff = np.abs(got.astype(int) - want.astype(int))
        assert diff.max() <= 1 and (diff == 1).mean() < 1e-3


def model_blur_tc2(img, sigma):
    """blur_tc2_kernel's arithmetic (blur_tc.cpp / blur.cpp blur_build): vertical pass with the table's own (border-
    renormalised) taps as digits, horizontal pass against ONE Toeplitz tile of the INTERIOR weights u (x 16, f16 halves),
    the border renormalisation as one f32 factor per output column, corr[j] = w_table(j, centre tap) / u_centre."""
    h, w, c = img.shape
    radius = int(np.ceil(2.0 * sigma - 0.5))
    n_ref = 4 * radius + 4
    rl, rc, rw = O.weight_table(O.GAUSSIAN_BLUR, n_ref, n_ref, sigma)
    o_ref = 2 * radius + 2
    assert rc[o_ref] == 2 * radius + 1
    u = rw[o_ref, :rc[o_ref]].astype(np.float32)
    vl, vc, vw = O.weight_table(O.GAUSSIAN_BLUR, h, h, sigma)
    hl, hc, hw = O.weight_table(O.GAUSSIAN_BLUR, w, w, sigma)
    sh = _weight_shift(vw)
    q = _dense(vl, vc, _lround(vw.astype(np.float64) * 2.0 ** sh), h, np.int64)
    dh, dm, dl = _digits(q)
    src = img.reshape(h, w * c).astype(np.int64)
    s_hi, s_mid, s_lo = dh @ src, dm @ src, dl @ src
    scale, scale_hi = np.float32(2.0 ** -sh), np.float32(2.0 ** (14 - sh))
    v = ((s_mid * 128 + s_lo).astype(np.float32).astype(np.float64) * np.float64(scale)).astype(np.float32)
    v = (s_hi.astype(np.float32).astype(np.float64) * np.float64(scale_hi) + v.astype(np.float64)).astype(np.float32)
    t_hi, t_lo = _f16_split(v)
    toe = np.zeros((w, w), np.float32)  # Toeplitz: output j takes u[k - j + R] from source k (zeros outside the image)
    for j in range(w):
        k0, k1 = max(0, j - radius), min(w, j + radius + 1)
        toe[j, k0:k1] = u[k0 - j + radius:k1 - j + radius]
    w_hi, w_lo = _f16_split((toe * np.float32(TC2_WSCALE)).astype(np.float32))
    T_hi, T_lo = t_hi.astype(np.float64).reshape(h, w, c), t_lo.astype(np.float64).reshape(h, w, c)
    d = (np.einsum("ok,rkc->roc", w_hi.astype(np.float64), T_hi) + np.einsum("ok,rkc->roc", w_hi.astype(np.float64), T_lo)
         + np.einsum("ok,rkc->roc", w_lo.astype(np.float64), T_hi)).astype(np.float32)
    corr = np.array([hw[j, j - hl[j]] for j in range(w)], np.float32) / u[radius]   # f32 division, as the host does
    corr16 = (corr * np.float32(1.0 / TC2_WSCALE)).astype(np.float32)
    pre = (d * corr16[None, :, None]).astype(np.float32)
    out = np.clip(np.trunc(pre.astype(np.float64) + 0.5), 0, 255).astype(np.uint8)
    exact = np.einsum("ok,rkc->roc", _dense(hl, hc, hw, w, np.float64),
                      (_dense(vl, vc, vw, h, np.float64) @ src.astype(np.float64)).reshape(h, w, c))
    return out, pre.astype(np.float64), exact


@pytest.mark.parametrize("sigma,h,w,c", [(10.0, 70, 95, 4), (20.0, 100, 130, 1), (10.3, 50, 64, 3), (12.0, 30, 25, 2)])
def test_modelled_blur_with_interior_weights_and_border_factors(sigma, h, w, c):
    # images narrower than the window too (25 columns under 49 taps): every column is a border column there
    img = synth_image(int(sigma * 10) + c, h, w, c)
    got, pre, exact = model_blur_tc2(img, sigma)
    want = O.blur(img, sigma)
    assert np.abs(pre - exact).max() < 2e-3
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1 and (diff == 1).mean() < 1e-3


def test_shift_rule_leaves_headroom_for_every_lanczos3_and_gaussian_table():
    # |q| <= 2 080 000 < 127 * 16384 + 63 * 128 + 63 (the largest three-digit value, 2 088 895) for every geometry
    # the BASELINE configs use; sh is at least 21 for a downscale or a blur, 20 for an upscale (centre weight up to ~1.0)
    for kind, n_in, n_out, sg in [(O.LANCZOS3, 1080, 169, 0), (O.LANCZOS3, 1920, 300, 0), (O.LANCZOS3, 2160, 1000, 0),
                                  (O.LANCZOS3, 3000, 1214, 0), (O.LANCZOS3, 512, 200, 0), (O.LANCZOS3, 100, 1000, 0),
                                  (O.GAUSSIAN_BLUR, 1000, 1000, 10.0), (O.GAUSSIAN_BLUR, 1000, 1000, 20.0)]:
        _, _, ws = O.weight_table(kind, n_in, n_out, sg)
        sh = _weight_shift(ws)
        q = _lround(ws.astype(np.float64) * 2.0 ** sh)
        hi, mid, lo = _digits(q)
        assert (21 if n_out <= n_in else 20) <= sh <= 30 and np.abs(q).max() <= 2080000
        assert hi.min() >= -128 and hi.max() <= 127 and mid.min() >= -64 and mid.max() <= 63 and lo.min() >= -64 and lo.max() <= 63
