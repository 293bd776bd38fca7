"""YCCK -> CMYK, the float loop of convert_jpeg_color_if_needed (reference src/handler.rs:420-439;
SURVEY.md 8f rank 3).  Unlike the image-crate arithmetic this loop is IN the reference tree, so the
oracle is pinned by the reference's own source: the known answers below are worked by hand from the
expression at :427-431 and checked against an independent numpy float32 restatement; the CUDA kernel
must be bit-exact on all 2^24 (Y, Cb, Cr) triples."""
import numpy as np
import pytest

from oracle import oracle as O


def np_ycck(raw):
    """Second restatement: numpy float32, one rounded operation at a time, Rust's left-to-right order."""
    a = np.asarray(raw, np.uint8).reshape(-1, 4)
    f = np.float32
    y, cb, cr = a[:, 0].astype(f), a[:, 1].astype(f), a[:, 2].astype(f)
    r = (y + f(1.40200) * cr) - f(179.456)
    g = ((y - f(0.34414) * cb) - f(0.71414) * cr) + f(135.45984)
    b = (y + f(1.77200) * cb) - f(226.816)
    out = np.empty_like(a)
    for k, v in enumerate((r, g, b)):
        assert v.dtype == np.float32
        out[:, k] = np.clip(v, f(0), f(255)).astype(np.uint8)  # truncation, like `as u8`
    out[:, 3] = 255 - a[:, 3]
    return out.reshape(-1)


def test_known_answers():
    # y = cb = cr = 128: r = 128 + 179.456 - 179.456 = 128; g = 128 - 44.04992 - 91.40992 + 135.45984 = 128;
    # b = (128 + 226.816) - 226.816 in f32 = 127.99998 -> 127 (truncating cast, as in the reference)
    assert list(O.ycck_to_cmyk([128, 128, 128, 0])) == [128, 128, 127, 255]
    # white-ish: everything clamps high except g = 255 - 87.76 - 182.11 + 135.46 = 120.59 -> 120
    assert list(O.ycck_to_cmyk([255, 255, 255, 255])) == [255, 120, 255, 0]
    # zeros: r = -179.456 -> 0, g = 135.46 -> 135, b = -226.8 -> 0
    assert list(O.ycck_to_cmyk([0, 0, 0, 10])) == [0, 135, 0, 245]
    # y=200 cb=30 cr=220: r = 200 + 308.44 - 179.456 = 328.98 -> 255; g = 200 - 10.32 - 157.11 + 135.46 = 168.02 -> 168;
    # b = 200 + 53.16 - 226.816 = 26.34 -> 26
    assert list(O.ycck_to_cmyk([200, 30, 220, 99])) == [255, 168, 26, 156]


def test_oracle_matches_numpy_restatement_on_all_triples():
    y, cb, cr = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(0, 256, 5, dtype=np.uint8), indexing="ij")
    raw = np.stack([y, cb, cr, (y ^ cb)], axis=-1).reshape(-1)
    assert np.array_equal(O.ycck_to_cmyk(raw), np_ycck(raw))


def test_partial_trailing_pixel_is_left_alone():
    raw = np.array([1, 2, 3, 4, 9, 9], np.uint8)
    out = O.ycck_to_cmyk(raw)
    assert list(out[4:]) == [9, 9]


@pytest.mark.gpu
def test_device_bit_exact_on_all_triples(fanlin):
    dev = fanlin.Device([0])
    try:
        y, cb, cr = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
        raw = np.stack([y, cb, cr, (y + 3 * cb + 7 * cr).astype(np.uint8)], axis=-1).reshape(-1)  # 64 MB: three chunks through two streams
        got = dev.ycck_to_cmyk(raw)
        assert np.array_equal(got, O.ycck_to_cmyk(raw))
        # odd pixel counts and unaligned buffers take the per-pixel tail
        for n_px, off in [(1, 0), (3, 1), (1027, 3), (4099, 2)]:
            buf = np.random.default_rng(n_px).integers(0, 256, 4 * n_px + off, dtype=np.uint8)
            assert np.array_equal(dev.ycck_to_cmyk(buf[off:]), O.ycck_to_cmyk(buf[off:])), (n_px, off)
    finally:
        dev.close()


@pytest.mark.gpu
def test_device_resident_in_place(fanlin):
    import torch

    dev = fanlin.Device([0])
    try:
        raw = np.random.default_rng(5).integers(0, 256, 4 * 1920 * 1080, dtype=np.uint8)
        t = torch.from_numpy(raw).cuda()
        torch.cuda.synchronize()
        dev.ycck_to_cmyk_device(t.data_ptr(), t.data_ptr(), raw.size // 4)  # in place, like the reference
        dev.ycck_to_cmyk_device(t.data_ptr(), t.data_ptr(), 0)
        import ctypes as C
        # the context's own stream: synchronise the device before reading
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), O.ycck_to_cmyk(raw))
    finally:
        dev.close()
