"""Regenerates tests/golden/ from the reference's fixture and the CPU oracle.

Run in the build container only (reads /root/reference/images/lenna.png):
    python tests/golden/make_golden.py
Writes
  lenna_512_rgb.png   the decoded pixels of the reference's images/lenna.png
                      (512x512 RGB8, lossless) -- the C1 input
  golden.json         per case: params, output dims/channels, sha256 of the
                      oracle's output bytes
The hashes are produced by oracle/fanlin_oracle.c, NOT by the reference (which
cannot be built here: no Rust toolchain, `image` 0.25.6 not vendored).  They pin
the oracle against regressions and let the GPU box check its locally compiled
oracle; they do not pin the oracle to the reference ("parity unpinned").
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from synth import synth_image  # noqa: E402

# name, input ("lenna" | [seed,h,w,c]), query-accessor values
CASES = [
    ("c1_lenna_fit_fill", "lenna", dict(w=300, h=200, rgb=[32, 32, 32])),
    ("c1_lenna_fit_fill_default_rgb", "lenna", dict(w=300, h=200)),
    ("lenna_crop", "lenna", dict(w=300, h=200, crop=True)),
    ("lenna_gray_fit", "lenna", dict(w=300, h=200, grayscale=True)),
    ("lenna_gray_beats_inverse", "lenna", dict(w=300, h=200, grayscale=True, inverse=True)),
    ("lenna_inverse_fit", "lenna", dict(w=300, h=200, inverse=True)),
    ("lenna_blur10", "lenna", dict(w=300, h=200, blur=10.0)),
    ("lenna_blur20_only", "lenna", dict(blur=20.0)),
    ("lenna_gray_only", "lenna", dict(grayscale=True)),
    ("lenna_inverse_only", "lenna", dict(inverse=True)),
    ("lenna_upscale_fit", "lenna", dict(w=2000, h=1000, rgb=[1, 2, 3])),
    ("lenna_same_dims", "lenna", dict(w=512, h=512)),
    ("lenna_same_w_fill", "lenna", dict(w=600, h=512, rgb=[200, 100, 0])),
    ("c2_small_1080p_fit_fill", [2000, 1080, 1920, 3], dict(w=300, h=200)),
    ("c3_small_rgba_crop_blur", [3007, 540, 960, 4], dict(w=404, h=250, crop=True, blur=10.0)),
    ("c3_opaque_rgba_crop_blur", [3000, 540, 960, 4], dict(w=404, h=250, crop=True, blur=10.0)),
    ("c4_gif_literal", [4000, 270, 480, 4], dict(grayscale=True, inverse=True, gif=True)),
    ("c4_gif_w200_h113", [4001, 270, 480, 4], dict(w=200, h=113, grayscale=True, inverse=True, gif=True)),
    ("c4_gif_w200_h200_alpha", [4007, 270, 480, 4], dict(w=200, h=200, grayscale=True, gif=True)),
    ("gif_inverse_fill_alpha", [4015, 270, 480, 4], dict(w=200, h=200, inverse=True, gif=True, rgb=[9, 8, 7])),
    ("gif_crop", [4002, 270, 480, 4], dict(w=100, h=100, crop=True, gif=True)),
    ("c5_small_crop", [5000, 750, 1000, 3], dict(w=404, h=250, crop=True, grayscale=True, blur=10.0)),
    ("c5_small_fit", [5001, 750, 1000, 3], dict(w=404, h=250, rgb=[10, 20, 30], grayscale=True, blur=10.0)),
    ("rgba_alpha_fit_fill", [7, 97, 131, 4], dict(w=120, h=120, rgb=[255, 0, 128])),
    ("la_alpha_fit_fill", [15, 97, 131, 2], dict(w=120, h=120, rgb=[0, 255, 0])),
    ("l8_crop_tall", [21, 131, 41, 1], dict(w=30, h=90, crop=True)),
    ("tiny_1x1_up", [22, 1, 1, 3], dict(w=20, h=20)),
    ("tiny_1xN", [23, 1, 77, 3], dict(w=20, h=20)),
    ("tiny_Nx1", [24, 77, 1, 4], dict(w=20, h=20, crop=True)),
    ("ragged_prime", [25, 211, 307, 3], dict(w=97, h=89, crop=True, blur=15.0)),
    ("blur_sigma20_small_img", [26, 17, 23, 4], dict(blur=20.0)),
    # EXIF orientation (handler.rs:221-223): rotate 180, rotate 90 clockwise, transverse
    ("orient3_lenna_fit_fill", "lenna", dict(w=300, h=200, rgb=[32, 32, 32], orientation=3)),
    ("orient6_portrait_crop_gray", [31, 300, 200, 3], dict(w=120, h=90, crop=True, grayscale=True, orientation=6)),
    ("orient7_rgba_fit_blur", [32, 150, 260, 4], dict(w=100, h=100, rgb=[5, 6, 7], blur=10.0, orientation=7)),
]


def load_input(spec):
    if spec == "lenna":
        return np.asarray(Image.open(os.path.join(HERE, "lenna_512_rgb.png")).convert("RGB"))
    seed, h, w, c = spec
    return synth_image(seed, h, w, c)


def main():
    ref = "/root/reference/images/lenna.png"
    if os.path.exists(ref):
        Image.open(ref).convert("RGB").save(os.path.join(HERE, "lenna_512_rgb.png"), optimize=True)
    table = []
    for name, spec, kw in CASES:
        img = load_input(spec)
        okw = {k: (tuple(v) if k == "rgb" else v) for k, v in kw.items()}
        out = O.process(img, **okw)
        table.append(dict(name=name, input=spec, params=kw, out_h=out.shape[0], out_w=out.shape[1],
                          out_c=out.shape[2], sha256=hashlib.sha256(out.tobytes()).hexdigest()))
        print(name, out.shape, table[-1]["sha256"][:12])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(oracle="oracle/fanlin_oracle.c", pinned_to_reference=False, cases=table), f, indent=1)


if __name__ == "__main__":
    main()
