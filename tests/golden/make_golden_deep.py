"""Regenerates tests/golden/golden_deep.json: SHA-256 of the oracle's output for requests on 16-bit and f32 images
(oracle/fanlin_oracle_deep.c).  As golden.json: the hashes pin the oracle against regressions and let the GPU box
check its locally compiled oracle; they do NOT come from the reference ("parity unpinned": no Rust toolchain here).
    python tests/golden/make_golden_deep.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from synth import synth_deep  # noqa: E402

# name, [seed, h, w, c, dtype], request (the accessor values of query::Query + orientation / encoder branch)
CASES = [
    ("rgb16_fit_fill", [100, 97, 131, 3, "uint16"], dict(w=120, h=120, rgb=[255, 0, 128])),
    ("rgba16_alpha_fit_fill", [107, 97, 131, 4, "uint16"], dict(w=120, h=120)),
    ("rgb16_crop", [101, 211, 307, 3, "uint16"], dict(w=97, h=89, crop=True)),
    ("rgba16_gray_crop_blur", [102, 120, 160, 4, "uint16"], dict(w=64, h=50, crop=True, grayscale=True, blur=10.0)),
    ("l16_inverse_fit", [103, 131, 41, 1, "uint16"], dict(w=30, h=90, inverse=True)),
    ("la16_alpha_fit_fill_blur", [111, 60, 90, 2, "uint16"], dict(w=80, h=80, blur=10.0)),
    ("rgb16_gray_only", [104, 40, 50, 3, "uint16"], dict(grayscale=True)),
    ("rgb16_blur_only", [105, 40, 50, 3, "uint16"], dict(blur=12.0)),
    ("rgb16_upscale", [106, 9, 13, 3, "uint16"], dict(w=40, h=40, crop=True)),
    ("rgb16_orient6_fit", [108, 90, 60, 3, "uint16"], dict(w=50, h=50, orientation=6)),
    ("rgba16_orient7_gray_crop", [109, 90, 60, 4, "uint16"], dict(w=40, h=30, crop=True, grayscale=True, orientation=7)),
    ("rgb16_to_rgb8", [110, 97, 131, 3, "uint16"], dict(w=60, h=45, crop=True, to_rgb8=True)),
    ("la16_to_rgba8", [119, 97, 131, 2, "uint16"], dict(w=60, h=45, crop=True, to_rgba8=True)),
    ("l16_same_dims", [112, 33, 44, 1, "uint16"], dict(w=44, h=33)),
    ("rgb32f_fit_fill", [120, 97, 131, 3, "float32"], dict(w=120, h=120, rgb=[1, 2, 3])),
    ("rgba32f_alpha_crop", [127, 97, 131, 4, "float32"], dict(w=70, h=60, crop=True)),
    ("rgb32f_gray_crop_blur", [121, 80, 120, 3, "float32"], dict(w=64, h=50, crop=True, grayscale=True, blur=10.0)),
    ("rgba32f_inverse_only", [122, 40, 50, 4, "float32"], dict(inverse=True)),
    ("rgba32f_gray_fit_fill", [135, 50, 90, 4, "float32"], dict(w=60, h=60, grayscale=True)),
    ("rgb32f_blur_only", [123, 40, 50, 3, "float32"], dict(blur=10.0)),
    ("rgb32f_orient3_to_rgb8", [124, 70, 50, 3, "float32"], dict(w=35, h=35, orientation=3, to_rgb8=True)),
    ("rgba32f_orient5_to_rgba8_blur", [125, 70, 50, 4, "float32"], dict(w=40, h=40, crop=True, blur=10.0, orientation=5, to_rgba8=True)),
]


def okw(params):
    return {k: (tuple(v) if k == "rgb" else v) for k, v in params.items()}


def main():
    out = []
    for name, (seed, h, w, c, dt), params in CASES:
        img = synth_deep(seed, h, w, c, np.dtype(dt))
        res = O.process_deep(img, **okw(params))
        out.append(dict(name=name, input=[seed, h, w, c, dt], params=params, out_h=res.shape[0], out_w=res.shape[1], out_c=res.shape[2],
                        out_dtype=str(res.dtype), sha256=hashlib.sha256(res.tobytes()).hexdigest()))
    with open(os.path.join(HERE, "golden_deep.json"), "w") as f:
        json.dump(dict(generator="oracle/fanlin_oracle_deep.c via tests/golden/make_golden_deep.py", pinned_to_reference=False, cases=out), f, indent=1)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
