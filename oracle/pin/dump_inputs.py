#!/usr/bin/env python
"""Writes the inputs of the golden cases (tests/golden/golden.json, and golden_deep.json: 16-bit and f32 images) as raw
pixel files + manifest.json for oracle/pin (the Rust program that runs them through image = 0.25.6).
Usage: python oracle/pin/dump_inputs.py <dir>"""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synth_deep, synth_image  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]
lenna = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "lenna_512_rgb.png")).convert("RGB"))
manifest = []
for c in golden:
    img = lenna if c["input"] == "lenna" else synth_image(*c["input"])
    img = np.ascontiguousarray(img if img.ndim == 3 else img[:, :, None])
    img.tofile(os.path.join(out, c["name"] + ".raw"))
    manifest.append({"name": c["name"], "width": img.shape[1], "height": img.shape[0], "channels": img.shape[2], "sample": "u8", "params": c["params"]})
deep = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_deep.json")))["cases"]
for c in deep:  # subpixels in native (little-endian) byte order, as ImageBuffer::as_raw() holds them
    seed, h, w, ch, dt = c["input"]
    img = np.ascontiguousarray(synth_deep(seed, h, w, ch, np.dtype(dt)))
    img.tofile(os.path.join(out, c["name"] + ".raw"))
    manifest.append({"name": c["name"], "width": w, "height": h, "channels": ch, "sample": {"uint16": "u16", "float32": "f32"}[dt], "params": c["params"]})
json.dump(manifest, open(os.path.join(out, "manifest.json"), "w"), indent=1)
print(f"{len(manifest)} inputs in {out}")
