#!/usr/bin/env python
"""Both-passes-on-the-tensor-cores kernel against the oracle and against the other device paths on a
few shapes; prints the |diff| histogram and the kernels that ran.  GPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from synth import synth_image  # noqa: E402

pkg = G.load_package()
O.build()
CASES = [  # (h, w, c, query)
    (1080, 1920, 3, "w=300&h=200"),
    (1080, 1920, 3, "w=300&h=200&crop=true"),
    (1000, 1777, 3, "w=211&h=160"),
    (2160, 3840, 3, "w=400&h=300"),
    (700, 2000, 3, "w=150&h=150"),
    (1080, 1920, 4, "w=300&h=200"),
    (3000, 4000, 1, "w=1333&h=1000"),
    (1500, 2000, 2, "w=300&h=300"),
    (512, 512, 3, "w=300&h=200"),
]
dev = {vp: pkg.Device([0], vertical_path=vp) for vp in (3, 2)}
bad = 0
for (h, w, c, q) in CASES:
    img = synth_image(77 + h + c, h, w, c)
    want = O.process(img, **{k: (v == "true" if v in ("true", "false") else int(v)) for k, v in (kv.split("=") for kv in q.split("&"))})
    line = f"{h}x{w}x{c} {q:28s}"
    for vp in (3, 2):
        got = pkg.process_image(dev[vp], img, pkg.Query(q))
        # which kernel: prepare a device batch with timing
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        line += f" | path {vp}: d1={int((d == 1).sum()):6d} d2+={int((d >= 2).sum()):6d} max={int(d.max())}"
        if vp == 3:
            bad += int((d >= 2).sum())
    print(line, flush=True)
print("FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
