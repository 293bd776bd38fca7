#!/usr/bin/env python
"""The tensor-core blur (both passes, kernels_blur_tc.cu) against the oracle on shapes that exercise its borders,
bands, ring wrap and channel counts; prints the |diff| histogram and the kernels that ran.  GPU only."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from synth import synth_image  # noqa: E402
import torch  # noqa: E402

pkg = G.load_package()
O.build()
CASES = [  # (h, w, c, sigma)
    (200, 300, 4, 10.0), (128, 64, 1, 10.0), (129, 257, 3, 10.0), (64, 48, 2, 10.0), (300, 200, 4, 12.0), (1000, 1618, 4, 10.0),
    (1000, 1618, 1, 10.0), (1000, 1618, 1, 20.0), (333, 1000, 3, 15.0), (40, 30, 4, 10.0), (500, 777, 1, 13.3), (257, 129, 2, 20.0),
    (97, 131, 3, 0.8), (97, 131, 4, 1.3),
]
if len(sys.argv) > 1:
    CASES = CASES[: int(sys.argv[1])]
dev = pkg.Device([0])
device = torch.device("cuda", 0)
bad = 0
for (h, w, c, sigma) in CASES:
    img = synth_image(11 + h + c, h, w, c)
    want = O.process(img, blur=float(np.float32(sigma)))
    pitch = (w * c + 15) // 16 * 16
    src = torch.zeros((h, pitch), dtype=torch.uint8, device=device)
    src[:, : w * c] = torch.from_numpy(img.reshape(h, w * c)).to(device)
    src[:, w * c:] = 0xAB  # pitch padding must not leak into the result
    dst = torch.zeros((h, w, c), dtype=torch.uint8, device=device)
    j = pkg.Job()
    q = pkg.Query("")
    pkg.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(j))
    j.src, j.src_w, j.src_h, j.src_channels, j.src_pitch = src.data_ptr(), w, h, c, pitch
    j.blur_sigma = sigma
    j.dst, j.dst_capacity = dst.data_ptr(), h * w * c
    torch.cuda.synchronize()
    b = dev.prepare([j], 0)
    b.set_timing(True)
    b.launch(None)
    torch.cuda.synchronize()
    names = [k for k, _ in b.kernel_times()]
    got = dst.cpu().numpy()
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    ys, xs = np.nonzero(d.max(axis=2) >= 2)
    where = f" first bad at y={ys[0]} x={xs[0]} got={got[ys[0], xs[0]]} want={want[ys[0], xs[0]]}" if len(ys) else ""
    print(f"{h}x{w}x{c} sigma={sigma}: d1={int((d == 1).sum())} d2+={int((d >= 2).sum())} max={int(d.max())} kernels={names}{where}", flush=True)
    bad += int((d >= 2).sum())
    b.free()
print("FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
