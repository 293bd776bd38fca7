#!/usr/bin/env python
"""Soak run of the default device path on random requests (the generator of tests/test_fuzz_gpu.py with other seeds, plus
larger images so that the both-passes tensor-core kernels see many band / ring / region geometries), single requests and
small ragged batches, against the oracle.  Usage: python tools/soak.py [n_requests] [seed]   (wrap it in `timeout`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from synth import synth_image  # noqa: E402
from test_fuzz_gpu import _case, _qs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(900000 + seed)
pkg = G.load_package()
dev = pkg.Device([0])
bad = d1 = vals = skipped = 0
i = 0
while i < n:
    batch = []
    for _ in range(int(rng.choice([1, 1, 1, 3, 9]))):
        h, w, c, p, exif = _case(rng)
        if rng.random() < 0.3:  # larger sources: strong and mild downscales through the tensor-core kernels
            h, w = int(rng.integers(300, 1400)), int(rng.integers(300, 2000))
        batch.append((synth_image(int(rng.integers(0, 1 << 20)), h, w, c), p, exif))
    jobs = [pkg.make_job(im, pkg.Query(_qs(p)), orientation=e) for im, p, e in batch]
    try:
        outs = pkg.stage._run(dev, jobs)
    except pkg.FanlinError as ex:  # a request the planner rejects (resized dimensions beyond 65535): not a device matter
        if "too large" not in str(ex):
            raise
        skipped += len(batch)
        i += len(batch)
        continue
    for (im, p, e), got in zip(batch, outs):
        want = O.process(im, orientation=e, **{k: v for k, v in p.items()})
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        if got.shape != want.shape or d.max() > 1:
            bad += 1
            print("MISMATCH", im.shape, p, e, got.shape, want.shape, int(d.max()) if got.shape == want.shape else None, flush=True)
        d1 += int((d == 1).sum()); vals += d.size
        i += 1
print(f"soak seed {seed}: {i} requests ({skipped} rejected by the planner), {bad} mismatches, {d1} of {vals} values off by one ({100.0 * d1 / max(vals, 1):.4f} %)", flush=True)
dev.close()
sys.exit(1 if bad else 0)
