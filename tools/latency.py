#!/usr/bin/env python
"""Single-request latency through fanlin_run (host buffers in, host buffers out), what one tokio
worker sees at src/main.rs:179: median / p95 over repeated calls, per shape."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import __graft_entry__ as G
    from synth import synth_image

    pkg = G.load_package()
    vp = int(sys.argv[sys.argv.index("--vertical-path") + 1]) if "--vertical-path" in sys.argv else 0
    dev = pkg.Device([0], vertical_path=vp)
    shapes = [("C1 512x512 RGB -> 300x200 fit+fill", 512, 512, 3, "w=300&h=200&rgb=32,32,32"),
              ("C2 1080p RGB -> 300x200 fit+fill", 1080, 1920, 3, "w=300&h=200"),
              ("C5 12MP RGB -> 1618x1000 crop+gray+blur", 3000, 4000, 3, "w=1618&h=1000&crop=true&grayscale=true&blur=10")]
    for name, h, w, c, qs in shapes:
        img = synth_image(1, h, w, c)
        q = pkg.Query(qs)
        for _ in range(5):
            pkg.process_image(dev, img, q)
        ts = []
        for _ in range(40):
            t0 = time.perf_counter()
            pkg.process_image(dev, img, q)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print(json.dumps(dict(vertical_path=vp, shape=name, median_ms=ts[len(ts) // 2], p95_ms=ts[int(len(ts) * 0.95)], min_ms=ts[0])), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
