#!/usr/bin/env python
"""Times the blur alone (device-resident, blur-only jobs) on the shapes of the BASELINE configs.  With a -DBT_PROF build
one CTA prints where its consumer warps spend their cycles.  Usage: python tools/prof_blur.py [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
import torch  # noqa: E402

pkg = G.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = pkg.Device([0])
device = torch.device("cuda", 0)
for (h, w, c, sigma) in [(1000, 1618, 4, 10.0), (1000, 1618, 1, 10.0), (1000, 1618, 3, 10.0), (1000, 1617, 1, 10.0), (1000, 1617, 3, 10.0)]:  # (1617: rows at odd addresses)
    pitch = (w * c + 15) // 16 * 16
    src = torch.randint(0, 256, (n, h, pitch), dtype=torch.uint8, device=device)
    dst = torch.zeros((n, h, w * c), dtype=torch.uint8, device=device)
    jobs = (pkg.Job * n)()
    q = pkg.Query("")
    for i in range(n):
        pkg.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(jobs[i]))
        jobs[i].src, jobs[i].src_w, jobs[i].src_h, jobs[i].src_channels, jobs[i].src_pitch = src.data_ptr() + i * h * pitch, w, h, c, pitch
        jobs[i].blur_sigma = sigma
        jobs[i].dst, jobs[i].dst_capacity = dst.data_ptr() + i * h * w * c, h * w * c
    torch.cuda.synchronize()
    b = dev.prepare(jobs, 0)
    b.set_timing(True)
    for _ in range(3):
        b.launch(None)
    torch.cuda.synchronize()
    kt = b.kernel_times()
    per = {}
    for k, v in kt:
        per.setdefault(k, []).append(v)
    print(f"{h}x{w}x{c} sigma={sigma} batch {n}:", {k: round(min(v) * 1e3 / n, 2) for k, v in per.items()}, "us per image", flush=True)
    b.free()
dev.close()
