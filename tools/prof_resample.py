#!/usr/bin/env python
"""Times the Lanczos3 resample alone (device-resident) on a BASELINE shape; with a -DT3_PROF / -DTC2_PROF build one CTA
prints where its roles spend their cycles.  Usage: python tools/prof_resample.py [c1|c2|c3|c5] [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G  # noqa: E402
import torch  # noqa: E402

SHAPES = {"c1": (512, 512, 3, "w=300&h=200&rgb=32,32,32"), "c2": (1080, 1920, 3, "w=300&h=200"), "c3": (2160, 3840, 4, "w=1618&h=1000&crop=true"),
          "c5": (3000, 4000, 1, "w=1618&h=1000&crop=true"),
          "c1crop": (512, 512, 3, "w=300&h=200&crop=true"), "c2crop": (1080, 1920, 3, "w=300&h=200&crop=true")}  # plain RGB out (the JPEG crop request)
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148
rgb8 = "--rgb8" in sys.argv  # FANLIN_TO_RGB8: the JPEG branch's layout from the kernel's epilogue
h, w, c, qs = SHAPES[which]
pkg = G.load_package()
dev = pkg.Device([0], vertical_path=3)
device = torch.device("cuda", 0)
src = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device=device)
q = pkg.Query(qs)
proto = pkg.Job()
pkg.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(proto))
proto.src_w, proto.src_h, proto.src_channels = w, h, c
if rgb8:
    proto.flags |= 1 << 5
plan = pkg.plan_job(proto)
dst = torch.zeros((n, plan.out_h, plan.out_w, plan.out_channels), dtype=torch.uint8, device=device)
jobs = (pkg.Job * n)()
for i in range(n):
    C.memmove(C.byref(jobs, i * C.sizeof(pkg.Job)), C.byref(proto), C.sizeof(pkg.Job))
    jobs[i].src = src.data_ptr() + i * h * w * c
    jobs[i].dst = dst.data_ptr() + i * plan.out_bytes
    jobs[i].dst_capacity = plan.out_bytes
torch.cuda.synchronize()
b = dev.prepare(jobs, 0)
b.set_timing(True)
for _ in range(3):
    b.launch(None)
torch.cuda.synchronize()
per = {}
for k, v in b.kernel_times():
    per.setdefault(k, []).append(v)
print(which + (" rgb8" if rgb8 else ""), f"batch {n}:", {k: round(min(v) * 1e3 / n, 2) for k, v in per.items()}, "us per image", flush=True)
b.free()
dev.close()
