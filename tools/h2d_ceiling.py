#!/usr/bin/env python
"""The end-to-end roofline per N: pure cudaMemcpyAsync H2D (+ the matching D2H) from pinned host buffers on N GPUs of one
box, no kernels -- what fanlin_run's host-buffer path could reach at best.  Two modes:
    python tools/h2d_ceiling.py --gpus N            one process, one thread + stream per device (the product's layout)
    torchrun --nproc-per-node N tools/h2d_ceiling.py   N processes, one device each (bench.py's SCALE layout)
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--gb", type=float, default=8.0, help="GB copied per device and repetition")
ap.add_argument("--piece-mb", type=float, default=6.2208, help="bytes per copy call: one C2 image")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--pool", type=int, default=256, help="distinct pinned pieces cycled per device (bench.py's e2e leg cycles 1024 images = 6.4 GB)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
devices = [local] if world > 1 else list(range(args.gpus))
piece = int(args.piece_mb * 1e6)
n_pieces = max(1, int(args.gb * 1e9 / piece))
pool = min(n_pieces, args.pool)  # distinct pinned pieces cycled (256: 1.6 GB)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("gloo")


def worker(d, out):
    torch.cuda.set_device(d)
    host = torch.empty((pool, piece), dtype=torch.uint8, pin_memory=True)
    host.fill_(d + 1)
    dev = torch.empty((2, pool, piece), dtype=torch.uint8, device=f"cuda:{d}")
    back = torch.empty((pool, piece // 25), dtype=torch.uint8, pin_memory=True)  # results are ~4 % of the inputs (C2: 240 KB per 6.2 MB)
    st = torch.cuda.Stream(device=d)
    st_out = torch.cuda.Stream(device=d)  # results go back on a stream of their own: the link is full duplex
    best = 0.0
    worst_dt = 0.0  # slowest repetition's time with every device (and, under torchrun, every process) copying at once
    for rep in range(args.reps + 1):
        out["barrier"].wait()
        if world > 1:
            dist.barrier()  # all processes start the repetition together: the aggregate below is bytes / the slowest rank's time
        t0 = time.perf_counter()
        for i in range(n_pieces):
            with torch.cuda.stream(st):
                dev[i & 1, i % pool].copy_(host[i % pool], non_blocking=True)
            with torch.cuda.stream(st_out):
                back[i % pool].copy_(dev[i & 1, i % pool, : piece // 25], non_blocking=True)
        st.synchronize()
        st_out.synchronize()
        dt = time.perf_counter() - t0
        if rep:
            best = max(best, n_pieces * piece / dt / 1e9)
            out.setdefault("dts", {}).setdefault(rep, []).append(dt)
    out[d] = best


res = {"barrier": threading.Barrier(len(devices))}
th = [threading.Thread(target=worker, args=(d, res)) for d in devices]
[t.start() for t in th]
[t.join() for t in th]
per = [res[d] for d in devices]
# synchronised figure: per repetition, all bytes / the slowest device's time; the best repetition counts
rep_dt = torch.tensor([max(v) for _, v in sorted(res.get("dts", {}).items())], dtype=torch.float64)
if world > 1:
    t = torch.tensor([sum(per)], dtype=torch.float64)
    dist.all_reduce(t)
    dist.all_reduce(rep_dt, op=dist.ReduceOp.MAX)
    total, mode, n = float(t.item()), "processes", world
else:
    total, mode, n = sum(per), "threads", len(devices)
sync_total = n * n_pieces * piece / float(rep_dt.min().item()) / 1e9 if len(rep_dt) else 0.0
if rank == 0:
    print(json.dumps({"what": "H2D copy ceiling (pinned, one C2 image per copy, + 4 % D2H)", "mode": mode, "n_gpus": n, "aggregate_gb_s": total,
                      "per_gpu_gb_s": total / n, "out_mpix_s_at_c2": total * 1e9 / 6220800 * 0.06,
                      "synchronised": {"what": "every rank starts a repetition together; all bytes / the slowest rank's time (what a max-over-ranks bench can reach)",
                                       "aggregate_gb_s": sync_total, "per_gpu_gb_s": sync_total / n, "out_mpix_s_at_c2": sync_total * 1e9 / 6220800 * 0.06},
                      "pool_gb_per_gpu": pool * piece / 1e9}))
