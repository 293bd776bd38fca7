#!/usr/bin/env python
"""Benchmark of the pixel-transform stage (BASELINE.json metric: output Mpix/s of the
fused resize+fill(+blur) pipeline, and % of HBM peak).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- a batch of 4096 synthetic
1920x1080 RGB images -> w=300&h=200 aspect-fit + letterbox fill, per GPU (weak
scaling: every rank owns its own 4096 images; images are independent, so there is
no collective on the data path).  A step is one pass of the stage over the batch.

  value   device-resident throughput (inputs in HBM before the timed region),
          CUDA events on the launching stream, max over ranks.
  e2e     the same metric through the C ABI's blocking host entry point
          (fanlin_run) with pinned HOST buffers: H2D of the inputs and D2H of the
          results inside the timed region.
  roofline  dominant kernel: algorithmic bytes per launch / its average device
          time (CUDA events bracketing the kernel), against MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (C restatement of the image-crate path the
          reference calls) on the host cores, on a bounded sample.
`--impl reference` times that CPU path alone, as the reference arm.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SRC_W, SRC_H, SRC_C = 1920, 1080, 3
REQ = "w=300&h=200"
OUT_W, OUT_H, OUT_C = 300, 200, 4
BATCH = 4096
MPIX_PER_IMAGE = OUT_W * OUT_H / 1e6
WORKLOAD = "C2: 4096 x 1920x1080 RGB -> w=300&h=200 fit + letterbox fill (RGBA8 300x200)"
METRIC = "output Mpix/s, fused resize+fill+blur pipeline, 1/2/4/8 B200; % of HBM peak"  # BASELINE.json's string
# what the dominant kernel computes in: vertical pass u8 x s8 (three base-128 digits of the 2^-21-quantised weights) -> s32 on
# tcgen05 kind::i8, recombined exactly in f32; horizontal pass f16 hi/lo x f16 hi/lo -> f32 on tcgen05 kind::f16; u8 out
DTYPE = "u8*s8->s32 (vertical, exact) + f16 hi/lo*f16 hi/lo->f32 (horizontal), u8 out"
DTYPE_CPU = "f32 (image-crate operation order)"

# The other BASELINE.json configs, device-resident, parity-checked, reported under the line's `configs` key.
#   key, BASELINE config, h, w, c, query, gif, images per GPU and launch, total images of a strong-scaling run (or None)
OTHER_CONFIGS = [
    ("C1", "configs[0] shape: 512x512 RGB -> w=300&h=200&rgb=32,32,32 (fit + letterbox), batch 1024", 512, 512, 3, "w=300&h=200&rgb=32,32,32", False, 1024, None),
    ("C3_resize_crop", "configs[2] without blur: 3840x2160 RGBA -> w=1618&h=1000&crop=true, batch 1024", 2160, 3840, 4, "w=1618&h=1000&crop=true", False, 1024, None),
    ("C3", "configs[2]: 3840x2160 RGBA -> w=1618&h=1000&crop=true&blur=10, batch 1024", 2160, 3840, 4, "w=1618&h=1000&crop=true&blur=10", False, 1024, None),
    ("C4", "configs[3] literal: 200 GIF frames 480x270 RGBA, w=200&grayscale=true&inverse=true (no h: no resize; grayscale wins)", 270, 480, 4,
     "w=200&grayscale=true&inverse=true", True, 200, None),
    ("C4_h113", "configs[3] with h: 200 GIF frames 480x270 RGBA, w=200&h=113&grayscale=true&inverse=true (Nearest)", 270, 480, 4,
     "w=200&h=113&grayscale=true&inverse=true", True, 200, None),
    ("C5_crop", "configs[4] crop variant: 8192 x 4000x3000 RGB -> w=1618&h=1000&crop=true&grayscale=true&blur=10 (L8 out), resident chunks of <= 2048 per GPU",
     3000, 4000, 3, "w=1618&h=1000&crop=true&grayscale=true&blur=10", False, 2048, 8192),
    ("C5_fit", "configs[4] fit + fill variant: 8192 x 4000x3000 RGB -> w=1618&h=1000&rgb=32,32,32&grayscale=true&blur=10 (RGBA8 out), resident chunks of <= 2048 per GPU",
     3000, 4000, 3, "w=1618&h=1000&rgb=32,32,32&grayscale=true&blur=10", False, 2048, 8192),
]
FP32_FMA_PER_S = 148 * 128 * 1.965e9  # CUDA-core FMA peak (the direct-form blur is FP32-bound)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.p = gpu_index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pw.append(0.0)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = those drawing at least 60 % of the highest power seen (the sampler also runs through the
        # idle stretches around the warm-up and the timed region; idle SM clocks sit at the maximum on this part)
        top = max(pw) if pw else 0.0
        load = [c for c, w in zip(sm, pw) if w >= 0.6 * top] if top > 0 else sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "sm_mhz_min_under_load": min(load) if load else None, "power_w_max": top or None}


def synth_batch_on_device(torch, n, seed, device):
    """u8 noise blended 50/50 with a smooth gradient, generated on the device."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, SRC_H, SRC_W, SRC_C), dtype=torch.uint8, device=device)
    yy = torch.arange(SRC_H, device=device, dtype=torch.int32).view(SRC_H, 1, 1)
    xx = torch.arange(SRC_W, device=device, dtype=torch.int32).view(1, SRC_W, 1)
    k = torch.arange(SRC_C, device=device, dtype=torch.int32).view(1, 1, SRC_C)
    grad = (xx * (k + 1) * 255 // (SRC_W - 1) + yy * (3 - k % 3) * 255 // (SRC_H - 1)) % 511
    grad = torch.where(grad > 255, 510 - grad, grad).to(torch.int16)
    chunk = 64
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        noise = torch.randint(0, 256, (m, SRC_H, SRC_W, SRC_C), dtype=torch.int16, device=device, generator=g)
        out[i:i + m] = ((noise + grad + 1) // 2).to(torch.uint8)
    return out


def cpu_baseline(n_images, threads, steps=1):
    """The CPU oracle on `n_images` C2 images over `threads` host threads; returns (Mpix/s, seconds/step)."""
    import numpy as np
    from oracle import oracle as O
    from synth import synth_image

    O.build()
    base = [synth_image(2000 + i, SRC_H, SRC_W, SRC_C) for i in range(min(n_images, 8))]
    imgs = [base[i % len(base)] for i in range(n_images)]
    O.process_batch(imgs[:threads], n_threads=threads, w=OUT_W, h=OUT_H)  # warm-up (page faults, thread start)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        outs = O.process_batch(imgs, n_threads=threads, w=OUT_W, h=OUT_H)
        ts.append(time.perf_counter() - t0)
    assert outs[0].shape == (OUT_H, OUT_W, OUT_C)
    dt = sum(ts) / len(ts)
    return n_images * MPIX_PER_IMAGE / dt, dt


def config_dict(images_per_gpu, exact=False):
    """`config` of the JSON line: the same keys and values on both arms."""
    return {"workload": WORKLOAD, "images_per_gpu": images_per_gpu, "l2": "inputs (25.5 GB/GPU) larger than L2",
            "sharding": "by image index, no collective", "exact_mode": bool(exact)}


def run_other_configs(args, pkg, dev, torch, np, device, stream, peak, world, dist, rank):
    """Device-timed, parity-checked record of every other BASELINE.json config (what tools/bench_configs.py prints),
    at the stated batch sizes; C5 as a strong-scaling run of 8192 images over the ranks in resident chunks."""
    from oracle import oracle as O
    from synth import synth_image

    out = {}
    for key, desc, h, w, c, qs, gif, per_launch, total in OTHER_CONFIGS:
        if args.only_configs and key not in args.only_configs.split(","):
            continue
        n_rank = per_launch if total is None else max(1, total // world)   # images this rank owns per step
        n = min(per_launch, n_rank)                                         # resident chunk: images per launch
        launches = (n_rank + n - 1) // n
        base = [torch.from_numpy(synth_image(900 + 10 * rank + i, h, w, c)).to(device) for i in range(4)]
        src = torch.empty((n, h, w, c), dtype=torch.uint8, device=device)
        for i in range(4):
            src[i::4] = base[i]
        q = pkg.Query(qs)
        proto = pkg.Job()
        pkg.lib().fanlin_job_from_query(C.byref(q._q), int(gif), C.byref(proto))
        proto.src_w, proto.src_h, proto.src_channels = w, h, c
        plan = pkg.plan_job(proto)
        dst = torch.zeros((n, plan.out_h, plan.out_w, plan.out_channels), dtype=torch.uint8, device=device)
        jobs = (pkg.Job * n)()
        for i in range(n):
            C.memmove(C.byref(jobs, i * C.sizeof(pkg.Job)), C.byref(proto), C.sizeof(pkg.Job))
            jobs[i].src = src.data_ptr() + i * h * w * c
            jobs[i].dst = dst.data_ptr() + i * plan.out_bytes
            jobs[i].dst_capacity = plan.out_bytes
        torch.cuda.synchronize(device)
        b = dev.prepare(jobs, 0)
        b.set_timing(True)
        for _ in range(3):
            b.launch(stream.cuda_stream)
        torch.cuda.synchronize(device)
        b.kernel_times()
        steps = max(1, min(args.steps, 5))
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps * launches):
            b.launch(stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize(device)
        kt = {}
        for k, v in b.kernel_times():
            kt.setdefault(k, []).append(v)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if dist and world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps                      # one step = this rank's n_rank images (launches x n)
        imgs_step = world * n * launches
        alg = plan.algorithmic_bytes * n * launches
        rec = {"workload": desc, "query": qs, "images_per_gpu_per_launch": n, "launches_per_step": launches, "images_per_step_all_gpus": imgs_step,
               "steps": steps, "ms_per_step": ms, "us_per_image": ms * 1e3 / (n * launches),
               "out_mpix_s": imgs_step * plan.out_w * plan.out_h / 1e6 / (ms * 1e-3),
               "algorithmic_bytes_per_image": int(plan.algorithmic_bytes), "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
               # summed over the kernel instances of one batch launch (a batch whose scratch is chunked launches a kernel several times)
               "kernels_ms_per_launch": {k: sum(v) / (steps * launches) for k, v in kt.items()},
               "scaling": "weak" if total is None else f"strong: {total} images over {world} GPU(s)"}
        if q.blur() and not gif:
            # the direct-form blur is FP32-bound -- report the blur kernels against the CUDA-core FMA peak too
            # (2 passes x taps FMAs per output element), whatever unit actually runs them
            taps = 2 * int(np.ceil(2 * q.blur() - 0.5)) + 1
            fma = 2.0 * taps * plan.out_w * plan.out_h * plan.out_channels * n
            blur_ms = sum(v for k, v in rec["kernels_ms_per_launch"].items() if "blur" in k)
            if blur_ms > 0:
                rec["blur"] = {"taps": taps, "ms_per_launch": blur_ms, "fp32_frac": fma / (blur_ms * 1e-3) / FP32_FMA_PER_S}
        if rank == 0:  # parity of two images of the batch (one per distinct base image) at full size
            kw = dict(grayscale=q.grayscale(), inverse=q.inverse(), crop=q.cropping(), blur=0.0 if gif else q.blur(), rgb=q.fill_color(), gif=gif)
            if q.dimensions():
                kw["w"], kw["h"] = q.dimensions()
            d1 = d2 = nn = 0
            for i in (1, n - 1):
                want = O.process(base[i % 4].cpu().numpy(), **kw)
                d = np.abs(dst[i].cpu().numpy().astype(np.int16) - want.astype(np.int16))
                d1 += int((d == 1).sum()); d2 += int((d >= 2).sum()); nn += int(d.size)
            rec["parity"] = {"images_checked": 2, "values": nn, "diff1": d1, "diff_ge2": d2}
            assert d2 == 0, (key, rec["parity"])
        out[key] = rec
        b.free()
        del src, dst, base
        torch.cuda.empty_cache()
    return out


def reduce_timing(ms_total_local, steps, images_per_rank, world, dist, device):
    """Max over ranks of the device time, and the whole-job aggregate it implies (weak scaling:
    every rank processed images_per_rank per step).  Returns (ms_per_step, Mpix/s)."""
    import torch

    t = torch.tensor([ms_total_local], dtype=torch.float64, device=device)
    if dist is not None and world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    return ms_per_step, world * images_per_rank * MPIX_PER_IMAGE / (ms_per_step * 1e-3)


def run_reference(args, rank, out_fd):
    """Reference arm: the reference's CPU implementation of the path (the oracle port: the
    image crate cannot be built here) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(16 * cores, 64)
    for _ in range(args.warmup):
        cpu_baseline(cores, cores)
    v, dt = cpu_baseline(n, cores, steps=args.steps)
    sample = f"{n} of the {BATCH} C2 images per step, one image per thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_CPU, "data": "synthetic",
        "config": config_dict(BATCH), "sample": sample,
        "cpu_baseline": {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line, out_fd)


def emit(line, out_fd):
    """The one JSON line of the run, alone on the process's real stdout."""
    os.write(out_fd, (json.dumps(line) + "\n").encode())


def main():
    # Everything any library prints to fd 1 from here on (NCCL prints its version banner there under torchrun when
    # NCCL_DEBUG is set) goes to stderr: stdout carries the JSON line and nothing else.
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step (default: the C2 batch)")
    ap.add_argument("--e2e-pool", type=int, default=1024, help="distinct pinned host images cycled by the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle spot check (kernel ablation runs only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the records of the other BASELINE configs")
    ap.add_argument("--only-configs", default="", help="comma-separated keys of OTHER_CONFIGS to run")
    ap.add_argument("--exact", action="store_true", help="bit-exact kernels (crate operation order)")
    ap.add_argument("--vertical-path", type=int, default=0, help="0: tensor cores (default), 1: CUDA cores, 2: tensor cores for the vertical pass only, 3: both passes whatever the batch size")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, out_fd)
        return

    import numpy as np
    import torch
    import __graft_entry__ as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    pkg = G.load_package()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)
        host_group = dist.new_group(backend="gloo")  # CPU-side barrier for the leg in which rank 0 drives every GPU itself
    peak, peak_src = load_peaks()
    warm = max(args.warmup, 3)

    dev = pkg.Device([local_rank], exact=args.exact, vertical_path=args.vertical_path)
    n = args.batch
    src = synth_batch_on_device(torch, n, 2000 + rank, device)
    dst = torch.zeros((n, OUT_H, OUT_W, OUT_C), dtype=torch.uint8, device=device)
    q = pkg.Query(REQ)
    jobs = (pkg.Job * n)()
    img_bytes, out_bytes = SRC_H * SRC_W * SRC_C, OUT_H * OUT_W * OUT_C
    proto = pkg.Job()
    pkg.lib().fanlin_job_from_query(C.byref(q._q), 0, C.byref(proto))
    for i in range(n):
        C.memmove(C.byref(jobs, i * C.sizeof(pkg.Job)), C.byref(proto), C.sizeof(pkg.Job))
        jobs[i].src = src.data_ptr() + i * img_bytes
        jobs[i].src_w, jobs[i].src_h, jobs[i].src_channels = SRC_W, SRC_H, SRC_C
        jobs[i].dst = dst.data_ptr() + i * out_bytes
        jobs[i].dst_capacity = out_bytes
    batch = dev.prepare(jobs, 0)
    alg_bytes = sum(p.algorithmic_bytes for p in batch.plans)
    assert batch.plans[0].out_w == OUT_W and batch.plans[0].out_h == OUT_H and batch.plans[0].out_channels == OUT_C
    stream = torch.cuda.Stream(device)  # a real handle: NULL would select the library's own stream
    assert stream.cuda_stream != 0

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident leg ------------------------------------------------------------
    batch.set_timing(True)
    # the clock sampler starts before the warm-up: nvidia-smi needs a few hundred ms before its first line, and the timed
    # region is tens of ms (its idle samples are dropped in stop(): the upper half of the SM clocks is what ran under load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.5)
    for _ in range(warm):
        batch.launch(stream.cuda_stream)
    torch.cuda.synchronize(device)
    batch.kernel_times()  # drop the warm-up record
    ktimes = {}
    launches0 = dev.stats()["kernel_launches"]
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    stream.synchronize()
    ev[0].record(stream)
    for s in range(args.steps):
        batch.launch(stream.cuda_stream)
        ev[s + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    for name, ms in batch.kernel_times():  # every kernel of the K timed launches, bracketed by events
        ktimes.setdefault(name, []).append(ms)
    total_ms = ev[0].elapsed_time(ev[-1])
    gpu_launches = dev.stats()["kernel_launches"] - launches0
    ms_per_step, value = reduce_timing(total_ms, args.steps, n, world, dist, device)

    # parity spot check at full size: a few images of the batch against the oracle
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import oracle as O

        d2 = 0
        d1 = 0
        for i in (0, n // 2, n - 1):
            want = O.process(src[i].cpu().numpy(), w=OUT_W, h=OUT_H)
            got = dst[i].cpu().numpy()
            d = np.abs(got.astype(np.int16) - want.astype(np.int16))
            d1 += int((d == 1).sum()); d2 += int((d >= 2).sum())
        parity = {"images_checked": 3, "diff1": d1, "diff_ge2": d2}
        assert d2 == 0, parity

    # dominant kernel and its roofline
    def ncu_traffic(kernel, images):
        """DRAM bytes per launch of the dominant kernel from this round's committed `ncu --set full` capture of the
        same workload (profiles/r02/ncu_traffic.json), scaled from the captured launch's image count to this launch's:
        the kernel streams every image once, traffic is per image.  ncu cannot run inside a timed bench."""
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02", "ncu_traffic.json")
        try:
            t = json.load(open(path)).get(kernel)
        except OSError:
            return None
        if not t or not t["workload"].startswith("C2"):
            return None
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["images"] * images

    dom = max(ktimes, key=lambda k: sum(ktimes[k])) if ktimes else None
    roofline = None
    if dom:
        avg_ms = sum(ktimes[dom]) / len(ktimes[dom])
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None if args.exact else ncu_traffic(dom, n),
                    "traffic_source": "profiles/r02/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch of this kernel on this workload, per image x images",
                    "peak_source": peak_src,
                    "kernel_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                    "kernel_share_of_step": avg_ms / ms_per_step,
                    "all_kernels_ms": {k: sum(v) / len(v) for k, v in ktimes.items()}}
    batch.set_timing(False)

    # ---- end-to-end leg: pinned host buffers through fanlin_run ------------------------
    e2e = None
    if not args.no_e2e:
        pool = min(args.e2e_pool, n)
        hin = dev.host_alloc(pool * img_bytes)
        hout = dev.host_alloc(n * out_bytes)
        hin_t = torch.from_numpy(hin).view(pool, SRC_H, SRC_W, SRC_C)
        hin_t.copy_(src[:pool])  # same synthetic images, now host-resident
        hjobs = (pkg.Job * n)()
        for i in range(n):
            C.memmove(C.byref(hjobs, i * C.sizeof(pkg.Job)), C.byref(proto), C.sizeof(pkg.Job))
            hjobs[i].src = hin.ctypes.data + (i % pool) * img_bytes
            hjobs[i].src_w, hjobs[i].src_h, hjobs[i].src_channels = SRC_W, SRC_H, SRC_C
            hjobs[i].dst = hout.ctypes.data + i * out_bytes
            hjobs[i].dst_capacity = out_bytes
        e_steps = max(1, min(args.steps, 5))
        e_warm = 1
        for _ in range(e_warm):
            dev.run(hjobs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            dev.run(hjobs)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * MPIX_PER_IMAGE * e_steps / dt, "unit": "Mpix/s",
               "h2d_bytes_per_step": n * img_bytes, "d2h_bytes_per_step": n * out_bytes,
               "steps": e_steps, "ms_per_step": dt / e_steps * 1e3, "host_pool_images": pool,
               "api": "fanlin_run (C ABI), pinned host buffers from fanlin_host_alloc"}
        # every image of every sub-batch against the device-resident leg's result for the same source image
        dref = dst[:pool].cpu().numpy().reshape(pool, out_bytes)
        hall = hout[:n * out_bytes].reshape(n, out_bytes)
        bad = [i for i in range(n) if not np.array_equal(hall[i], dref[i % pool])]
        assert not bad, f"e2e output differs from the device-resident leg for {len(bad)} images, first {bad[:5]}"
        e2e["images_compared_with_device_leg"] = n
        # ---- the same leg from PAGEABLE host buffers (what the reference's decoders hand over: a Vec<u8> per image,
        # src/handler.rs:219): the library stages them through its own pinned buffers, sub-batch by sub-batch
        if world == 1:  # (N = 1 only: N ranks x 16 copy threads would measure the host's memory system, not the library)
            pin = np.empty(pool * img_bytes, np.uint8)
            pin[:] = hin[:pool * img_bytes]
            pout = np.empty(n * out_bytes, np.uint8)
            pjobs = (pkg.Job * n)()
            for i in range(n):
                C.memmove(C.byref(pjobs, i * C.sizeof(pkg.Job)), C.byref(hjobs, i * C.sizeof(pkg.Job)), C.sizeof(pkg.Job))
                pjobs[i].src = pin.ctypes.data + (i % pool) * img_bytes
                pjobs[i].dst = pout.ctypes.data + i * out_bytes
            dev.run(pjobs)  # warm-up: the staging buffers enter the pinned pool
            barrier()
            p_steps = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(p_steps):
                dev.run(pjobs)
            torch.cuda.synchronize(device)
            tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if dist:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            dtp = float(tp.item())
            assert np.array_equal(pout.reshape(n, out_bytes)[n // 2], dref[(n // 2) % pool]) and np.array_equal(pout.reshape(n, out_bytes)[n - 1], dref[(n - 1) % pool])
            e2e["pageable"] = {"value": world * n * MPIX_PER_IMAGE * p_steps / dtp, "unit": "Mpix/s", "steps": p_steps, "ms_per_step": dtp / p_steps * 1e3,
                               "frac_of_pinned": (world * n * MPIX_PER_IMAGE * p_steps / dtp) / e2e["value"],
                               "api": "fanlin_run (C ABI), pageable host buffers (numpy arrays): staged through the library's pinned pool"}
            del pin, pout
        # ---- the product's own multi-GPU path: ONE process, ONE context over all N devices, one fanlin_run call that shards
        # the N x 4096 images by index (one host thread + streams per device, src/main.rs keeps a single State); the other
        # ranks wait at the barrier.  Same pinned pool, outputs checked against the device-resident leg.
        if world > 1:
            barrier()
            if rank == 0:
                devn = pkg.Device(list(range(world)), exact=args.exact, vertical_path=args.vertical_path)
                nn = world * n
                hout_n = devn.host_alloc(nn * out_bytes)
                njobs = (pkg.Job * nn)()
                for i in range(nn):
                    C.memmove(C.byref(njobs, i * C.sizeof(pkg.Job)), C.byref(hjobs, (i % n) * C.sizeof(pkg.Job)), C.sizeof(pkg.Job))
                    njobs[i].dst = hout_n.ctypes.data + i * out_bytes
                devn.run(njobs)  # warm-up (contexts, pools, tables on every device)
                n_steps = max(1, min(e_steps, 3))
                t0 = time.perf_counter()
                for _ in range(n_steps):
                    devn.run(njobs)
                dtn = time.perf_counter() - t0
                alln = hout_n[:nn * out_bytes].reshape(nn, out_bytes)
                badn = [i for i in range(0, nn, 7) if not np.array_equal(alln[i], dref[(i % n) % pool])]
                assert not badn, f"single-process multi-GPU output differs for images {badn[:5]}"
                e2e["single_process"] = {"value": nn * MPIX_PER_IMAGE * n_steps / dtn, "unit": "Mpix/s", "devices": devn.device_count, "images_per_step": nn,
                                         "steps": n_steps, "ms_per_step": dtn / n_steps * 1e3,
                                         "frac_of_n_processes": (nn * MPIX_PER_IMAGE * n_steps / dtn) / e2e["value"],
                                         "api": "one fanlin_ctx over all devices, one fanlin_run per step (shards by image index, one host thread per device)"}
                devn.host_free(hout_n)
                devn.close()
            dist.barrier(group=host_group)  # on the host: no collective kernel spins on a GPU rank 0 is using
        dev.host_free(hin)
        dev.host_free(hout)

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = max(64 * cores, 256)  # ~25 s of CPU work: one image takes ~25 ms on one core
        v, dt = cpu_baseline(ns, cores)
        cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{ns} of the {n} C2 images, one image per thread, {dt:.2f} s"}

    batch.free()
    del src, dst
    torch.cuda.empty_cache()
    configs = None
    if not args.no_configs and not args.exact:
        configs = run_other_configs(args, pkg, dev, torch, np, device, stream, peak, world, dist, rank)
    dev.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE if not args.exact else DTYPE_CPU, "data": "synthetic",
            "config": config_dict(n, args.exact), "hbm_frac_whole_step": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "clocks": clocks, "parity": parity, "configs": configs,
        }
        emit(line, out_fd)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
