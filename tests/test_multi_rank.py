"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch by image index
with the library's own partition function, plan their shards, and reduce the timing
the way bench.py does (max over ranks, aggregate = all units / that time).  The path
has no data-path collective: ranks only meet at the barrier and the
timing reduction."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_jobs, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as G
        import bench

        pkg = G.load_package()
        lo, hi = pkg.shard_range(n_jobs, world, rank)
        # plan this rank's shard (pure host): every image independent, same request
        qy = pkg.Query(bench.REQ)
        import ctypes as C

        j = pkg.Job()
        pkg.lib().fanlin_job_from_query(C.byref(qy._q), 0, C.byref(j))
        j.src_w, j.src_h, j.src_channels = bench.SRC_W, bench.SRC_H, bench.SRC_C
        alg = sum(pkg.plan_job(j).algorithmic_bytes for _ in range(lo, hi))
        # all ranks learn every shard: disjoint, contiguous, covering
        mine = torch.tensor([lo, hi, alg], dtype=torch.int64)
        allr = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        # timing reduction as in bench.py: device ms per rank -> max over ranks
        ms_local = 10.0 + 5.0 * rank
        ms, value = bench.reduce_timing(ms_local, steps=2, images_per_rank=hi - lo, world=world, dist=dist, device="cpu")
        dist.barrier()
        if rank == 0:
            q.put(([t.tolist() for t in allr], ms, value))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_jobs", [4096, 7])
def test_two_ranks_shard_by_image_index(n_jobs):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_jobs, q)) for r in range(world)]
    for p in procs:
        p.start()
    shards, ms, value = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shards[0][0] == 0 and shards[0][1] == shards[1][0] and shards[1][1] == n_jobs
    assert abs((shards[0][1] - shards[0][0]) - (shards[1][1] - shards[1][0])) <= 1
    assert shards[0][2] == 6460800 * (shards[0][1] - shards[0][0])
    assert ms == pytest.approx(15.0 / 2)  # max over ranks, per step
    import bench

    # rank 0 reports world x its own per-rank count (weak scaling) over the slowest rank's time
    assert value == pytest.approx(world * (shards[0][1] - shards[0][0]) * bench.MPIX_PER_IMAGE / 7.5e-3)


def test_shard_range_properties(fanlin):
    for n, k in [(0, 4), (1, 8), (8, 8), (4096, 8), (4097, 8), (5, 3)]:
        ranges = [fanlin.shard_range(n, k, s) for s in range(k)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
