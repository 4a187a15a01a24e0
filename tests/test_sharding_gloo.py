"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + all-reduce used by the multi-GPU
Monte Carlo path, with the CPU oracle standing in for the per-rank kernel (the compute itself is
covered by the -m gpu tests; there is no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from gokalman_b200.sharding import shard_range
    for total in (1, 7, 8, 1000, 10 ** 6 + 3):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, trials, steps, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import fixtures as fx
    from gokalman_b200.sharding import chisquare_means, shard_range
    from oracle import gko
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = fx.jerk3()
    lo, hi = shard_range(trials, rank, world)
    # this rank's shard: Philox keyed by the GLOBAL trial index (trial_offset = lo)
    r = gko.mc_chisquare(gko.VANILLA, f["F"], f["G"], f["H"], f["Q"], f["R"], f["x0"], f["x0"], f["P0"], hi - lo, steps,
                         seed=77, trial_offset=lo)
    nis, nees = chisquare_means(r["NIS"] * (hi - lo), r["NEES"] * (hi - lo), trials)
    out_q.put((rank, nis, nees))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fixtures as fx
    from oracle import gko
    gko.build()
    trials, steps, world = 301, 40, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, trials, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = fx.jerk3()
    full = gko.mc_chisquare(gko.VANILLA, f["F"], f["G"], f["H"], f["Q"], f["R"], f["x0"], f["x0"], f["P0"], trials, steps,
                            seed=77, trial_offset=0)
    for rank, nis, nees in results:
        assert fx.scaled_err(nis, full["NIS"]) <= 1e-12, rank
        assert fx.scaled_err(nees, full["NEES"]) <= 1e-12, rank
    # every rank holds the same global answer
    assert np.array_equal(results[0][1], results[1][1]) and np.array_equal(results[0][2], results[1][2])
