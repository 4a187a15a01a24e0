"""Multi-GPU sharding of the Monte Carlo / chi-square path.

Trials are independent (montecarlo.go:108-117) and chi-square only averages across them
(chisquare.go:85-92), so the path shards with no data-path exchange: rank g of G runs the contiguous
trial range shard_range(T, g, G) with Philox keyed by the GLOBAL trial index, and the only
collective is one all-reduce (SUM) of the 2 x steps per-step NEES/NIS sums, after which every rank
divides by T.  torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(total, rank, world):
    """[lo, hi) of the trials rank `rank` owns: contiguous, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %d / world %d" % (rank, world))
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums(sums, group=None):
    """In-place SUM all-reduce of a torch tensor (any device) or numpy array; no-op when
    torch.distributed is not initialised (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    if isinstance(sums, np.ndarray):
        t = torch.from_numpy(sums)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
            dist.all_reduce(t, group=group)
            sums[...] = t.cpu().numpy()
        else:
            dist.all_reduce(t, group=group)
        return sums
    dist.all_reduce(sums, group=group)
    return sums


def chisquare_means(local_nis_sum, local_nees_sum, total_trials, group=None):
    """Per-step means over ALL ranks' trials from each rank's per-step sums (chisquare.go:85-92)."""
    stacked = np.stack([np.asarray(local_nis_sum, dtype=np.float64), np.asarray(local_nees_sum, dtype=np.float64)])
    stacked = allreduce_sums(stacked, group)
    return stacked[0] / float(total_trials), stacked[1] / float(total_trials)


def sharded_chisquare(make_filters, samples, steps, rowsH, controls, withNEES=True, withNIS=True, group=None):
    """NewMonteCarloRuns + NewChiSquare over all ranks of the default process group.

    make_filters(device) -> (mckf, chikf) builds this rank's pure predictor and tested filter on its
    GPU.  Returns the global (NISmeans, NEESmeans), identical on every rank."""
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(samples, rank, world)
    device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    mckf, chikf = make_filters(device)
    runs = api.NewMonteCarloRuns(hi - lo, steps, rowsH, controls, mckf, trial_offset=lo)
    nis, nees = api.NewChiSquare(chikf, runs, controls, withNEES, withNIS, sums=True)
    return chisquare_means(nis, nees, samples, group)
