"""gkb_mc_chisquare_multi: the Monte Carlo + chi-square run sharded over several GPUs of ONE process with the
collective inside the C-ABI (what a cgo caller uses: no torch.distributed).  On a 1-GPU box the PEER mode accepts
the same device for every shard, which exercises the sharding, the global-trial Philox keying and the rank-ordered
reduction; with >= 2 GPUs (gpurun --gpus 2) the NCCL all-reduce and the NVLink peer loads run for real."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


def _setup(trials, steps, devices, reduce, seed=77):
    import gokalman_b200 as gk
    gk.load()
    f = fx.robot_1d()
    controls = list(fx.robot_controls(steps))
    mckf, _ = gk.NewPurePredictorVanilla(np.array([0.7, -0.4]), f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=seed))
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(0.3 * f["Q"], f["R"]))
    runs = gk.NewMonteCarloRuns(trials, steps, 1, controls, mckf, devices=devices, reduce=reduce)
    return gk, kf, runs, controls


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_multi_same_device_shards_equal_single_run():
    trials, steps = 5003, 300  # ragged shards, more than one 256-step chunk
    gk, kf, runs1, controls = _setup(trials, steps, None, "peer")
    nis1, nees1 = gk.NewChiSquare(kf, runs1, controls, True, True)
    for shards in (2, 3):
        gk, kf, runs, controls = _setup(trials, steps, [0] * shards, "peer")
        nis, nees = gk.NewChiSquare(kf, runs, controls, True, True)
        assert fx.scaled_err(nis, nis1) <= 1e-13 and fx.scaled_err(nees, nees1) <= 1e-13
        for k in (0, 100, steps - 1):  # Mean / StdDev sums reduce across the shards too (montecarlo.go:18-59)
            assert fx.scaled_err(runs.Mean(k), runs1.Mean(k)) <= 1e-12
            assert fx.scaled_err(runs.StdDev(k), runs1.StdDev(k)) <= 1e-9
    # rank-ordered sum: the same shard list gives the same bits
    nis_b, nees_b = gk.NewChiSquare(kf, runs, controls, True, True)
    assert np.array_equal(nis, nis_b) and np.array_equal(nees, nees_b)


def test_multi_failed_update_is_reported_from_any_shard():
    import gokalman_b200 as gk
    gk.load()
    f = fx.jerk3()
    mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=3))
    bad, _ = gk.NewVanilla(f["x0"], np.zeros((3, 3)), f["F"], f["G"], f["H"], gk.NewNoiseless(np.zeros((3, 3)), np.zeros((1, 1))))
    runs = gk.NewMonteCarloRuns(300, 10, 1, [np.zeros(1)], mckf, devices=[0, 0], reduce="peer")
    with pytest.raises(gk.GkbError):
        gk.NewChiSquare(bad, runs, [np.zeros(1)], True, True)


@pytest.mark.parametrize("reduce", ["nccl", "peer"])
def test_multi_gpu_collective_inside_the_abi(reduce):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs in this process (gpurun --gpus 2)")
    trials, steps = 200000, 120
    gk, kf, runs1, controls = _setup(trials, steps, None, reduce)
    nis1, nees1 = gk.NewChiSquare(kf, runs1, controls, True, True)
    gk, kf, runs, controls = _setup(trials, steps, list(range(n)), reduce)
    nis, nees = gk.NewChiSquare(kf, runs, controls, True, True)
    assert fx.scaled_err(nis, nis1) <= 1e-12 and fx.scaled_err(nees, nees1) <= 1e-12
    assert fx.scaled_err(runs.Mean(steps - 1), runs1.Mean(steps - 1)) <= 1e-12
    # one host thread driving handles on two devices (the per-device timing events of the engine)
    f = fx.robot_1d()
    for d in range(n):
        h, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=64, device=d)
        est = h.UpdateBatch(np.zeros((5, 1)), fx.robot_controls(5), every_step=False)
        assert np.all(est.status == 0)
