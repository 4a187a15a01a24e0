"""GPU parity, Monte Carlo + chi-square (montecarlo.go + chisquare.go fused kernel) against the oracle.

Parity protocol: the reference's AWGN stream (Go ziggurat on a clock seed) cannot
be reproduced, so the kernel DUMPS the coloured noise it generated from Philox and the oracle
replays exactly those samples through its restatement of NewMonteCarloRuns + NewChiSquare.
Tolerance 1e-10 relative to the array's max-abs."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


def _mc_pair(gk, oracle, fxm, kind, trials, steps, controls, seed=0x5EED, trial_offset=0):
    noise = gk.NewAWGN(fxm["Q"], fxm["R"], seed=seed)
    mckf, _ = gk.NewPurePredictorVanilla(fxm["x0_truth"], fxm["P0"], fxm["F"], fxm["G"], fxm["H"], noise)
    tested_noise = gk.NewNoiseless(fxm["Q"], fxm["R"])
    ctor = {"vanilla": gk.NewVanilla, "information": gk.NewInformationFromState, "sqrt": gk.NewSquareRoot}[kind]
    chikf, _ = ctor(fxm["x0"], fxm["P0"], fxm["F"], fxm["G"], fxm["H"], tested_noise)
    runs = gk.NewMonteCarloRuns(trials, steps, fxm["H"].shape[0], controls, mckf, trial_offset=trial_offset)
    with_nis = kind != "information"
    nis, nees = gk.NewChiSquare(chikf, runs, controls, True, with_nis)
    tx, ty, w, v = runs.Truth(with_noise=True)
    okind = {"vanilla": oracle.VANILLA, "information": oracle.INFORMATION, "sqrt": oracle.SQRT}[kind]
    ctrl = None
    if controls is not None and len(controls) == steps:
        ctrl = np.stack(controls)
    ref = oracle.mc_chisquare(okind, fxm["F"], fxm["G"], fxm["H"], fxm["Q"], fxm["R"], fxm["x0_truth"], fxm["x0"],
                              fxm["P0"], trials, steps, controls=ctrl,
                              w=np.ascontiguousarray(w.transpose(2, 0, 1)), v=np.ascontiguousarray(v.transpose(2, 0, 1)),
                              with_nees=True, with_nis=with_nis, want_truth=True, want_stats=True)
    return dict(nis=nis, nees=nees, tx=tx, ty=ty, w=w, v=v, ref=ref, runs=runs)


def _jerk3():
    f = fx.jerk3()
    f["x0_truth"] = f["x0"]
    return f


def _robot():
    f = fx.robot_1d()
    f["x0_truth"] = np.array([0.7, -0.4])  # "a random initial state" drawn once (examples/robot/main.go:28-30)
    return f


@pytest.mark.parametrize("kind", ["vanilla", "information", "sqrt"])
def test_mc_chisquare_robot_matches_oracle(oracle, kind):
    """examples/robot/main.go (n=2, m=1, cos controls), all three tested filter kinds (config 3)."""
    gk = _gpu()
    steps, trials = 120, 300
    controls = list(fx.robot_controls(steps))
    r = _mc_pair(gk, oracle, _robot(), kind, trials, steps, controls)
    ref = r["ref"]
    assert fx.scaled_err(r["tx"], ref["truth_x"].transpose(1, 2, 0)) <= TOL
    assert fx.scaled_err(r["ty"], ref["truth_y"].transpose(1, 2, 0)) <= TOL
    assert fx.scaled_err(r["nees"], ref["NEES"]) <= TOL, fx.scaled_err(r["nees"], ref["NEES"])
    if kind != "information":
        assert fx.scaled_err(r["nis"], ref["NIS"]) <= TOL


def test_mc_chisquare_jerk3_matches_oracle(oracle):
    """BASELINE config 2 shape (3-state jerk fixture of montecarlo_test.go:12-20, m=1, zero controls
    from a single control vector), reduced trial count."""
    gk = _gpu()
    steps, trials = 300, 257  # > one 256-step chunk, ragged trial count
    r = _mc_pair(gk, oracle, _jerk3(), "vanilla", trials, steps, [np.zeros(1)])
    ref = r["ref"]
    assert fx.scaled_err(r["nees"], ref["NEES"]) <= TOL
    assert fx.scaled_err(r["nis"], ref["NIS"]) <= TOL
    # MonteCarloRuns.Mean / StdDev (montecarlo.go:18-59)
    for k in (0, 17, steps - 1):
        assert fx.scaled_err(r["runs"].Mean(k), ref["mean"][k]) <= 1e-9
        assert fx.scaled_err(r["runs"].StdDev(k), ref["std"][k]) <= 1e-7  # one-pass vs two-pass variance


@pytest.mark.parametrize("kind", ["vanilla", "sqrt"])
def test_mc_chisquare_multid4_two_philox_blocks(oracle, kind):
    """vanilla_test.go:97-115's 4-state / 2-measurement system: n + m = 6 normals per step, i.e. two Philox blocks
    per (trial, step) and the general (non-hoisted) generator; m = 2 exercises the 2 x 2 NIS inverse."""
    gk = _gpu()
    f = fx.multid4()
    # a positive definite Q for the AWGN colouring (the fixture's 3 x 3 block is numerically singular)
    f["Q"] = f["Q"] + 1e-9 * np.eye(4)
    f["x0_truth"] = f["x0"]
    steps, trials = 70, 200
    controls = [np.array([0.1 * np.sin(0.05 * k)]) for k in range(steps)]
    r = _mc_pair(gk, oracle, f, kind, trials, steps, controls)
    ref = r["ref"]
    assert fx.scaled_err(r["nees"], ref["NEES"]) <= TOL
    assert fx.scaled_err(r["nis"], ref["NIS"]) <= TOL
    # and the noise the kernel drew is the oracle's own stream for the same (seed, trial, step)
    LQ, _ = oracle.chol_lower(f["Q"])
    LR, _ = oracle.chol_lower(f["R"])
    for t in (0, 199):
        z = oracle.philox_normals(0x5EED, t, 5, 6)
        assert np.max(np.abs(r["w"][5, :, t] - LQ @ z[:4])) <= 1e-14 * np.max(np.abs(LQ))
        assert np.max(np.abs(r["v"][5, :, t] - LR @ z[4:])) <= 1e-14


def test_philox_stream_matches_oracle(oracle):
    """The device Philox4x32-10 + inverse-CDF stream is the one the oracle restates: same integers, the same
    table and fused Horner form, so the normals are bit-identical (the colouring L z may differ by an FMA
    rounding: 1e-15), keyed by the GLOBAL trial index."""
    gk = _gpu()
    f = _jerk3()
    steps, trials, seed, off = 5, 64, 0xC0FFEE, 12345
    noise = gk.NewAWGN(f["Q"], f["R"], seed=seed)
    mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], noise)
    runs = gk.NewMonteCarloRuns(trials, steps, 1, [np.zeros(1)], mckf, trial_offset=off)
    _, _, w, v = runs.Truth(with_noise=True)
    LQ, _ = oracle.chol_lower(f["Q"])
    LR, _ = oracle.chol_lower(f["R"])
    for t in (0, 31, 63):
        for k in range(steps):
            z = oracle.philox_normals(seed, off + t, k, 4)
            assert np.max(np.abs(w[k, :, t] - LQ @ z[:3])) <= 4e-15 * np.max(np.abs(LQ))
            assert abs(v[k, 0, t] - (LR @ z[3:])[0]) <= 4e-15
            assert v[k, 0, t] == LR[0, 0] * z[3]  # one multiplication: the normal itself is bit-identical


def test_mc_sharding_invariance():
    """8(e): trials shard across GPUs by contiguous ranges with Philox keyed by global trial id, so
    the per-step SUMS of two half-shards add up to the full run's (to summation rounding)."""
    gk = _gpu()
    f = _jerk3()
    steps, trials, seed = 64, 1000, 99

    def sums(n_trials, offset):
        noise = gk.NewAWGN(f["Q"], f["R"], seed=seed)
        mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], noise)
        chikf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
        runs = gk.NewMonteCarloRuns(n_trials, steps, 1, [np.zeros(1)], mckf, trial_offset=offset)
        nis, nees = gk.NewChiSquare(chikf, runs, [np.zeros(1)], True, True)
        return nis * n_trials, nees * n_trials
    full = sums(trials, 0)
    a, b = sums(400, 0), sums(600, 400)
    assert fx.scaled_err(a[0] + b[0], full[0]) <= 1e-12
    assert fx.scaled_err(a[1] + b[1], full[1]) <= 1e-12


def test_chisquare_philox_end_to_end_and_statistics(oracle):
    """examples/robot/main.go:32-58 at a larger size.  (a) The whole PHILOX pipeline (device RNG,
    colouring, truth, filter, NEES/NIS means) equals the oracle running its OWN restatement of the
    same Philox stream (no noise hand-over): the normals are bit-identical (same inverse-CDF table on both
    sides), so this holds to the filter parity bar, 1e-10.  (b) Statistical sanity at 2e5 trials: NIS -> m.  (NEES does
    not tend to n here: the reference pairs the state x_{k+1} with the measurement of x_k,
    montecarlo.go:110-113 / vanilla.go:155-157 -- both sides reproduce that.)"""
    gk = _gpu()
    f = _robot()
    f["x0_truth"] = np.zeros(2)
    steps = 120
    controls = list(fx.robot_controls(steps))

    def gpu_run(trials, seed):
        noise = gk.NewAWGN(f["Q"], f["R"], seed=seed)
        mckf, _ = gk.NewPurePredictorVanilla(f["x0_truth"], f["P0"], f["F"], f["G"], f["H"], noise)
        chikf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
        runs = gk.NewMonteCarloRuns(trials, steps, 1, controls, mckf)
        return gk.NewChiSquare(chikf, runs, controls, True, True)
    nis, nees = gpu_run(4000, 2024)
    ref = oracle.mc_chisquare(oracle.VANILLA, f["F"], f["G"], f["H"], f["Q"], f["R"], f["x0_truth"], f["x0"], f["P0"],
                              4000, steps, controls=np.stack(controls), seed=2024, threads=4)
    assert fx.scaled_err(nis, ref["NIS"]) <= 1e-10
    assert fx.scaled_err(nees, ref["NEES"]) <= 1e-10
    nis, nees = gpu_run(200000, 7)
    assert nis.shape == nees.shape == (steps,)
    assert np.all(np.isfinite(nees)) and np.all(nees > 0)
    assert abs(np.mean(nis[40:]) - 1.0) < 0.02


def test_chisquare_errors_mirror_reference():
    """montecarlo_test.go:54-88"""
    gk = _gpu()
    f = _jerk3()
    noise = gk.NewAWGN(f["Q"], f["R"], seed=1)
    mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], noise)
    with pytest.raises(ValueError):  # two control vectors for ten steps
        gk.NewMonteCarloRuns(5, 10, 1, [np.zeros(1), np.zeros(1)], mckf)
    notpred, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
    with pytest.raises(ValueError):  # not a pure predictor
        gk.NewMonteCarloRuns(5, 10, 1, [np.zeros(1)], notpred)
    runs = gk.NewMonteCarloRuns(5, 10, 1, [np.zeros(1)], mckf)
    with pytest.raises(gk.GkbError):  # neither NEES nor NIS
        gk.NewChiSquare(notpred, runs, [np.zeros(1)], False, False)
    with pytest.raises(ValueError):  # too few controls
        gk.NewChiSquare(notpred, runs, [np.zeros(1), np.zeros(1)], True, False)
    nis, nees = gk.NewChiSquare(notpred, runs, [np.zeros(1)], True, True)
    assert len(nis) == len(nees) == 10


@pytest.mark.parametrize("kind", ["vanilla", "information", "sqrt"])
def test_chisquare_tested_filter_carries_its_own_model(oracle, kind):
    """chisquare.go:16-42,64-66: NewChiSquare runs ANY LDKF -- with its own F / G / H and its own Noise -- on the
    stored truth.  Here the truth is the robot model and the tested filter is deliberately mis-tuned: Q ten times
    too small, a slightly wrong F, another G, a scaled H row and another R.  The GPU must (a) equal the oracle
    running the same two-model experiment to 1e-10 and (b) differ visibly from the matched-filter statistics
    (that difference is what NEES / NIS exist to show)."""
    gk = _gpu()
    f = _robot()
    steps, trials = 100, 400
    controls = list(fx.robot_controls(steps))
    tF = f["F"] + np.array([[0.0, 0.01], [0.0, -0.02]])
    tG = 0.9 * f["G"]
    tH = np.array([[1.05, 0.0]])
    tQ, tR = 0.1 * f["Q"], np.array([[0.08]])
    mckf, _ = gk.NewPurePredictorVanilla(f["x0_truth"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=77))
    ctor = {"vanilla": gk.NewVanilla, "information": gk.NewInformationFromState, "sqrt": gk.NewSquareRoot}[kind]
    tested, _ = ctor(f["x0"], f["P0"], tF, tG, tH, gk.NewNoiseless(tQ, tR))
    matched, _ = ctor(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
    runs = gk.NewMonteCarloRuns(trials, steps, 1, controls, mckf)
    with_nis = kind != "information"
    nis, nees = gk.NewChiSquare(tested, runs, controls, True, with_nis)
    nis_m, nees_m = gk.NewChiSquare(matched, runs, controls, True, with_nis)
    tx, ty, w, v = runs.Truth(with_noise=True)
    okind = {"vanilla": oracle.VANILLA, "information": oracle.INFORMATION, "sqrt": oracle.SQRT}[kind]
    ref = oracle.mc_chisquare(okind, f["F"], f["G"], f["H"], f["Q"], f["R"], f["x0_truth"], f["x0"], f["P0"], trials, steps,
                              controls=np.stack(controls), w=np.ascontiguousarray(w.transpose(2, 0, 1)),
                              v=np.ascontiguousarray(v.transpose(2, 0, 1)), with_nees=True, with_nis=with_nis,
                              tested=dict(F=tF, G=tG, H=tH, Q=tQ, R=tR), want_truth=True)
    # the truth is the PREDICTOR's trajectory (its own Q / R colour the noise), whatever the tested filter carries
    assert fx.scaled_err(tx, ref["truth_x"].transpose(1, 2, 0)) <= TOL
    assert fx.scaled_err(nees, ref["NEES"]) <= TOL, fx.scaled_err(nees, ref["NEES"])
    if with_nis:
        assert fx.scaled_err(nis, ref["NIS"]) <= TOL, fx.scaled_err(nis, ref["NIS"])
    # the mis-tuned filter is visibly inconsistent: its NEES is far from the matched filter's (and from n = 2)
    assert np.mean(nees[20:]) > 1.5 * np.mean(nees_m[20:]), (np.mean(nees[20:]), np.mean(nees_m[20:]))
    if with_nis:
        assert abs(np.mean(nis[20:]) - np.mean(nis_m[20:])) > 0.1


def test_chisquare_information_nis_when_n_equals_m(oracle):
    """chisquare.go:61-77 with an information filter: Innovation() is the n-vector i+ (information.go:272-274), so
    the NIS product only has matching dimensions when n == m; then it is i+^T inv(H inv(I-) H^T + R) i+."""
    gk = _gpu()
    rng = np.random.default_rng(12)
    n = m = 2
    F = np.array([[1.0, 0.1], [-0.05, 0.98]])
    H = np.array([[1.0, 0.2], [0.0, 1.0]])
    Q, R = np.array([[2e-2, 1e-3], [1e-3, 1e-2]]), np.array([[0.05, 0.01], [0.01, 0.08]])
    x0, P0 = np.array([0.3, -0.2]), np.diag([2.0, 1.0])
    steps, trials = 60, 200
    mckf, _ = gk.NewPurePredictorVanilla(x0, P0, F, None, H, gk.NewAWGN(Q, R, seed=5))
    kf, _ = gk.NewInformationFromState(x0, P0, F, None, H, gk.NewNoiseless(Q, R))
    runs = gk.NewMonteCarloRuns(trials, steps, m, None, mckf)
    nis, nees = gk.NewChiSquare(kf, runs, None, True, True)
    _, _, w, v = runs.Truth(with_noise=True)
    ref = oracle.mc_chisquare(oracle.INFORMATION, F, None, H, Q, R, x0, x0, P0, trials, steps,
                              w=np.ascontiguousarray(w.transpose(2, 0, 1)), v=np.ascontiguousarray(v.transpose(2, 0, 1)))
    assert np.all(nis > 0)
    assert fx.scaled_err(nees, ref["NEES"]) <= TOL, fx.scaled_err(nees, ref["NEES"])
    assert fx.scaled_err(nis, ref["NIS"]) <= TOL, fx.scaled_err(nis, ref["NIS"])


def test_chisquare_failed_update_is_reported():
    """chisquare.go:40-42: the reference panics when the tested filter's Update returns an error.  A tested filter
    with P0 = 0, Q = 0, R = 0 has S = H P- H^T + R = 0 at the first step (vanilla.go:164-167): the call must raise,
    not return silently biased means."""
    gk = _gpu()
    f = _jerk3()
    mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=3))
    bad, _ = gk.NewVanilla(f["x0"], np.zeros((3, 3)), f["F"], f["G"], f["H"], gk.NewNoiseless(np.zeros((3, 3)), np.zeros((1, 1))))
    runs = gk.NewMonteCarloRuns(64, 10, 1, [np.zeros(1)], mckf)
    with pytest.raises(gk.GkbError) as ei:
        gk.NewChiSquare(bad, runs, [np.zeros(1)], True, True)
    assert ei.value.code == -2


@pytest.mark.parametrize("kind", ["vanilla", "information", "sqrt"])
@pytest.mark.parametrize("n,m", [(8, 3), (7, 2)])
def test_mc_chisquare_n7_n8_matches_oracle(oracle, kind, n, m):
    """The north star's n <= 8: the fused Monte Carlo + chi-square kernels at n = 7 and 8 (kernels_mc.cu parts 4-6),
    a seeded random model with a control, all three tested filter kinds, against the oracle on the dumped noise."""
    gk = _gpu()
    rng = np.random.default_rng(100 * n + m)
    A = rng.standard_normal((n, n))
    B = rng.standard_normal((m, m))
    f = dict(F=np.eye(n) + 0.05 * rng.standard_normal((n, n)), G=rng.standard_normal((n, 1)), H=rng.standard_normal((m, n)),
             Q=1e-3 * (A @ A.T + n * np.eye(n)), R=1e-1 * (B @ B.T + m * np.eye(m)), x0=rng.standard_normal(n),
             P0=np.diag(rng.uniform(0.5, 2.0, n)))
    f["x0_truth"] = f["x0"] + 0.1 * rng.standard_normal(n)
    steps, trials = 40, 70
    controls = [np.array([0.2 * np.cos(0.1 * k)]) for k in range(steps)]
    r = _mc_pair(gk, oracle, f, kind, trials, steps, controls)
    ref = r["ref"]
    assert fx.scaled_err(r["tx"], ref["truth_x"].transpose(1, 2, 0)) <= TOL
    assert fx.scaled_err(r["nees"], ref["NEES"]) <= TOL, fx.scaled_err(r["nees"], ref["NEES"])
    if kind != "information":
        assert fx.scaled_err(r["nis"], ref["NIS"]) <= TOL, fx.scaled_err(r["nis"], ref["NIS"])


def test_statod5044_example_monte_carlo_and_csv(oracle):
    """examples/statOD5044/main.go:36-116: the 4-state / 2-control / 2-measurement statOD model -- truth from a pure
    predictor with AWGN on the closed-loop F - G T (zero control matrix there: here G with zero controls, the
    `len(controls) == 1` rule of montecarlo.go:98-104), NEES / NIS of a Vanilla filter on the same model, and
    MonteCarloRuns.AsCSV (montecarlo.go:62-89) against Mean / StdDev / the dumped truth."""
    gk = _gpu()
    f = fx.statod4()
    steps, trials = 108, 15  # numMC = 15 (main.go:74); samples shortened from 1086
    fm = dict(F=f["Fcl"], G=f["G"], H=f["H"], Q=f["Q"], R=f["R"], x0=f["x0"], P0=f["P0"], x0_truth=f["x0"])
    r = _mc_pair(gk, oracle, fm, "vanilla", trials, steps, [np.zeros(2)])
    ref = r["ref"]
    assert fx.scaled_err(r["tx"], ref["truth_x"].transpose(1, 2, 0)) <= TOL
    assert fx.scaled_err(r["nees"], ref["NEES"]) <= TOL and fx.scaled_err(r["nis"], ref["NIS"]) <= TOL
    headers = ["dr", "dr_dot", "dtheta", "dtheta_dot"]  # main.go:77
    csvs = r["runs"].AsCSV(headers)
    assert len(csvs) == 4
    for i, text in enumerate(csvs):
        lines = text.split("\n")
        assert lines[0] == "".join("%s-%d," % (headers[i], k) for k in range(trials)) + headers[i] + "-mean," + headers[i] + "-stddev"
        assert len(lines) == steps + 1
        for k in (0, steps // 2, steps - 1):
            vals = [float(v) for v in lines[1 + k].split(",")]
            assert len(vals) == trials + 2
            assert np.allclose(vals[:trials], r["tx"][k, i, :], atol=5e-7)            # %f: six decimals
            assert abs(vals[trials] - ref["mean"][k, i]) <= 5e-7 + 1e-9 * abs(ref["mean"][k, i])
            assert abs(vals[trials + 1] - ref["std"][k, i]) <= 5e-7 + 1e-6 * abs(ref["std"][k, i])
