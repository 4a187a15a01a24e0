"""AWGN (noise.go:109-159) on ordinary filters: `gkb_set_philox_noise` + the host-side AWGN.Process / Measurement.

The reference's own filter tests drive every LDKF with an AWGN noise object (vanilla_test.go:29-75,
information_test.go:45-109, squareroot_test.go:29-75).  Its samples are clock-seeded and cannot be reproduced; here
they are a function of (seed, filter, step), so the same run can be replayed through the oracle."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


@pytest.mark.parametrize("kind", ["vanilla", "information", "sqrt"])
def test_awgn_on_ordinary_filter_matches_oracle(oracle, kind):
    """vanilla_test.go:29-75 shape: the jerk-car 3-state fixture with AWGN(Q, R) ON THE FILTER, 100 Update() calls with
    the yacc measurements pattern (here seeded values) and a control.  The GPU filter (device-generated samples) must
    equal the oracle fed, through replay, with the samples AWGN.Process(k) / Measurement(k) return on the host --
    including Vanilla's second, different Process(k) draw (vanilla.go:195)."""
    gk = _gpu()
    f = fx.jerk3()
    rng = np.random.default_rng(3)
    steps, seed = 100, 4242
    y = 0.05 * rng.standard_normal((steps, 1))
    u = np.full((steps, 1), 0.1)
    noise = gk.NewAWGN(f["Q"], f["R"], seed=seed)
    ctor = {"vanilla": gk.NewVanilla, "information": gk.NewInformationFromState, "sqrt": gk.NewSquareRoot}[kind]
    kf, _ = ctor(f["x0"], f["P0"], f["F"], f["G"], f["H"], noise)
    ests = [kf.Update(y[k], u[k]) for k in range(steps)]       # one Update() per call: the drop-in shape
    # host-side samples of the same stream
    w = np.stack([noise.Process(k) for k in range(steps)])
    w2 = np.stack([(noise.Process(k), noise.Process(k))[1] for k in range(steps)])
    v = np.stack([noise.Measurement(k) for k in range(steps)])
    assert not np.allclose(w, w2) and abs(np.std(v) / np.sqrt(f["R"][0, 0]) - 1.0) < 0.25
    octor = {"vanilla": oracle.NewVanilla, "information": oracle.NewInformationFromState, "sqrt": oracle.NewSquareRoot}[kind]
    o = octor(f["x0"], f["P0"], f["F"], f["G"], f["H"], f["Q"], f["R"])
    o.SetReplayNoise(w, v, w2 if kind == "vanilla" else None)
    for k in range(steps):
        eo = o.Update(y[k], u[k])
        for name in ("State", "Measurement", "Covariance"):
            a, b = np.asarray(getattr(ests[k], name)()), np.asarray(getattr(eo, name)())
            assert fx.scaled_err(a, b) <= TOL, (kind, k, name, fx.scaled_err(a, b))
    # batched: 5 filters get 5 different noise streams (keyed by the filter index), filter 0 the one above
    kfb, _ = ctor(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=seed), n_filters=5)
    eb = kfb.UpdateBatch(y, u, every_step=True)
    xs = np.asarray(eb.State())
    assert fx.scaled_err_steps(xs[:, :, 0], np.stack([np.asarray(e.State()) for e in ests])) <= 1e-13
    # (the information filter's noise only enters y-hat, information.go:192-194: compare the measurements there)
    ys = np.asarray(eb.Measurement())
    assert not np.allclose(ys[:, :, 0], ys[:, :, 1])
    if kind != "information":
        assert not np.allclose(xs[:, :, 0], xs[:, :, 1])
    # Reset() re-arms the same seeded stream (a seeded AWGN: the reference would re-seed from the clock)
    kfb.Reset()
    eb2 = kfb.UpdateBatch(y, u, every_step=True)
    assert np.array_equal(np.asarray(eb2.State()), xs)


def test_awgn_predictor_reproduces_monte_carlo_truth():
    """A pure predictor handle with AWGN noise and n_filters = trials draws, for (filter, step), the samples the Monte
    Carlo kernel draws for (trial, step): its Estimates ARE the runs NewMonteCarloRuns describes (montecarlo.go:108-117:
    the reference builds the runs by calling this very Update in a loop)."""
    gk = _gpu()
    f = fx.robot_1d()
    steps, trials, seed = 40, 96, 99
    controls = fx.robot_controls(steps)
    x0 = np.array([0.3, -0.1])
    mckf, _ = gk.NewPurePredictorVanilla(x0, f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=seed))
    runs = gk.NewMonteCarloRuns(trials, steps, 1, list(controls), mckf)
    tx, ty = runs.Truth()
    pred, _ = gk.NewPurePredictorVanilla(x0, f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=seed), n_filters=trials)
    est = pred.UpdateBatch(np.zeros((steps, 1)), controls, every_step=True)
    assert np.array_equal(np.asarray(est.State()), tx)
    assert np.array_equal(np.asarray(est.Measurement()), ty)


def test_awgn_requires_positive_definite_noise():
    gk = _gpu()
    f = fx.jerk3()
    with pytest.raises(ValueError):  # noise.go:149-156 panics
        gk.NewAWGN(np.zeros((3, 3)), f["R"])


def test_awgn_samples_like_noise_test_go():
    """The sample half of TestAWGN (noise_test.go:136-170): Process / Measurement return vectors of the sizes of Q / R,
    different at two different steps; plus what a clock-seeded stream cannot promise: the same (seed, step) gives the same
    sample again, and the sample covariance over many steps approaches Q and R."""
    import gokalman_b200 as gk
    gk.load()
    Q, R = np.eye(2), np.array([[20.0, 0.05], [0.05, 20.0]])
    n = gk.NewAWGN(Q, R, seed=11)
    pk0, pk1, mk0, mk1 = n.Process(0), n.Process(1), n.Measurement(0), n.Measurement(1)
    assert pk0.shape == (2,) and mk0.shape == (2,)
    assert not np.array_equal(pk0, pk1) and not np.array_equal(mk0, mk1)
    n2 = gk.NewAWGN(Q, R, seed=11)
    assert np.array_equal(n2.Process(0), pk0) and np.array_equal(n2.Measurement(1), mk1)
    W = np.array([gk.NewAWGN(Q, R, seed=11).Process(k) for k in range(4000)])
    V = np.array([n.Measurement(k) for k in range(4000)])
    assert np.allclose(np.cov(W.T), Q, atol=0.08) and np.allclose(np.cov(V.T), R, atol=1.6)
    assert abs(W.mean()) < 0.05 and abs(V.mean()) < 0.25
