"""GPU parity, LDKF kinds: the CUDA engine (through the C-ABI) against the CPU oracle on the same
inputs.  Tolerance: 1e-10 relative to the array's max-abs (BASELINE.json north_star; metric of
SURVEY.md 8(c)), FP64 throughout."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


def _jerkcar_gpu_filters(gk):
    def make(f):
        n2 = gk.NewNoiseless(f["Q"], f["Ra"])
        v, v0 = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H2"], n2)
        i, _ = gk.NewInformation(np.zeros(4), np.zeros((4, 4)), f["F"], f["G"], f["H2"], gk.NewNoiseless(f["Q"], f["Ra"]))
        s, s0 = gk.NewSquareRoot(f["x0"], f["P0"], f["F"], f["G"], f["H2"], gk.NewNoiseless(f["Q"], f["Ra"]))

        class Adapt:  # run_jerkcar calls SetNoise(Q, R) with raw matrices
            def __init__(self, kf):
                self.kf = kf

            def SetMeasurementMatrix(self, H):
                self.kf.SetMeasurementMatrix(H)

            def SetNoise(self, Q, R):
                self.kf.SetNoise(gk.NewNoiseless(Q, R))

            def Update(self, y, u):
                return self.kf.Update(y, u)

        class Zero:  # est0 of the information filter: zero state / covariance row (information.csv:3)
            def State(self):
                return np.zeros(4)

            def Covariance(self):
                return np.zeros((4, 4))
        return [("vanilla", Adapt(v), v0), ("information", Adapt(i), Zero()), ("sqrt", Adapt(s), s0)]
    return make


def _jerkcar_oracle_filters(gko):
    def make(f):
        v = gko.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        i = gko.NewInformation(np.zeros(4), np.zeros((4, 4)), f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        s = gko.NewSquareRoot(f["x0"], f["P0"], f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        return [("vanilla", v, v.InitialEstimate()), ("information", i, i.InitialEstimate()),
                ("sqrt", s, s.InitialEstimate())]
    return make


def test_jerkcar_example_drop_in(oracle):
    """BASELINE config 1: examples/jerkcar/main.go replayed through Update()/SetMeasurementMatrix()/
    SetNoise() one step at a time (n = 4, m switching 1 <-> 2 every 10th step, 2000 steps), for the
    vanilla, information and square-root filters.  Checked against the reference's golden CSVs
    (print precision) and against the oracle (1e-10)."""
    gk = _gpu()
    got = fx.run_jerkcar(_jerkcar_gpu_filters(gk))
    ref = fx.run_jerkcar(_jerkcar_oracle_filters(oracle))
    gold = fx.load_jerkcar_golden()
    for name in ("vanilla", "information", "sqrt"):
        assert np.abs(got[name] - gold[name]).max() <= 5.0e-7 + 1e-9, name
        # per column (a state component and its 2-sigma bounds), scaled by the column's max-abs
        for col in range(12):
            err = fx.scaled_err(got[name][:, col], ref[name][:, col])
            assert err <= TOL, (name, col, err)


def _rand_spd(rng, n, scale=1.0):
    A = rng.standard_normal((n, n))
    return scale * (A @ A.T + n * np.eye(n))


def _model(rng, n, m, c):
    F = np.eye(n) + 0.05 * rng.standard_normal((n, n))
    G = rng.standard_normal((n, c)) if c else None
    H = rng.standard_normal((m, n))
    Q = _rand_spd(rng, n, 1e-3)
    R = _rand_spd(rng, m, 1e-1)
    P0 = _rand_spd(rng, n, 1.0)
    x0 = rng.standard_normal(n)
    return F, G, H, Q, R, x0, P0


def _compare(est_gpu, ests_ref, fields, tag):
    """est_gpu: batched Estimate [steps, ..., N]; ests_ref[f][k]: oracle estimates."""
    getters = {"state": "State", "meas": "Measurement", "innov": "Innovation", "covar": "Covariance",
               "pred_covar": "PredCovariance", "gain": "Gain", "obs_dev": "ObservationDev"}
    for fld in fields:
        g = getattr(est_gpu, getters[fld])()
        nf = len(ests_ref)
        steps = len(ests_ref[0])
        for f in range(nf):
            ref = np.stack([np.asarray(getattr(ests_ref[f][k], getters[fld])()) for k in range(steps)])
            got = g[..., f] if nf > 1 else g
            got = got.reshape(ref.shape)
            err = fx.scaled_err_steps(got, ref)  # every step scaled by its own max-abs
            assert err <= TOL, (tag, fld, f, err)


SHAPES = [(2, 1, 1), (3, 1, 1), (4, 2, 1), (6, 2, 0), (5, 2, 2), (3, 3, 1), (1, 1, 0), (8, 3, 2), (7, 2, 1), (8, 1, 0)]


@pytest.mark.parametrize("n,m,c", SHAPES)
@pytest.mark.parametrize("kind", ["vanilla", "predictor", "information", "sqrt"])
def test_batched_update_matches_oracle(oracle, kind, n, m, c):
    """UpdateBatch (many steps in one launch, per-filter measurements, replayed noise) against the
    oracle run filter by filter, step by step: every Estimate field, every step, every filter."""
    gk = _gpu()
    rng = np.random.default_rng(1000 * n + 10 * m + c)
    F, G, H, Q, R, x0, P0 = _model(rng, n, m, c)
    nf, steps = 37, 60
    y = rng.standard_normal((steps, m, nf))
    u = rng.standard_normal((steps, c)) if c else None
    w = 0.03 * rng.standard_normal((steps, n, nf))
    v = 0.3 * rng.standard_normal((steps, m, nf))
    noise = gk.ReplayNoise(Q, R, w, v)
    if kind == "vanilla":
        kf, _ = gk.NewVanilla(x0, P0, F, G, H, noise, n_filters=nf)
    elif kind == "predictor":
        kf, _ = gk.NewPurePredictorVanilla(x0, P0, F, G, H, noise, n_filters=nf)
    elif kind == "information":
        kf, _ = gk.NewInformationFromState(x0, P0, F, G, H, noise, n_filters=nf)
    else:
        kf, _ = gk.NewSquareRoot(x0, P0, F, G, H, noise, n_filters=nf)
    est = kf.UpdateBatch(y, u, every_step=True)
    assert np.all(est.status == 0)
    refs = []
    for f in range(nf):
        if kind in ("vanilla", "predictor"):
            o = oracle.NewVanilla(x0, P0, F, G, H, Q, R, predictor=(kind == "predictor"))
        elif kind == "information":
            o = oracle.NewInformationFromState(x0, P0, F, G, H, Q, R)
        else:
            o = oracle.NewSquareRoot(x0, P0, F, G, H, Q, R)
        o.SetReplayNoise(w[:, :, f], v[:, :, f])
        refs.append([o.Update(y[k, :, f], None if u is None else u[k]) for k in range(steps)])
    fields = ["state", "meas", "innov", "covar", "pred_covar"] + ([] if kind == "information" else ["gain"])
    _compare(est, refs, fields, kind)
    # the raw device state after the batch equals the oracle's internal representation
    vec, mat = kf.GetState()
    for f in range(0, nf, 9):
        rv, rm, _ = refs[f][-1].raw()
        if kind in ("vanilla", "predictor"):
            rv, rm = refs[f][-1].State(), refs[f][-1].Covariance()
        elif kind == "sqrt":
            rv = refs[f][-1].State()
        assert fx.scaled_err(vec[:, f], rv) <= TOL
        assert fx.scaled_err(mat[:, :, f], rm) <= TOL


def test_update_then_update_equals_batch(oracle):
    """Step counter and state persist across calls: 3 launches of 20 steps == 1 launch of 60."""
    gk = _gpu()
    rng = np.random.default_rng(7)
    F, G, H, Q, R, x0, P0 = _model(rng, 4, 2, 1)
    nf, steps = 5, 60
    y, u = rng.standard_normal((steps, 2, nf)), rng.standard_normal((steps, 1))
    w, v = 0.03 * rng.standard_normal((steps, 4, nf)), 0.3 * rng.standard_normal((steps, 2, nf))
    a, _ = gk.NewVanilla(x0, P0, F, G, H, gk.ReplayNoise(Q, R, w, v), n_filters=nf)
    b, _ = gk.NewVanilla(x0, P0, F, G, H, gk.ReplayNoise(Q, R, w, v), n_filters=nf)
    ea = a.UpdateBatch(y, u, every_step=True)
    parts = [b.UpdateBatch(y[i:i + 20], u[i:i + 20], every_step=True) for i in (0, 20, 40)]
    assert np.array_equal(ea.State(), np.concatenate([p.State() for p in parts]))
    assert np.array_equal(ea.Covariance(), np.concatenate([p.Covariance() for p in parts]))
    b.Reset()
    eb = b.UpdateBatch(y, u, every_step=True)
    assert np.array_equal(ea.State(), eb.State())


def test_error_paths_mirror_reference():
    """vanilla_test.go:9-27,86-92 / information_test.go / squareroot_test.go: dimension errors."""
    gk = _gpu()
    f = fx.jerk3()
    noise = gk.NewNoiseless(f["Q"], f["R"])
    with pytest.raises(gk.GkbError):
        gk.NewVanilla(np.zeros(2), f["P0"], f["F"], f["G"], f["H"], noise)  # x0 vs Covar0
    with pytest.raises(gk.GkbError):
        gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], np.zeros((1, 2)), noise)  # H vs x0
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], noise)
    with pytest.raises(gk.GkbError):
        kf.Update(np.zeros(1), np.zeros(2))  # control size (vanilla_test.go:86-88)
    with pytest.raises(gk.GkbError):
        kf.Update(np.zeros(2), np.zeros(1))  # measurement size (vanilla_test.go:89-91)
    # a singular innovation covariance is reported, not silently used (vanilla.go:164-167)
    kz, _ = gk.NewVanilla(np.zeros(2), np.zeros((2, 2)), np.eye(2), None, np.array([[1.0, 0]]),
                          gk.NewNoiseless(np.zeros((2, 2)), np.zeros((1, 1))))
    with pytest.raises(gk.GkbError) as ei:
        kz.Update(np.zeros(1), None)
    assert ei.value.code == -2
