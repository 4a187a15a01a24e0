"""GPU parity, large-state Vanilla (BASELINE configs[4]: synthetic 32-state filters; warp-per-filter
FP64 tensor-core kernel, kernels_tile.cu) against the CPU oracle's Vanilla.Update (vanilla.go:128-220)
on the same inputs.  Tolerance 1e-10 of each array's max-abs, every Estimate field of every step."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10
FIELDS = ("State", "Measurement", "Innovation", "Covariance", "PredCovariance", "Gain")


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


def _oracle_run(oracle, f, ys, x0=None):
    kf = oracle.NewVanilla(f["x0"] if x0 is None else x0, f["P0"], f["F"], None, f["H"], f["Q"], f["R"])
    return [kf.Update(y, None) for y in ys]


@pytest.mark.parametrize("n,m,nf", [(32, 8, 19), (32, 3, 9), (16, 8, 40), (16, 1, 5), (24, 5, 11), (64, 8, 7), (64, 2, 4), (48, 5, 5), (40, 8, 6), (56, 3, 5)])
def test_tile_vanilla_every_step_matches_oracle(oracle, n, m, nf):
    gk = _gpu()
    f = fx.synth_lti(n, m, seed=100 + n + m)
    rng = np.random.default_rng(n * 100 + m)
    steps = 25
    y = rng.standard_normal((steps, m, nf))
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], None, f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=nf)
    assert kf._fm  # the large-state (filter-major, warp-per-filter) path
    est = kf.UpdateBatch(y, None, every_step=True)
    assert np.all(est.status == 0)
    for fi in sorted(set([0, 1, nf // 2, nf - 1])):
        refs = _oracle_run(oracle, f, y[:, :, fi])
        for name in FIELDS:
            got = np.asarray(getattr(est, name)())
            for k in (0, 1, steps // 2, steps - 1):
                ref = np.asarray(getattr(refs[k], name)())
                g = got[k][..., fi]
                err = fx.scaled_err(g, ref.reshape(g.shape))
                assert err <= TOL, (name, fi, k, err)


def test_tile_vanilla_state_carries_over_calls_and_reset(oracle):
    """Two calls of 10 steps == one call of 20 (state written back between calls), per-filter x0,
    final-only outputs, GetState, Reset."""
    gk = _gpu()
    n, m, nf, steps = 32, 8, 300, 20  # 300 filters: more than one round of the persistent grid's first CTA
    f = fx.synth_lti(n, m, seed=7)
    rng = np.random.default_rng(77)
    y = rng.standard_normal((steps, m, nf))
    x0 = rng.standard_normal((n, nf))
    kf, _ = gk.NewVanilla(x0, f["P0"], f["F"], None, f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=nf)
    e1 = kf.UpdateBatch(y[:10], None, every_step=False, want=("state",))
    e2 = kf.UpdateBatch(y[10:], None, every_step=False, want=("state", "covar"))
    vec, mat = kf.GetState()
    kf.Reset()
    e3 = kf.UpdateBatch(y, None, every_step=False, want=("state", "covar"))
    assert np.array_equal(np.asarray(e2.State()), np.asarray(e3.State()))
    assert np.array_equal(np.asarray(e2.Covariance()), np.asarray(e3.Covariance()))
    assert np.array_equal(vec, np.asarray(e3.State())) and np.array_equal(mat, np.asarray(e3.Covariance()))
    for fi in (0, 150, 299):
        refs = _oracle_run(oracle, f, y[:, :, fi], x0=x0[:, fi])
        assert fx.scaled_err(np.asarray(e1.State())[:, fi], refs[9].State()) <= TOL
        assert fx.scaled_err(np.asarray(e3.State())[:, fi], refs[-1].State()) <= TOL
        assert fx.scaled_err(np.asarray(e3.Covariance())[:, :, fi], refs[-1].Covariance()) <= TOL


def test_tile_vanilla_shared_measurement_and_singular_s(oracle):
    gk = _gpu()
    n, m, nf = 16, 8, 6
    f = fx.synth_lti(n, m, seed=3)
    y = np.random.default_rng(1).standard_normal((12, m))
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], None, f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=nf)
    est = kf.UpdateBatch(y, None, every_step=False, want=("state", "covar"))
    refs = _oracle_run(oracle, f, y)
    for fi in range(nf):
        assert fx.scaled_err(np.asarray(est.State())[:, fi], refs[-1].State()) <= TOL
    # H = 0, R = 0 -> S = 0: the reference fails with "could not invert" (vanilla.go:164-167)
    kf2, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], None, np.zeros((m, n)), gk.NewNoiseless(f["Q"], np.zeros((m, m))),
                           n_filters=nf)
    est2 = kf2.UpdateBatch(y, None, every_step=False, want=("state",))
    assert np.all(est2.status == -2)


@pytest.mark.parametrize("n,m", [(16, 5), (32, 8), (64, 3)])
def test_tile_vanilla_ill_conditioned_s_is_the_reference_error(oracle, n, m):
    """mat64.Dense.Inverse also fails on cond(S) > 1e16 (vanilla.go:164-167 returns the error): an S that is positive
    definite with exact, positive Gauss-Jordan pivots but badly scaled (R = diag(1e12, 1e-8, 1, ...) against P ~ 1e-8) is
    refused by the oracle's condition test and must be refused by the large-state kernel too -- while the same model with a
    benign R runs."""
    gk = _gpu()
    nf = 5
    f = fx.synth_lti(n, m, seed=3)
    P0, Q = 1e-8 * np.eye(n), 1e-8 * f["Q"]
    R_bad = np.diag(([1e12, 1e-8] + [1.0] * 8)[:m])
    y = np.zeros((3, m))
    with pytest.raises(oracle.OracleError):
        oracle.NewVanilla(f["x0"], P0, f["F"], None, f["H"], Q, R_bad).Update(y[0], None)
    kf, _ = gk.NewVanilla(f["x0"], P0, f["F"], None, f["H"], gk.NewNoiseless(Q, R_bad), n_filters=nf)
    assert kf._fm
    est = kf.UpdateBatch(y, None, every_step=False, want=("state",))
    assert np.all(est.status == -2)
    R_ok = np.diag(([1e3, 1e-3] + [1.0] * 8)[:m])
    kf2, _ = gk.NewVanilla(f["x0"], P0, f["F"], None, f["H"], gk.NewNoiseless(Q, R_ok), n_filters=nf)
    est2 = kf2.UpdateBatch(y, None, every_step=False, want=("state", "covar"))
    assert np.all(est2.status == 0)
    ref = oracle.NewVanilla(f["x0"], P0, f["F"], None, f["H"], Q, R_ok)
    for k in range(3):
        r = ref.Update(y[k], None)
    assert fx.scaled_err(np.asarray(est2.Covariance())[:, :, 0], r.Covariance()) <= TOL


def test_tile_vanilla_with_input_control(oracle):
    """x- = F x + G u (vanilla.go:138-143) on a large-state handle: the control term G u is formed once per step
    for the whole batch; a missing control is the reference's dimension error."""
    gk = _gpu()
    n, m, c, nf, steps = 32, 8, 2, 12, 15
    f = fx.synth_lti(n, m, seed=21)
    rng = np.random.default_rng(210)
    G = rng.standard_normal((n, c))
    y = rng.standard_normal((steps, m, nf))
    u = rng.standard_normal((steps, c))
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], G, f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=nf)
    assert kf._fm
    est = kf.UpdateBatch(y, u, every_step=True, want=("state", "covar", "innov"))
    for fi in (0, nf - 1):
        o = oracle.NewVanilla(f["x0"], f["P0"], f["F"], G, f["H"], f["Q"], f["R"])
        refs = [o.Update(y[k, :, fi], u[k]) for k in range(steps)]
        for k in (0, steps - 1):
            assert fx.scaled_err(np.asarray(est.State())[k][:, fi], refs[k].State()) <= TOL
            assert fx.scaled_err(np.asarray(est.Innovation())[k][:, fi], refs[k].Innovation()) <= TOL
            assert fx.scaled_err(np.asarray(est.Covariance())[k][:, :, fi], refs[k].Covariance()) <= TOL
    with pytest.raises(gk.GkbError):
        kf.UpdateBatch(y, None, every_step=False)


def test_tile_vanilla_setters_between_calls(oracle):
    """SetStateTransition / SetMeasurementMatrix / SetNoise on a large-state handle between two batched calls
    (the jerkcar example's pattern, examples/jerkcar/main.go:141-159, at n = 16): m switches 8 -> 3 -> 8."""
    gk = _gpu()
    n, nf = 16, 10
    fa, fb = fx.synth_lti(n, 8, seed=31), fx.synth_lti(n, 3, seed=32)
    rng = np.random.default_rng(33)
    ya, yb = rng.standard_normal((6, 8, nf)), rng.standard_normal((5, 3, nf))
    kf, _ = gk.NewVanilla(fa["x0"], fa["P0"], fa["F"], None, fa["H"], gk.NewNoiseless(fa["Q"], fa["R"]), n_filters=nf)
    o = oracle.NewVanilla(fa["x0"], fa["P0"], fa["F"], None, fa["H"], fa["Q"], fa["R"])
    kf.UpdateBatch(ya, None, every_step=False, want=("state",))
    for k in range(6):
        ref = o.Update(ya[k, :, 4], None)
    kf.SetStateTransition(fb["F"]); kf.SetMeasurementMatrix(fb["H"]); kf.SetNoise(gk.NewNoiseless(fb["Q"], fb["R"]))
    o.SetStateTransition(fb["F"]); o.SetMeasurementMatrix(fb["H"]); o.SetNoise(fb["Q"], fb["R"])
    est = kf.UpdateBatch(yb, None, every_step=False, want=("state", "covar", "gain"))
    for k in range(5):
        ref = o.Update(yb[k, :, 4], None)
    assert fx.scaled_err(np.asarray(est.State())[:, 4], ref.State()) <= TOL
    assert fx.scaled_err(np.asarray(est.Covariance())[:, :, 4], ref.Covariance()) <= TOL
    assert fx.scaled_err(np.asarray(est.Gain())[:, :, 4], ref.Gain()) <= TOL
    kf.SetMeasurementMatrix(fa["H"]); kf.SetNoise(gk.NewNoiseless(fa["Q"], fa["R"]))
    o.SetMeasurementMatrix(fa["H"]); o.SetNoise(fa["Q"], fa["R"])
    est = kf.UpdateBatch(ya[:2], None, every_step=False, want=("state",))
    for k in range(2):
        ref = o.Update(ya[k, :, 4], None)
    assert fx.scaled_err(np.asarray(est.State())[:, 4], ref.State()) <= TOL
