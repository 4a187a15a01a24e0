"""The independent pins of tests/test_oracle_crosscheck.py, run on the CUDA path (through the C-ABI):
hybrid CKF == vanilla on an LTI model (GPU against GPU), the SmoothAll identity, BatchKF against numpy's normal
equations, and the Monte Carlo + chi-square means at 10^6 trials against the exact analytic moments."""
import numpy as np
import pytest

import fixtures as fx
from chi2_moments import chi2_moments
from test_oracle_crosscheck import _band_check, _lti3

pytestmark = pytest.mark.gpu


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


@pytest.mark.parametrize("snc", [False, True])
def test_gpu_hybrid_ckf_equals_gpu_vanilla_on_lti(snc):
    """SURVEY 8(c)(i) on the device: the hybrid kernels (production and strict) and the vanilla kernel compute the
    same filter when Phi = F, Htilde = H, computed = 0."""
    gk = _gpu()
    from gokalman_b200 import _lib as L
    rng = np.random.default_rng(5)
    F, H, Q, R, x0, P0 = _lti3(rng)
    n, m, steps, nf = 3, 2, 50, 7
    ys = rng.standard_normal((steps, m, nf))
    v, _ = gk.NewVanilla(x0, P0, F, None, H, gk.NewNoiseless(Q if snc else np.zeros((n, n)), R), n_filters=nf)
    ev = v.UpdateBatch(ys, None, every_step=True)
    flags = np.full(steps, L.F_MEAS | (L.F_SNC if snc else 0), dtype=np.uint8)
    Phi = np.repeat(F[None], steps, axis=0)
    Ht = np.repeat(H[None], steps, axis=0)
    Gamma = np.repeat(np.eye(n)[None], steps, axis=0) if snc else None
    for strict in (False, True):
        h, _ = gk.NewHybridKF(x0, P0, gk.NewNoiseless(Q, R), m, n_filters=nf)
        h.SetStrict(strict)
        eh = h.RunBatch(flags, Phi, Ht, ys, np.zeros((steps, m, nf)), Gamma, every_step=True)
        for name in ("State", "Covariance", "PredCovariance", "Gain", "Innovation"):
            a, b = np.asarray(getattr(eh, name)()), np.asarray(getattr(ev, name)())
            for f in range(nf):
                err = fx.scaled_err_steps(a[..., f], b[..., f])
                assert err <= 1e-12, (strict, name, f, err)


def test_gpu_smooth_all_identity():
    gk = _gpu()
    from gokalman_b200 import _lib as L
    rng = np.random.default_rng(8)
    n, m, steps, nf = 6, 2, 30, 40
    Phi = np.eye(n)[None, :, :, None] + 0.05 * rng.standard_normal((steps, n, n, nf))
    Ht = rng.standard_normal((steps, m, n, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.05 * rng.standard_normal((steps, m, nf))
    kf, _ = gk.NewHybridKF(np.zeros(n), np.diag([10, 10, 10, 1, 1, 1.0]), gk.NewNoiseless(np.diag([1e-12] * 3), np.diag([1e-2, 1e-2])), m,
                           n_filters=nf)
    est = kf.RunBatch(np.full(steps, L.F_MEAS, dtype=np.uint8), Phi, Ht, real, comp, None, every_step=True)
    last_x, last_P = est.State()[-1].copy(), est.Covariance()[-1].copy()
    kf.SmoothAll(est)
    xs, Ps = est.State(), est.Covariance()
    assert np.array_equal(xs[-1], last_x) and np.array_equal(Ps[-1], last_P)
    for f in (0, 13, nf - 1):
        for k in range(steps - 1):
            Pk = Phi[k + 1, :, :, f]
            assert fx.scaled_err(Pk @ xs[k, :, f], xs[k + 1, :, f]) <= 1e-11, (f, k)
            assert fx.scaled_err(Pk @ Ps[k, :, :, f] @ Pk.T, Ps[k + 1, :, :, f]) <= 1e-10, (f, k)


def test_gpu_batch_kf_against_normal_equations():
    gk = _gpu()
    rng = np.random.default_rng(9)
    n, m, count, nf = 6, 2, 40, 33
    H = rng.standard_normal((count, m, n, nf))
    real = rng.standard_normal((count, m, nf))
    comp = real + 0.1 * rng.standard_normal((count, m, nf))
    R = np.array([[0.5, 0.1], [0.1, 0.3]])
    bk = gk.NewBatchKF(count, gk.NewNoiseless(np.zeros((n, n)), R))
    x, P, status = bk.SolveBatch(H, real, comp, n_filters=nf)
    assert np.all(status == 0)
    for f in (0, 16, nf - 1):
        Lam = sum(H[k, :, :, f].T @ R @ H[k, :, :, f] for k in range(count))  # batch.go:50: R, not inv(R)
        Nv = sum(H[k, :, :, f].T @ R @ (real[k, :, f] - comp[k, :, f]) for k in range(count))
        assert fx.scaled_err(P[:, :, f], np.linalg.inv(Lam)) <= 1e-11
        assert fx.scaled_err(x[:, f], np.linalg.solve(Lam, Nv)) <= 1e-11


@pytest.mark.parametrize("mistuned", [False, True])
def test_gpu_mc_chisquare_means_match_analytic_moments_1e6(mistuned):
    """examples/robot/main.go:32-58 at 10^6 trials: the per-step NEES / NIS means of the fused kernel (Philox noise,
    truth, filter, reduction) against the EXACT expected values of the same experiment (analytic linear-Gaussian
    recursion, numpy), inside a 5-sigma confidence band of the sample mean (sigma ~ 2e-3) -- for the matched filter
    and for a mis-tuned one (own Q, R, F: chisquare.go:16).  Independent of the oracle."""
    gk = _gpu()
    f = fx.robot_1d()
    steps, trials = 120, 1000000
    controls = fx.robot_controls(steps)
    x0t = np.array([0.7, -0.4])
    tested = dict(Q=0.2 * f["Q"], R=np.array([[0.08]]), F=f["F"] + np.array([[0, 0.01], [0, -0.02]])) if mistuned else {}
    mckf, _ = gk.NewPurePredictorVanilla(x0t, f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=31))
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], tested.get("F", f["F"]), f["G"], f["H"],
                          gk.NewNoiseless(tested.get("Q", f["Q"]), tested.get("R", f["R"])))
    runs = gk.NewMonteCarloRuns(trials, steps, 1, list(controls), mckf)
    nis, nees = gk.NewChiSquare(kf, runs, list(controls), True, True)
    mom = chi2_moments(f["F"], f["G"], f["H"], f["Q"], f["R"], x0t, steps, controls=controls, tested=tested or None,
                       x0_filter=f["x0"], P0=f["P0"])
    _band_check(nees, mom["nees_mean"], mom["nees_var"], trials, "NEES")
    _band_check(nis, mom["nis_mean"], mom["nis_var"], trials, "NIS")
    if not mistuned:  # a consistent filter: NIS -> m = 1 once the prior is forgotten
        assert abs(np.mean(nis[40:]) - 1.0) < 5e-3
