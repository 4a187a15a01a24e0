"""Pins for the oracle paths no runnable reference test covers (HybridKF, SmoothAll, BatchKF, Monte
Carlo + chi-square are "parity unpinned" by the reference's own tests -- hybrid_test.go / srif_test.go need the smd
propagator, AWGN is clock-seeded).  Each check below is INDEPENDENT of the restated function it pins:

  (i)   hybrid CKF == vanilla on an LTI model (Phi = F, Htilde = H, computed = 0, no control, Noiseless; Q = 0, or
        PreparePNT(Gamma = I) with the same Q): hybrid.go:114-182 and vanilla.go:149-205 are the same formulas, and
        vanilla is pinned by the reference's golden CSVs;
  (ii)  SmoothAll: Phi_{k+1} x_k = x_{k+1} and Phi_{k+1} P_k Phi_{k+1}^T = P_{k+1} on the smoothed history (the
        defining identity of hybrid.go:221-232);
  (iii) BatchKF against the normal equations solved by numpy (batch.go:49-57,64-79, R-not-inverse quirk kept);
  (iv)  Monte Carlo + chi-square means against the exact moments of an analytic linear-Gaussian recursion
        (tests/chi2_moments.py), inside a 5-sigma band, for a matched and for a mis-tuned tested filter.
CPU only: the same checks run on the GPU in tests/test_gpu_crosscheck.py."""
import numpy as np
import pytest

import fixtures as fx
from chi2_moments import chi2_moments


def _lti3(rng):
    n, m = 3, 2
    F = np.eye(n) + 0.05 * rng.standard_normal((n, n))
    H = rng.standard_normal((m, n))
    A = rng.standard_normal((n, n))
    Q = 1e-2 * (A @ A.T + n * np.eye(n))
    R = np.diag([0.05, 0.2])
    return F, H, Q, R, rng.standard_normal(n), np.diag([4.0, 2.0, 1.0])


@pytest.mark.parametrize("snc", [False, True])
def test_hybrid_ckf_equals_vanilla_on_lti(oracle, snc):
    rng = np.random.default_rng(5)
    F, H, Q, R, x0, P0 = _lti3(rng)
    n, m, steps = 3, 2, 50
    Qv = Q if snc else np.zeros((n, n))
    v = oracle.NewVanilla(x0, P0, F, None, H, Qv, R)
    h = oracle.NewHybridKF(x0, P0, Q, R, m)
    ys = rng.standard_normal((steps, m))
    for k in range(steps):
        ev = v.Update(ys[k])
        h.Prepare(F, H)
        if snc:
            h.PreparePNT(np.eye(n))  # Gamma Q Gamma^T = Q exactly
        eh = h.UpdateNL(ys[k], np.zeros(m))
        for name in ("State", "Covariance", "PredCovariance", "Gain", "Innovation"):
            a, b = np.asarray(getattr(eh, name)()), np.asarray(getattr(ev, name)())
            assert fx.scaled_err(a, b) <= 1e-13, (k, name, fx.scaled_err(a, b))


def test_smooth_all_identity(oracle):
    rng = np.random.default_rng(8)
    n, steps = 6, 30
    Phi = np.eye(n)[None] + 0.05 * rng.standard_normal((steps, n, n))
    x = rng.standard_normal((steps, n))
    A = rng.standard_normal((steps, n, n))
    P = A @ A.transpose(0, 2, 1) + n * np.eye(n)[None]
    xs, Ps = oracle.smooth_all(Phi, x, P)
    assert np.array_equal(xs[-1], x[-1]) and np.array_equal(Ps[-1], P[-1])  # the last estimate is the anchor
    for k in range(steps - 1):
        assert fx.scaled_err(Phi[k + 1] @ xs[k], xs[k + 1]) <= 1e-12, k
        assert fx.scaled_err(Phi[k + 1] @ Ps[k] @ Phi[k + 1].T, Ps[k + 1]) <= 1e-11, k


def test_batch_kf_against_normal_equations(oracle):
    rng = np.random.default_rng(9)
    n, m, count = 6, 2, 40
    H = rng.standard_normal((count, m, n))
    real = rng.standard_normal((count, m))
    comp = real + 0.1 * rng.standard_normal((count, m))
    R = np.array([[0.5, 0.1], [0.1, 0.3]])
    x, P = oracle.batch_solve(R, H, real, comp)
    Lam = sum(H[k].T @ R @ H[k] for k in range(count))       # batch.go:50: R, not inv(R) (reference formula)
    Nv = sum(H[k].T @ R @ (real[k] - comp[k]) for k in range(count))
    assert fx.scaled_err(P, np.linalg.inv(Lam)) <= 1e-11
    assert fx.scaled_err(x, np.linalg.solve(Lam, Nv)) <= 1e-11


def _band_check(mean, exp_mean, exp_var, trials, tag, sigmas=5.0):
    z = (mean - exp_mean) / np.sqrt(exp_var / trials)
    assert np.max(np.abs(z)) <= sigmas, (tag, float(np.max(np.abs(z))), int(np.argmax(np.abs(z))))
    # and not trivially: the standardised deviations look like N(0, 1) noise, not like a bias
    assert abs(np.mean(z)) <= 5.0 / np.sqrt(len(z)) + 0.5, (tag, float(np.mean(z)))


@pytest.mark.parametrize("mistuned", [False, True])
def test_mc_chisquare_means_match_analytic_moments(oracle, mistuned):
    f = fx.robot_1d()
    steps, trials = 60, 20000
    controls = fx.robot_controls(steps)
    tested = dict(Q=0.2 * f["Q"], R=np.array([[0.08]]), F=f["F"] + np.array([[0, 0.01], [0, -0.02]])) if mistuned else None
    x0t = np.array([0.7, -0.4])
    ref = oracle.mc_chisquare(oracle.VANILLA, f["F"], f["G"], f["H"], f["Q"], f["R"], x0t, f["x0"], f["P0"], trials, steps,
                              controls=controls, seed=31, threads=4, tested=tested)
    mom = chi2_moments(f["F"], f["G"], f["H"], f["Q"], f["R"], x0t, steps, controls=controls, tested=tested,
                       x0_filter=f["x0"], P0=f["P0"])
    _band_check(ref["NEES"], mom["nees_mean"], mom["nees_var"], trials, "NEES")
    _band_check(ref["NIS"], mom["nis_mean"], mom["nis_var"], trials, "NIS")
    if mistuned:  # visibly inconsistent, and the analytic recursion predicts by how much
        assert np.mean(mom["nees_mean"][20:]) > 3.0
