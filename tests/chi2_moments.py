"""Exact first and second moments of the per-step NEES / NIS samples of NewMonteCarloRuns + NewChiSquare, by a
linear-Gaussian covariance recursion in numpy -- a pin that is INDEPENDENT of both the oracle's restatement and
the CUDA kernels (it never runs a filter on samples).

Experiment (montecarlo.go:92-119, chisquare.go:16-95, quirks kept):
  truth      x_{k+1} = F x_k + G u_k + w_k,   y_k = H x_k + v_k   (sample k pairs the state x_{k+1} with the
             measurement of x_k: they are one step apart),  w ~ N(0, Q), v ~ N(0, R), x_0 fixed
  tested KF  any vanilla model (Ft, Gt, Ht, Qt, Rt), Noiseless:  x-_k = Ft x^_{k-1} + Gt u_k,  nu_k = y_k - Ht x-_k,
             x^_k = x-_k + K_k nu_k,  deterministic (P-_k, K_k, P_k) from the Riccati recursion with Joseph update
  samples    NEES_k = d^T inv(P_k) d, d = x_{k+1} - x^_k;   NIS_k = nu^T inv(Ht P-_k Ht^T + Rt) nu
Everything is linear in the jointly Gaussian s_k = [x_k; x^_{k-1}], so mean and covariance of d and nu are exact,
and for q = z^T A z with z ~ N(mu, S):  E q = tr(A S) + mu^T A mu,  Var q = 2 tr((A S)^2) + 4 mu^T A S A mu.
The mean over N trials then lies within c sqrt(Var q / N) of E q.
"""
import numpy as np


def _quad_moments(A, mu, S):
    AS = A @ S
    mean = np.trace(AS) + mu @ A @ mu
    var = 2.0 * np.trace(AS @ AS) + 4.0 * mu @ A @ S @ A @ mu
    return float(mean), float(max(var, 0.0))


def chi2_moments(F, G, H, Q, R, x0_truth, steps, controls=None, tested=None, x0_filter=None, P0=None):
    """Returns dict(nees_mean, nees_var, nis_mean, nis_var), each [steps]."""
    F, H, Q, R = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (F, H, Q, R))
    n, m = F.shape[0], H.shape[0]
    G = None if G is None else np.asarray(G, dtype=np.float64).reshape(n, -1)
    t = dict(F=F, G=G, H=H, Q=Q, R=R)
    for k, v in (tested or {}).items():
        if v is not None:
            t[k] = np.atleast_2d(np.asarray(v, dtype=np.float64))
    Ft, Gt, Ht, Qt, Rt = t["F"], t["G"], t["H"], t["Q"], t["R"]
    if Gt is not None:
        Gt = Gt.reshape(n, -1)
    c = 0 if G is None else G.shape[1]
    u = np.zeros((steps, max(c, 1))) if controls is None else np.asarray(controls, dtype=np.float64).reshape(steps, -1)
    P = np.asarray(P0, dtype=np.float64)
    mu = np.concatenate([np.asarray(x0_truth, dtype=np.float64), np.asarray(x0_filter, dtype=np.float64)])
    S = np.zeros((2 * n, 2 * n))
    out = {k: np.zeros(steps) for k in ("nees_mean", "nees_var", "nis_mean", "nis_var")}
    I = np.eye(n)
    for k in range(steps):
        Pm = Ft @ P @ Ft.T + Qt
        Sk = Ht @ Pm @ Ht.T + Rt
        K = Pm @ Ht.T @ np.linalg.inv(Sk)
        A = I - K @ Ht
        P = A @ Pm @ A.T + K @ Rt @ K.T
        gu = np.zeros(n) if G is None else G @ u[k, :c]
        gut = np.zeros(n) if Gt is None else Gt @ u[k, :Gt.shape[1]]
        # nu = [H, -Ht Ft] s + v - Ht gut
        Mnu = np.hstack([H, -Ht @ Ft])
        nu_mu = Mnu @ mu - Ht @ gut
        nu_S = Mnu @ S @ Mnu.T + R
        out["nis_mean"][k], out["nis_var"][k] = _quad_moments(np.linalg.inv(Sk), nu_mu, nu_S)
        # x^_k = A Ft x^_{k-1} + A gut + K H x_k + K v ;  x_{k+1} = F x_k + gu + w
        Mx = np.vstack([np.hstack([F, np.zeros((n, n))]), np.hstack([K @ H, A @ Ft])])
        b = np.concatenate([gu, A @ gut])
        N = np.zeros((2 * n, 2 * n))
        N[:n, :n] = Q
        N[n:, n:] = K @ R @ K.T
        mu = Mx @ mu + b
        S = Mx @ S @ Mx.T + N
        D = np.hstack([I, -I])  # d = x_{k+1} - x^_k
        out["nees_mean"][k], out["nees_var"][k] = _quad_moments(np.linalg.inv(P), D @ mu, D @ S @ D.T)
    return out
