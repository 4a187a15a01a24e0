"""Model fixtures taken from the reference's tests and examples (values only)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def robot_1d():
    """examples/robot/main.go:16-26, helper_test.go:10-15"""
    dt = 0.1
    return dict(F=np.array([[1, dt], [0, 1.0]]), G=np.array([[0.5 * dt * dt], [dt]]), H=np.array([[1.0, 0]]),
                R=np.array([[0.05]]), Q=np.array([[5e-2, 5e-4], [5e-4, 1e-3]]), x0=np.zeros(2), P0=2.0 * np.eye(2),
                dt=dt)


def robot_controls(steps):
    """examples/robot/main.go:36-39"""
    return np.array([[np.cos(0.75 * (k + 1) * 0.1)] for k in range(steps)])


def jerk3():
    """helper_test.go:17-22 (Midterm2Matrices) + montecarlo_test.go:12-19"""
    dt = 0.01
    F = np.array([[1, 0.01, 5e-5], [0, 1, 0.01], [0, 0, 1.0]])
    G = np.array([[(5e-7) / 3], [5e-5], [0.01]])
    Q = np.array([[2.5e-15, 6.25e-13, (25e-11) / 3], [6.25e-13, (5e-7) / 3, 2.5e-8], [(25e-11) / 3, 2.5e-8, 5e-6]])
    R = np.array([[0.005 / dt]])
    H = np.array([[1.0, 0, 0]])
    return dict(F=F, G=G, H=H, Q=Q, R=R, x0=np.array([0, 0.35, 0]), P0=10.0 * np.eye(3), dt=dt)


def jerkcar4():
    """examples/jerkcar/main.go:94-109,118-119"""
    F = np.array([[1, 0.01, 0.00005, 0], [0, 1, 0.01, 0], [0, 0, 1, 0], [0, 0, 0, 1.0005125020836]])
    G = np.array([[0.0], [0.0001], [0.01], [0.0]])
    H1 = np.array([[1.0, 0, 0, 0], [0, 0, 1, 1]])
    H2 = np.array([[0.0, 0, 1, 1]])
    Q = np.array([[0.0000000000025, 0.000000000625, 0.000000083333333, 0],
                  [0.000000000625, 0.000000166666667, 0.000025, 0],
                  [0.000000083333333, 0.000025, 0.005, 0],
                  [0, 0, 0, 0.530265088355421]]) * 1e-3
    R = np.array([[0.5, 0], [0, 0.05]])
    Ra = np.array([[0.05]])
    return dict(F=F, G=G, H1=H1, H2=H2, Q=Q, R=R, Ra=Ra, x0=np.array([0, 0.45, 0, 0.09]), P0=10.0 * np.eye(4))


def multid4():
    """vanilla_test.go:97-115 (4-state / 2-measurement system)"""
    F = np.array([[1, 0.01, 0.00005, 0], [0, 1, 0.01, 0], [0, 0, 1, 0], [0, 0, 0, 1.0005]])
    G = np.array([[0.0], [0.0001], [0.01], [0.0]])
    H = np.array([[1.0, 0, 0, 0], [0, 0, 1, 1]])
    Q = np.array([[0.0000000000025, 0.000000000625, 0.000000083333333, 0],
                  [0.000000000625, 0.000000166666667, 0.000025, 0],
                  [0.000000083333333, 0.000025, 0.005, 0],
                  [0, 0, 0, 0.530265088355421]]) * 1e-3
    R = np.array([[0.5, 0], [0, 0.05]])
    return dict(F=F, G=G, H=H, Q=Q, R=R, x0=np.array([0, 0.35, 0, 0]), P0=10.0 * np.eye(4))


def statod4():
    """examples/statOD5044/main.go:36-61: the 4-state / 2-control / 2-measurement DT system of the statOD
    example (dt = 0.1), R = diag(2e-3, 2e-5)/dt, the feedback gain T with Fcl = F - G T, x0, P0."""
    dt = 0.1
    F = np.array([[1, 0.1, 0, 7.726e-2], [4.015e-7, 1, 0, 1.545], [-2.319e-16, -1.732e-9, 1, 0.1],
                  [-6.956e-15, -3.465e-8, 0, 1]])
    G = np.array([[5e-3, 3.85e-7], [0.1, 1.157e-5], [-5.775e-11, 7.487e-7], [1.732e-9, 1.498e-5]])
    H = np.array([[1.0, 0, 0, 0], [0, 0, 1, 0]])
    # mat64.NewSymDense keeps the upper triangle of the literal (main.go:42)
    Qlit = np.array([[6.669e-16, 1.001e-14, 3.823e-19, 5.150e-18], [1.001e-14, 2.002e-13, 1.030e-17, 1.545e-16],
                     [3.862e-19, 1.030e-17, 6.667e-19, 1.000e-17], [5.150e-18, 1.545e-16, 1.000e-17, 2.000e-16]])
    Q = np.triu(Qlit) + np.triu(Qlit, 1).T
    R = np.diag([2e-3, 2e-5]) / dt
    T = np.array([[0.930124736616832, 1.395260337125255, -0.000008568056356, 15.440297905873823],
                  [0.000001749639349, 0.000000859493456, 0.001999922457941, 5.177881640687808]])
    return dict(F=F, G=G, H=H, Q=Q, R=R, T=T, Fcl=F - G @ T, x0=np.array([2, 0.5, 0, 0.0]),
                P0=np.diag([5, 1, 0.01, 0.00001]), dt=dt)


def load_jerkcar_golden():
    d = np.load(os.path.join(GOLDEN_DIR, "jerkcar.npz"))
    return {k: d[k] for k in d.files}


def run_jerkcar(make_filters, steps=2000):
    """Replays examples/jerkcar/main.go:133-161 against filters built by `make_filters(fixture)`.

    `make_filters` returns a list of (name, filter, est0) where filter exposes Update /
    SetMeasurementMatrix / SetNoise like the Go LDKF interface (the oracle and the CUDA engine's
    Python mirror both do).  Returns {name: array[steps+1, 12]} in the CSV exporter's layout
    (exporter.go:34-45: value, +2 sigma, -2 sigma per state component)."""
    g = load_jerkcar_golden()
    fx = jerkcar4()
    yacc, ypos, uvec = g["yacc"], np.nan_to_num(g["ypos"], nan=0.0), g["uvec"]
    filters = make_filters(fx)
    rows = {name: [csv_row(est0)] for name, _, est0 in filters}
    for k in range(min(steps, len(yacc))):
        for name, kf, _ in filters:
            if (k + 1) % 10 == 0:
                kf.SetMeasurementMatrix(fx["H1"])
                kf.SetNoise(fx["Q"], fx["R"])
                y = np.array([ypos[k], yacc[k]])
            else:
                y = np.array([yacc[k]])
            est = kf.Update(y, np.array([uvec[k]]))
            rows[name].append(csv_row(est))
            if (k + 1) % 10 == 0:
                kf.SetMeasurementMatrix(fx["H2"])
                kf.SetNoise(fx["Q"], fx["Ra"])
    return {k: np.array(v) for k, v in rows.items()}


def csv_row(est):
    x = np.asarray(est.State())
    P = np.asarray(est.Covariance())
    row = []
    for i in range(len(x)):
        b = 2.0 * np.sqrt(P[i, i])
        row += [x[i], b, -b]
    return row


def scaled_err(a, ref):
    """SURVEY 8(c) parity metric: max |a-ref| / max(|ref|_entry, ||ref||_max), per array."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    floor = np.max(np.abs(ref))
    if floor == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor)))


def scaled_err_steps(a, ref):
    """"for each output array A AT EACH STEP, |A_gpu - A_ref| <= tol * max(|A_ref|_entry,
    ||A_ref||_max)".  Axis 0 of `a` / `ref` is the step; every step is scaled by ITS OWN max-abs (a covariance
    that shrinks 100x over a run is held to the same relative bar at the end as at the start).  Returns the
    worst step's error."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    a, ref = a.reshape(a.shape[0], -1), ref.reshape(ref.shape[0], -1)
    floor = np.max(np.abs(ref), axis=1, keepdims=True)
    if ref.shape[1] == 1:
        # a ONE-entry array (m = 1: Measurement(), Innovation()) has no max-abs other than the entry itself, so
        # per-step scaling would be the bare per-entry relative error, which SURVEY section 7 calls ill-posed (the
        # entry crosses zero while its absolute error is set by the state's scale): its floor is the run's max-abs
        floor = np.full_like(floor, np.max(np.abs(ref)))
    d = np.abs(a - ref)
    den = np.maximum(np.abs(ref), floor)
    err = np.where(den > 0.0, d / np.where(den > 0.0, den, 1.0), d)
    return float(np.max(err))


def strict_rel_err(a, ref, tol=1e-10):
    """Strict per-entry relative error |a - ref| / |ref| (no floor) with the entries above `tol` listed:
    those are the zero crossings the SURVEY asks to name (entries whose magnitude is far below the array's).
    Returns (max over entries with ref != 0, [(flat index, ref value, rel err), ...] for rel err > tol)."""
    a, ref = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(ref, dtype=np.float64).reshape(-1)
    nz = ref != 0.0
    if not np.any(nz):
        return 0.0, []
    rel = np.zeros_like(ref)
    rel[nz] = np.abs(a[nz] - ref[nz]) / np.abs(ref[nz])
    bad = [(int(i), float(ref[i]), float(rel[i])) for i in np.nonzero(rel > tol)[0][:16]]
    return float(rel.max()), bad


def synth_lti(n, m, seed=5):
    """BASELINE configs[4]: seeded synthetic LTI model, shared by all filters.
    F = I + dt A with A ~ U(-1, 1)/n; Q = 1e-3 L L^T; R = diag U(0.1, 1); H = first m rows of a seeded
    orthogonal matrix; P0 = I; x0 = 0."""
    rng = np.random.default_rng(seed)
    dt = 0.1
    F = np.eye(n) + dt * rng.uniform(-1.0, 1.0, (n, n)) / n
    L = np.tril(rng.uniform(-1.0, 1.0, (n, n))) + np.eye(n)
    Q = 1e-3 * (L @ L.T)
    Q = 0.5 * (Q + Q.T)
    R = np.diag(rng.uniform(0.1, 1.0, m))
    Hfull, _ = np.linalg.qr(rng.standard_normal((n, n)))
    H = np.ascontiguousarray(Hfull[:m])
    return dict(F=F, G=None, H=H, Q=Q, R=R, x0=np.zeros(n), P0=np.eye(n))
