"""ctypes binding of the CPU oracle (oracle/gko.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (gokalman_b200/) must never import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgko.so")

MAXN, MAXM = 64, 16
VANILLA, PREDICTOR, INFORMATION, SQRT, HYBRID, SRIF = range(6)

ERR_NAMES = {0: "ok", -1: "dims", -2: "singular_S", -3: "asymmetric", -4: "locked", -5: "singular_Phi",
             -6: "singular_R", -7: "noise_range"}


class Estimate(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("m", C.c_int),
        ("state", C.c_double * MAXN),
        ("meas", C.c_double * MAXM),
        ("innov", C.c_double * MAXN),
        ("innov_len", C.c_int),
        ("covar", C.c_double * (MAXN * MAXN)),
        ("pred_covar", C.c_double * (MAXN * MAXN)),
        ("gain", C.c_double * (MAXN * MAXM)),
        ("obs_dev", C.c_double * MAXM),
        ("raw_vec", C.c_double * MAXN),
        ("raw_mat", C.c_double * (MAXN * MAXN)),
        ("raw_pred_mat", C.c_double * (MAXN * MAXN)),
        ("covar_ok", C.c_int), ("pred_covar_ok", C.c_int),
    ]

    def _vec(self, field, k):
        return np.frombuffer(getattr(self, field), dtype=np.float64, count=k).copy()

    def State(self):
        return self._vec("state", self.n)

    def Measurement(self):
        return self._vec("meas", self.m)

    def Innovation(self):
        return self._vec("innov", self.innov_len)

    def ObservationDev(self):
        return self._vec("obs_dev", self.m)

    def Covariance(self):
        return self._vec("covar", self.n * self.n).reshape(self.n, self.n)

    def PredCovariance(self):
        return self._vec("pred_covar", self.n * self.n).reshape(self.n, self.n)

    def Gain(self):
        return self._vec("gain", self.n * self.m).reshape(self.n, self.m)

    def raw(self):
        n = self.n
        return (self._vec("raw_vec", n), self._vec("raw_mat", n * n).reshape(n, n),
                self._vec("raw_pred_mat", n * n).reshape(n, n))


class McConfig(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("m", C.c_int), ("c", C.c_int), ("kind", C.c_int),
        ("F", C.c_void_p), ("G", C.c_void_p), ("H", C.c_void_p), ("Q", C.c_void_p), ("R", C.c_void_p),
        ("x0_truth", C.c_void_p), ("x0_filter", C.c_void_p), ("P0", C.c_void_p),
        ("trials", C.c_int), ("steps", C.c_int),
        ("controls", C.c_void_p), ("w", C.c_void_p), ("v", C.c_void_p),
        ("seed", C.c_uint64), ("trial_offset", C.c_int64),
        ("with_nees", C.c_int), ("with_nis", C.c_int), ("threads", C.c_int),
        ("tF", C.c_void_p), ("tG", C.c_void_p), ("tH", C.c_void_p), ("tQ", C.c_void_p), ("tR", C.c_void_p),
    ]


def build(force=False):
    """Compile oracle/_build/libgko.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("gko.c", "gko_linalg.c", "gko_od.c", "gko_c2d.c", "gko.h", "gko_linalg.h", "Makefile",
                                               os.path.join("..", "include", "gokalman_b200_icdf.inc"))]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    dp, ip = C.c_void_p, C.c_int
    L.gko_new_vanilla.restype = C.c_void_p
    L.gko_new_vanilla.argtypes = [ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, ip]
    L.gko_new_information.restype = C.c_void_p
    L.gko_new_information.argtypes = [ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]
    L.gko_new_information_from_state.restype = C.c_void_p
    L.gko_new_information_from_state.argtypes = [ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]
    L.gko_new_sqrt.restype = C.c_void_p
    L.gko_new_sqrt.argtypes = [ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]
    L.gko_new_hybrid.restype = C.c_void_p
    L.gko_new_hybrid.argtypes = [ip, ip, ip, dp, dp, dp, dp]
    L.gko_new_srif.restype = C.c_void_p
    L.gko_new_srif.argtypes = [ip, ip, dp, dp, dp, ip]
    L.gko_free.argtypes = [C.c_void_p]
    L.gko_set_state_transition.argtypes = [C.c_void_p, dp]
    L.gko_set_input_control.argtypes = [C.c_void_p, ip, dp]
    L.gko_set_measurement_matrix.argtypes = [C.c_void_p, ip, dp]
    L.gko_set_noise.argtypes = [C.c_void_p, dp, ip, dp]
    L.gko_set_replay.argtypes = [C.c_void_p, ip, dp, dp, ip]
    L.gko_reset.argtypes = [C.c_void_p]
    L.gko_initial_estimate.argtypes = [C.c_void_p, C.POINTER(Estimate)]
    L.gko_update.argtypes = [C.c_void_p, dp, dp, C.POINTER(Estimate)]
    L.gko_prepare.argtypes = [C.c_void_p, dp, dp]
    L.gko_prepare_pnt.argtypes = [C.c_void_p, dp]
    L.gko_enable_ekf.argtypes = [C.c_void_p, ip]
    L.gko_nl_predict.argtypes = [C.c_void_p, C.POINTER(Estimate)]
    L.gko_nl_update.argtypes = [C.c_void_p, dp, dp, C.POINTER(Estimate)]
    L.gko_srif_set_non_tri_r.argtypes = [C.c_void_p, ip]
    L.gko_measurement_srif_update.argtypes = [ip, ip, dp, dp, dp, dp, dp, dp, dp]
    L.gko_smooth_all.argtypes = [ip, ip, dp, dp, dp]
    L.gko_run_nl_batch.argtypes = [ip, ip, ip, C.c_int64, ip, C.c_void_p, dp, dp, dp, dp, dp, dp, dp, ip, dp, dp]
    L.gko_run_vanilla_batch.argtypes = [ip, ip, C.c_int64, ip, dp, dp, dp, dp, dp, dp, dp, ip, dp, dp]
    L.gko_batch_solve.argtypes = [ip, ip, ip, dp, dp, dp, dp, dp, dp]
    L.gko_mc_chisquare.argtypes = [C.POINTER(McConfig), dp, dp, dp, dp, dp, dp]
    L.gko_philox4x32_10.argtypes = [dp, dp, dp]
    L.gko_icdf_normal.argtypes = [C.c_uint32]
    L.gko_icdf_normal.restype = C.c_double
    L.gko_philox_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, ip, dp]
    for name in ("gko_inverse",):
        getattr(L, name).argtypes = [dp, dp, ip, dp]
    L.gko_chol_lower.argtypes = [dp, dp, ip]
    L.gko_qr_r.argtypes = [dp, dp, ip, ip]
    L.gko_householder_transf.argtypes = [dp, ip, ip]
    L.gko_as_sym.argtypes = [dp, ip]
    _lib = L
    return L


def _a(x):
    """contiguous float64 array (kept alive by the caller) or None"""
    if x is None:
        return None
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__("oracle error %d (%s)" % (code, ERR_NAMES.get(code, "?")))
        self.code = code


class Filter:
    """Stateful oracle filter mirroring the Go objects (LDKF and NLDKF)."""

    def __init__(self, handle, kind, n, m):
        if not handle:
            raise OracleError(-1)
        self.h, self.kind, self.n, self.m = handle, kind, n, m

    def __del__(self):
        if getattr(self, "h", None):
            lib().gko_free(self.h)
            self.h = None

    # -- LDKF
    def Update(self, y, u=None):
        est = Estimate()
        ya, ua = _a(y), _a(u)
        rc = lib().gko_update(self.h, _p(ya), _p(ua), C.byref(est))
        if rc != 0:
            raise OracleError(rc)
        return est

    def SetStateTransition(self, F):
        lib().gko_set_state_transition(self.h, _p(_a(F)))

    def SetInputControl(self, G):
        G = _a(G)
        lib().gko_set_input_control(self.h, G.shape[1], _p(G))

    def SetMeasurementMatrix(self, H):
        H = np.atleast_2d(_a(H))
        self.m = H.shape[0]
        lib().gko_set_measurement_matrix(self.h, H.shape[0], _p(H))

    def SetNoise(self, Q, R):
        R = np.atleast_2d(_a(R))
        lib().gko_set_noise(self.h, _p(_a(Q)), R.shape[0], _p(R))

    def SetReplayNoise(self, w, v, w2=None):
        """w2: what the second Process(k) call of Vanilla.Update returns (AWGN draws afresh); None = w again."""
        w, v = _a(w), _a(v)
        steps = (w if w is not None else v).shape[0]
        mv = v.shape[1] if v is not None else self.m
        lib().gko_set_replay(self.h, steps, _p(w), _p(v), mv)
        if w2 is not None:
            L = lib()
            L.gko_set_replay_second_draw.argtypes = [C.c_void_p, C.c_void_p]
            L.gko_set_replay_second_draw.restype = None
            L.gko_set_replay_second_draw(self.h, _p(_a(w2)))

    def Reset(self):
        lib().gko_reset(self.h)

    def InitialEstimate(self):
        est = Estimate()
        lib().gko_initial_estimate(self.h, C.byref(est))
        return est

    # -- NLDKF
    def Prepare(self, Phi, Htilde):
        lib().gko_prepare(self.h, _p(_a(Phi)), _p(_a(Htilde)))

    def PreparePNT(self, Gamma):
        lib().gko_prepare_pnt(self.h, _p(_a(Gamma)))

    def EnableEKF(self):
        lib().gko_enable_ekf(self.h, 1)

    def DisableEKF(self):
        lib().gko_enable_ekf(self.h, 0)

    def Predict(self):
        est = Estimate()
        rc = lib().gko_nl_predict(self.h, C.byref(est))
        if rc != 0:
            raise OracleError(rc)
        return est

    def UpdateNL(self, real_obs, computed_obs):
        est = Estimate()
        rc = lib().gko_nl_update(self.h, _p(_a(real_obs)), _p(_a(computed_obs)), C.byref(est))
        if rc != 0:
            raise OracleError(rc)
        return est


def _dims(F, G, H):
    F, H = _a(F), np.atleast_2d(_a(H))
    n, m = F.shape[0], H.shape[0]
    G = None if G is None else _a(G).reshape(n, -1)
    c = 0 if G is None else G.shape[1]
    return F, G, H, n, m, c


def NewVanilla(x0, P0, F, G, H, Q, R, predictor=False):
    F, G, H, n, m, c = _dims(F, G, H)
    h = lib().gko_new_vanilla(n, m, c, _p(_a(x0)), _p(_a(P0)), _p(F), _p(G), _p(H), _p(_a(Q)),
                              _p(np.atleast_2d(_a(R))), int(predictor))
    return Filter(h, PREDICTOR if predictor else VANILLA, n, m)


def NewInformation(i0, I0, F, G, H, Q, R):
    F, G, H, n, m, c = _dims(F, G, H)
    h = lib().gko_new_information(n, m, c, _p(_a(i0)), _p(_a(I0)), _p(F), _p(G), _p(H), _p(_a(Q)),
                                  _p(np.atleast_2d(_a(R))))
    return Filter(h, INFORMATION, n, m)


def NewInformationFromState(x0, P0, F, G, H, Q, R):
    F, G, H, n, m, c = _dims(F, G, H)
    h = lib().gko_new_information_from_state(n, m, c, _p(_a(x0)), _p(_a(P0)), _p(F), _p(G), _p(H),
                                             _p(_a(Q)), _p(np.atleast_2d(_a(R))))
    return Filter(h, INFORMATION, n, m)


def NewSquareRoot(x0, P0, F, G, H, Q, R):
    F, G, H, n, m, c = _dims(F, G, H)
    h = lib().gko_new_sqrt(n, m, c, _p(_a(x0)), _p(_a(P0)), _p(F), _p(G), _p(H), _p(_a(Q)),
                           _p(np.atleast_2d(_a(R))))
    return Filter(h, SQRT, n, m)


def NewHybridKF(x0, P0, Q, R, meas_size):
    x0 = _a(x0)
    n = x0.shape[0]
    Q = None if Q is None else np.atleast_2d(_a(Q))
    q = 0 if Q is None else Q.shape[0]
    h = lib().gko_new_hybrid(n, meas_size, q, _p(x0), _p(_a(P0)), _p(Q), _p(np.atleast_2d(_a(R))))
    return Filter(h, HYBRID, n, meas_size)


def NewSRIF(x0, P0, meas_size, non_tri_r, R):
    x0 = _a(x0)
    n = x0.shape[0]
    h = lib().gko_new_srif(n, meas_size, _p(x0), _p(_a(P0)), _p(np.atleast_2d(_a(R))), int(non_tri_r))
    return Filter(h, SRIF, n, meas_size)


def measurement_srif_update(R, H, b, y):
    R, H, b, y = _a(R), np.atleast_2d(_a(H)), _a(b), _a(y)
    n, m = R.shape[0], H.shape[0]
    Rk, bk, ek = np.zeros((n, n)), np.zeros(n), np.zeros(m)
    lib().gko_measurement_srif_update(n, m, _p(R), _p(H), _p(b), _p(y), _p(Rk), _p(bk), _p(ek))
    return Rk, bk, ek


def householder_transf(A, n, m):
    A = _a(A).copy()
    lib().gko_householder_transf(_p(A), n, m)
    return A


def inverse(A):
    A = _a(A)
    n = A.shape[0]
    out = np.zeros((n, n))
    cond = C.c_double(0.0)
    rc = lib().gko_inverse(_p(out), _p(A), n, C.cast(C.byref(cond), C.c_void_p))
    return out, rc, cond.value


def chol_lower(A):
    A = _a(A)
    n = A.shape[0]
    L = np.zeros((n, n))
    ok = lib().gko_chol_lower(_p(L), _p(A), n)
    return L, bool(ok)


def qr_r(A):
    A = _a(A)
    R = np.zeros_like(A)
    lib().gko_qr_r(_p(R), _p(A), A.shape[0], A.shape[1])
    return R


def smooth_all(Phi, x, P):
    Phi, x, P = _a(Phi), _a(x).copy(), _a(P).copy()
    steps, n = x.shape
    rc = lib().gko_smooth_all(n, steps, _p(Phi), _p(x), _p(P))
    if rc != 0:
        raise OracleError(rc)
    return x, P


_fma_lib = None


def fma_lib():
    """The FMA-contracted build of the same restatement (oracle/Makefile: libgko_fma.so) -- a rounding-sensitivity
    probe, never the parity oracle."""
    global _fma_lib
    if _fma_lib is None:
        path = os.path.join(_HERE, "_build", "libgko_fma.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(path)
        dp, ip = C.c_void_p, C.c_int
        L.gko_run_nl_batch.argtypes = [ip, ip, ip, C.c_int64, ip, C.c_void_p, dp, dp, dp, dp, dp, dp, dp, ip, dp, dp]
        _fma_lib = L
    return _fma_lib


def run_nl_batch(kind, x0, P0, R, flags, Phi, Htilde, real_obs, computed_obs, threads=1, fma=False):
    """Hybrid / SRIF over SoA streams Phi [steps, n*n, nf], Htilde [steps, m*n, nf], observations [steps, m, nf]:
    last State() [n, nf] and Covariance() [n*n, nf] of every filter (OpenMP over filters).  fma=True runs the
    FMA-contracted build instead (rounding-sensitivity probe)."""
    Phi, Htilde, real_obs, computed_obs = _a(Phi), _a(Htilde), _a(real_obs), _a(computed_obs)
    steps, nn, nf = Phi.shape
    n = int(round(nn ** 0.5))
    m = real_obs.shape[1]
    flags = None if flags is None else np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
    xs, Ps = np.zeros((n, nf)), np.zeros((n * n, nf))
    rc = (fma_lib() if fma else lib()).gko_run_nl_batch(kind, n, m, nf, steps, None if flags is None else flags.ctypes.data, _p(_a(x0)), _p(_a(P0)),
                                _p(np.atleast_2d(_a(R))), _p(Phi), _p(Htilde), _p(real_obs), _p(computed_obs), threads,
                                _p(xs), _p(Ps))
    if rc != 0:
        raise OracleError(rc)
    return xs, Ps


def run_vanilla_batch(x0, P0, F, H, Q, R, y, threads=1, want_covar=True):
    """Vanilla (Noiseless, no control) over y [steps, nf, m]: last State() [nf, n], Covariance() [nf, n*n]."""
    F, H, y = _a(F), np.atleast_2d(_a(H)), _a(y)
    n, m = F.shape[0], H.shape[0]
    steps, nf = y.shape[0], y.shape[1]
    xs = np.zeros((nf, n))
    Ps = np.zeros((nf, n * n)) if want_covar else None
    rc = lib().gko_run_vanilla_batch(n, m, nf, steps, _p(_a(x0)), _p(_a(P0)), _p(F), _p(H), _p(_a(Q)), _p(np.atleast_2d(_a(R))),
                                     _p(y), threads, _p(xs), _p(Ps))
    if rc != 0:
        raise OracleError(rc)
    return xs, Ps


def batch_solve(R, H, real_obs, computed_obs):
    """batch.go:34-79: H [count, m, n], observations [count, m] -> (xHat0 [n], P0 [n, n])."""
    H, R = _a(H), np.atleast_2d(_a(R))
    count, m, n = H.shape
    real_obs, computed_obs = _a(real_obs).reshape(count, m), _a(computed_obs).reshape(count, m)
    x, P = np.zeros(n), np.zeros((n, n))
    rc = lib().gko_batch_solve(n, m, count, _p(R), _p(H), _p(real_obs), _p(computed_obs), _p(x), _p(P))
    if rc != 0:
        raise OracleError(rc)
    return x, P


def icdf_normal(k):
    """The engine's Gaussian transform of one 32-bit word (piecewise-quintic inverse normal CDF)."""
    return lib().gko_icdf_normal(int(k) & 0xFFFFFFFF)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().gko_philox4x32_10(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                            o.ctypes.data_as(C.c_void_p))
    return o


def philox_normals(seed, trial, step, count):
    z = np.zeros(count)
    lib().gko_philox_normals(seed, trial, step, count, _p(z))
    return z


def mc_chisquare(kind, F, G, H, Q, R, x0_truth, x0_filter, P0, trials, steps, controls=None, w=None, v=None,
                 seed=0, trial_offset=0, with_nees=True, with_nis=True, threads=1, want_stats=False,
                 want_truth=False, tested=None):
    """montecarlo.go NewMonteCarloRuns + chisquare.go NewChiSquare. Returns dict with NIS/NEES means.
    tested: optional dict with the tested filter's OWN F / G / H / Q / R (chisquare.go:16 takes any LDKF); a
    missing key = the truth generator's matrix."""
    F, G, H, n, m, c = _dims(F, G, H)
    keep = [F, G, H, _a(Q), np.atleast_2d(_a(R)), _a(x0_truth), _a(x0_filter), _a(P0), _a(controls), _a(w), _a(v)]
    cfg = McConfig()
    cfg.n, cfg.m, cfg.c, cfg.kind = n, m, c, kind
    (cfg.F, cfg.G, cfg.H, cfg.Q, cfg.R, cfg.x0_truth, cfg.x0_filter, cfg.P0, cfg.controls, cfg.w,
     cfg.v) = [None if a is None else a.ctypes.data for a in keep]
    cfg.trials, cfg.steps = trials, steps
    cfg.seed, cfg.trial_offset = seed, trial_offset
    cfg.with_nees, cfg.with_nis, cfg.threads = int(with_nees), int(with_nis), threads
    for key in ("F", "G", "H", "Q", "R"):
        a = None if not tested or tested.get(key) is None else np.atleast_2d(_a(tested[key]))
        if a is not None:
            keep.append(a)
            setattr(cfg, "t" + key, a.ctypes.data)
    nis, nees = np.zeros(steps), np.zeros(steps)
    mean = np.zeros((steps, n)) if want_stats else None
    std = np.zeros((steps, n)) if want_stats else None
    tx = np.zeros((trials, steps, n)) if want_truth else None
    ty = np.zeros((trials, steps, m)) if want_truth else None
    rc = lib().gko_mc_chisquare(C.byref(cfg), _p(nis), _p(nees), _p(mean), _p(std), _p(tx), _p(ty))
    if rc != 0:
        raise OracleError(rc)
    return {"NIS": nis, "NEES": nees, "mean": mean, "std": std, "truth_x": tx, "truth_y": ty}


def od_synth(mu, j2, re, dt, orbit0, station, truth_obs, sigma_range, sigma_rate, seed, filter_offset=0):
    """gko_od.c: the engine's OD-input synthesis as the generic RK4 on state + STM.  orbit0 [6, nf]; station
    [steps, 6]; truth_obs [steps, 2].  Returns Phi [steps, 36, nf], Ht [steps, 12, nf], real, computed [steps, 2, nf],
    final orbits [6, nf]."""
    orbit0, station, truth_obs = _a(orbit0), _a(station), _a(truth_obs)
    nf, steps = orbit0.shape[1], station.shape[0]
    Phi, Ht = np.zeros((steps, 36, nf)), np.zeros((steps, 12, nf))
    real, comp, orb = np.zeros((steps, 2, nf)), np.zeros((steps, 2, nf)), np.zeros((6, nf))
    L = lib()
    L.gko_od_synth.restype = C.c_int
    L.gko_od_synth.argtypes = [C.c_double] * 4 + [C.c_int64, C.c_int] + [C.c_void_p] * 3 + [C.c_double, C.c_double,
                               C.c_uint64, C.c_int64] + [C.c_void_p] * 5
    rc = L.gko_od_synth(mu, j2, re, dt, nf, steps, _p(orbit0), _p(station), _p(truth_obs), sigma_range, sigma_rate, seed,
                        filter_offset, _p(Phi), _p(Ht), _p(real), _p(comp), _p(orb))
    if rc != 0:
        raise OracleError(rc)
    return Phi, Ht, real, comp, orb


def van_loan(A, Gamma, W, dt):
    """gko_c2d.c: c2d.go:13-75 (F, Q) of one system."""
    A, Gamma, W = np.atleast_2d(_a(A)), np.atleast_2d(_a(Gamma)), np.atleast_2d(_a(W))
    n, q = A.shape[0], W.shape[0]
    Gamma = np.ascontiguousarray(Gamma.reshape(n, q))
    F, Q = np.zeros((n, n)), np.zeros((n, n))
    L = lib()
    L.gko_van_loan.restype = C.c_int
    L.gko_van_loan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    rc = L.gko_van_loan(n, q, _p(A), _p(Gamma), _p(W), float(dt), _p(F), _p(Q))
    if rc != 0:
        raise OracleError(rc)
    return F, Q


def expm(A):
    A = np.atleast_2d(_a(A))
    E = np.zeros_like(A)
    L = lib()
    L.gko_expm.restype = C.c_int
    L.gko_expm.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    if L.gko_expm(_p(E), _p(A), A.shape[0]) != 0:
        raise OracleError(-1)
    return E
