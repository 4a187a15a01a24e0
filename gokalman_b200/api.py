"""Host-side mirror of the gokalman Go API over the CUDA engine's C-ABI.

Same names and argument meaning as the reference (the Go toolchain is absent here, so this Python
layer is the tested caller; go/gokalman_b200.go carries the cgo shim as source):

    NewVanilla / NewPurePredictorVanilla    vanilla.go:21-62
    NewInformation / NewInformationFromState information.go:20-81
    NewSquareRoot                            squareroot.go:21-50
    NewHybridKF                              hybrid.go:23-34
    NewSRIF                                  srif.go:14-49
    NewNoiseless / BatchNoise / NewAWGN      noise.go:23-159
    NewMonteCarloRuns, NewChiSquare          montecarlo.go:92-119, chisquare.go:16-95

Each constructor takes two extra keyword arguments, `n_filters` and `device`: a filter object is a
batch of `n_filters` independent filters advanced in lockstep on the GPU (1 = the drop-in case).
`Update` keeps the reference's one-step signature; `UpdateBatch` / `RunBatch` are the new batched
entry points that run many steps inside one kernel launch.  Go's `(value, error)` returns become
exceptions (`GkbError`); the reference's panics become exceptions too.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GkbError  # noqa: F401  (re-exported)


def _arr(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mat(x):
    a = _arr(x)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)
    return a


# --------------------------------------------------------------------------------------------------
# Noise (noise.go)
# --------------------------------------------------------------------------------------------------
class Noiseless:
    """noise.go:23-64: zero samples, carries Q and R."""

    def __init__(self, Q, R):
        if Q is None or R is None:
            raise ValueError("Q and R must be specified")  # noise.go:30-32 panics
        self.Q, self.R = _mat(Q), _mat(R)

    def Process(self, k):
        return np.zeros(self.Q.shape[0])

    def Measurement(self, k):
        return np.zeros(self.R.shape[0])

    def ProcessMatrix(self):
        return self.Q

    def MeasurementMatrix(self):
        return self.R

    def Reset(self):
        pass

    def __str__(self):
        return "Noiseless{\nQ=%s\nR=%s}\n" % (self.Q, self.R)


def NewNoiseless(Q, R):
    return Noiseless(Q, R)


class ReplayNoise(Noiseless):
    """Pre-baked samples indexed by step like noise.go BatchNoise (73-86), but carrying Q and R so
    that it can drive a real filter.  process: [steps, n] or [steps, n, n_filters]."""

    def __init__(self, Q, R, process, measurement):
        super().__init__(Q, R)
        self.process = None if process is None else _arr(process)
        self.measurement = None if measurement is None else _arr(measurement)

    def _get(self, arr, k, what):
        if arr is None:
            return None
        if k >= arr.shape[0]:
            raise IndexError("no %s noise defined at step k=%d" % (what, k))  # noise.go:74-76 panics
        return arr[k]

    def Process(self, k):
        v = self._get(self.process, k, "process")
        return np.zeros(self.Q.shape[0]) if v is None else v

    def Measurement(self, k):
        v = self._get(self.measurement, k, "measurement")
        return np.zeros(self.R.shape[0]) if v is None else v


class BatchNoise(ReplayNoise):
    """noise.go:67-106: stored vectors, Q = R = zero matrices."""

    def __init__(self, process, measurement):
        process, measurement = _arr(process), _arr(measurement)
        super().__init__(np.zeros((process.shape[1],) * 2), np.zeros((measurement.shape[1],) * 2), process, measurement)

    def __str__(self):
        return "BatchNoise"


class AWGN:
    """noise.go:109-159.  The reference seeds math/rand from the clock (irreproducible by design); here the samples
    come from the engine's counter-based Philox4x32-10 stream keyed by (seed, filter, step), generated on the device:
    inside the Monte Carlo kernel for the pure predictor NewMonteCarloRuns drives, and by `gkb_set_philox_noise` for
    any ordinary filter that carries this noise (vanilla_test.go:29-75, information_test.go, squareroot_test.go).
    Process(k) / Measurement(k) return, on the host, exactly the samples filter 0 of such a handle consumes at step
    k (`gkb_awgn_sample`); like the reference's AWGN a second Process(k) call at the same step is a fresh draw."""

    def __init__(self, Q, R, seed=None, device=0):
        self.Q, self.R = _mat(Q), _mat(R)
        for name, M in (("process", self.Q), ("measurement", self.R)):
            try:
                np.linalg.cholesky(M)
            except np.linalg.LinAlgError:
                raise ValueError("%s noise invalid" % name)  # noise.go:149-156 panics
        self._seed0, self._device = seed, device
        self.Reset()

    def Reset(self):
        # a new stream per Reset(), as the reference re-seeds (noise.go:145-159)
        if self._seed0 is None:
            self.seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
        else:
            self.seed = int(self._seed0)
        self._last_k, self._calls = None, 0

    def ProcessMatrix(self):
        return self.Q

    def MeasurementMatrix(self):
        return self.R

    def _sample(self, k, filter_index=0):
        n, m = self.Q.shape[0], self.R.shape[0]
        w, v, w2 = np.zeros(n), np.zeros(m), np.zeros(n)
        _lib.check(_lib.load().gkb_awgn_sample(n, m, _ptr(self.Q), _ptr(self.R), self.seed, filter_index, int(k), self._device,
                                               _ptr(w), _ptr(v), _ptr(w2)))
        return w, v, w2

    def Process(self, k, filter_index=0):
        """noise.go:127-131.  The first call at step k returns the draw of vanilla.go:146, a second consecutive call at
        the same k the (different) draw of vanilla.go:195."""
        self._calls = self._calls + 1 if self._last_k == k else 1
        self._last_k = k
        w, _, w2 = self._sample(k, filter_index)
        return w if self._calls % 2 == 1 else w2

    def Measurement(self, k, filter_index=0):
        """noise.go:133-137"""
        return self._sample(k, filter_index)[1]

    def __str__(self):
        return "AWGN{\nQ=%s\nR=%s}\n" % (self.Q, self.R)


def NewAWGN(Q, R, seed=None, device=0):
    return AWGN(Q, R, seed, device)


# --------------------------------------------------------------------------------------------------
# Estimate (kalman.go:64-72)
# --------------------------------------------------------------------------------------------------
class Estimate:
    """One Estimate, or a batch: arrays are [component] for a single filter / single step and gain
    leading `steps` and trailing `n_filters` axes in the batched calls."""

    def __init__(self, n, m, fields, status=None, squeeze=True):
        self._n, self._m, self._f, self._squeeze = n, m, fields, squeeze
        self.status = status

    def _get(self, name, shape):
        a = self._f.get(name)
        if a is None:
            return None
        steps, nf = a.shape[0], a.shape[-1]
        a = a.reshape((steps,) + shape + (nf,))
        if self._squeeze:
            if nf == 1:
                a = a[..., 0]
            if steps == 1:
                a = a[0]
        return a

    def State(self):
        return self._get("state", (self._n,))

    def Measurement(self):
        return self._get("meas", (self._m,))

    def Innovation(self):
        a = self._f.get("innov")
        if a is None:
            return None
        return self._get("innov", (a.shape[1],))

    def ObservationDev(self):
        return self._get("obs_dev", (self._m,))

    def Covariance(self):
        return self._get("covar", (self._n, self._n))

    def PredCovariance(self):
        return self._get("pred_covar", (self._n, self._n))

    def Gain(self):
        return self._get("gain", (self._n, self._m))

    def IsWithinNσ(self, N):
        """vanilla.go:231-239: |x_i| <= N sqrt(P_ii) for every component (single estimate only)."""
        x, P = np.asarray(self.State()), np.asarray(self.Covariance())
        if x.ndim != 1:
            raise ValueError("IsWithinNσ is defined for a single estimate")
        for i in range(x.shape[0]):
            b = N * np.sqrt(P[i, i])
            if x[i] > b or x[i] < -b:
                return False
        return True

    IsWithinNsigma = IsWithinNσ

    def IsWithin2σ(self):
        return self.IsWithinNσ(2)

    IsWithin2sigma = IsWithin2σ

    def __str__(self):
        return "{\ns=%s\ny=%s\nP=%s\nK=%s\nP-=%s\ni=%s\n}" % (self.State(), self.Measurement(), self.Covariance(),
                                                           self.Gain(), self.PredCovariance(), self.Innovation())


_OUT_FIELDS = ("state", "meas", "innov", "covar", "pred_covar", "gain", "obs_dev")


class _Filter:
    """Common handle plumbing."""

    _fm = False  # large-state handles use filter-major arrays at the C-ABI (gkb_filter_major)

    def __init__(self):
        self._h = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().gkb_destroy(h)
            except Exception:
                pass

    @property
    def n_filters(self):
        return self._nf

    def _alloc_out(self, steps, every_step, want, innov_len, buffers=None):
        """buffers: optional dict of caller-owned C-contiguous float64 arrays (e.g. pinned host memory) to receive the
        fields named by their keys (+ "status": int32 [n_filters]); anything missing is allocated here."""
        n, m, nf = self._n, self._m, self._nf
        rows = steps if every_step else 1
        sizes = {"state": n, "meas": m, "innov": innov_len, "covar": n * n, "pred_covar": n * n, "gain": n * m,
                 "obs_dev": m}
        shape = (lambda c: (rows, nf, c)) if self._fm else (lambda c: (rows, c, nf))
        buffers = buffers or {}

        def field(k):
            b = buffers.get(k)
            if b is None:
                return np.zeros(shape(sizes[k]))
            if b.dtype != np.float64 or not b.flags["C_CONTIGUOUS"] or b.size != int(np.prod(shape(sizes[k]))):
                raise GkbError(-10, "output buffer %r must be a C-contiguous float64 array of %d elements" % (k, int(np.prod(shape(sizes[k])))))
            return b.reshape(shape(sizes[k]))
        fields = {k: field(k) for k in _OUT_FIELDS if k in want and sizes[k] > 0}
        status = buffers.get("status")
        if status is None:
            status = np.zeros(nf, dtype=np.int32)
        out = _lib.Outputs()
        out.mem, out.every_step = _lib.HOST, int(every_step)
        for k, a in fields.items():
            setattr(out, k, a.ctypes.data)
        out.status = status.ctypes.data
        return out, fields, status

    def _raise_status(self, status):
        bad = status[status != 0]
        if bad.size and self._nf == 1:
            raise GkbError(int(bad[0]), "filter 0 failed during Update")

    def GetState(self):
        """raw (vector [n, N], matrix [n, n, N]) of the filters: x,P / i,I / x,S / b,R by kind"""
        if self._fm:
            vec, mat = np.zeros((self._nf, self._n)), np.zeros((self._nf, self._n * self._n))
            _lib.check(_lib.load().gkb_get_state(self._h, _ptr(vec), _ptr(mat)))
            return np.ascontiguousarray(vec.T), np.ascontiguousarray(mat.T).reshape(self._n, self._n, self._nf)
        vec, mat = np.zeros((self._n, self._nf)), np.zeros((self._n * self._n, self._nf))
        _lib.check(_lib.load().gkb_get_state(self._h, _ptr(vec), _ptr(mat)))
        return vec, mat.reshape(self._n, self._n, self._nf)


# --------------------------------------------------------------------------------------------------
# LDKF: Vanilla / Information / SquareRoot (kalman.go:35-48)
# --------------------------------------------------------------------------------------------------
class _LDKF(_Filter):
    _kind = None
    _want = ("state", "meas", "innov", "covar", "pred_covar", "gain")

    def _create(self, kind, x0, P0, F, G, H, noise, n_filters, device, from_state=False):
        lib = _lib.load()
        F, H = _arr(F), _mat(H)
        n, m = F.shape[0], H.shape[0]
        x0 = _arr(x0)
        P0 = _arr(P0)
        # checkMatDims (vanilla.go:23-31): x0 rows vs Covar0 cols, F rows vs Covar0 cols, H cols vs x0 rows
        if x0.shape[0] != P0.shape[1]:
            raise GkbError(-1, "x0(%dx...) Covar0(...x%d)" % (x0.shape[0], P0.shape[1]))
        if F.shape[0] != P0.shape[1]:
            raise GkbError(-1, "F(%dx...) Covar0(...x%d)" % (F.shape[0], P0.shape[1]))
        if H.shape[1] != x0.shape[0]:
            raise GkbError(-1, "H(...x%d) x0(%dx...)" % (H.shape[1], x0.shape[0]))
        G = None if G is None else _arr(G).reshape(n, -1)
        c = 0 if G is None else G.shape[1]
        Q, R = _mat(noise.ProcessMatrix()), _mat(noise.MeasurementMatrix())
        per_filter = 1 if x0.ndim == 2 else 0
        x0_abi = x0
        if per_filter and kind == _lib.VANILLA and n > 8:  # large-state handles are filter-major at the C-ABI
            x0_abi = np.ascontiguousarray(x0.T)
        h = C.c_void_p()
        if from_state:
            _lib.check(lib.gkb_create_information_from_state(n, m, c, n_filters, device, _ptr(x0), _ptr(P0), _ptr(F),
                                                             _ptr(G), _ptr(H), _ptr(Q), _ptr(R), C.byref(h)))
        else:
            _lib.check(lib.gkb_create_lti(kind, n, m, c, n_filters, device, _ptr(x0_abi), per_filter, _ptr(P0), _ptr(F),
                                          _ptr(G), _ptr(H), _ptr(Q), _ptr(R), C.byref(h)))
        self._fm = bool(lib.gkb_filter_major(h))
        self._h, self._n, self._m, self._c, self._nf, self._device = h, n, m, c, n_filters, device
        self.F, self.G, self.H = F, G, H
        self.needCtrl = not (G is None or not np.any(G))
        self.Noise = None
        self._apply_noise(noise, set_matrices=False)
        self._x0, self._P0, self._from_state = x0, P0, from_state

    # -- getters / setters (vanilla.go:80-118)
    def GetStateTransition(self):
        return self.F

    def GetInputControl(self):
        return self.G

    def GetMeasurementMatrix(self):
        return self.H

    def GetNoise(self):
        return self.Noise

    def SetStateTransition(self, F):
        self.F = _arr(F)
        _lib.check(_lib.load().gkb_set_state_transition(self._h, _ptr(self.F)))

    def SetInputControl(self, G):
        self.G = _arr(G).reshape(self._n, -1)
        self._c = self.G.shape[1]
        _lib.check(_lib.load().gkb_set_input_control(self._h, self._c, _ptr(self.G)))

    def SetMeasurementMatrix(self, H):
        self.H = _mat(H)
        self._m = self.H.shape[0]
        _lib.check(_lib.load().gkb_set_measurement_matrix(self._h, self._m, _ptr(self.H)))

    def _apply_noise(self, noise, set_matrices=True):
        lib = _lib.load()
        if isinstance(noise, AWGN) and self._fm:
            raise NotImplementedError("large-state handles carry Noiseless noise only")
        if set_matrices:
            Q, R = _mat(noise.ProcessMatrix()), _mat(noise.MeasurementMatrix())
            _lib.check(lib.gkb_set_noise(self._h, _ptr(Q), R.shape[0], _ptr(R)))
        if isinstance(noise, ReplayNoise):
            w, v = noise.process, noise.measurement
            steps = (w if w is not None else v).shape[0]

            def soa(a, comps):
                if a is None:
                    return None
                if a.ndim == 2:  # [steps, comps] shared by all filters -> broadcast
                    a = np.repeat(a[:, :, None], self._nf, axis=2)
                return np.ascontiguousarray(a.reshape(steps, comps, self._nf))
            w, v = soa(w, self._n), soa(v, self._m)
            _lib.check(lib.gkb_set_replay_noise(self._h, steps, _ptr(w), _ptr(v), _lib.HOST))
        if isinstance(noise, AWGN):  # samples drawn on the device, keyed by (seed, filter, step)
            _lib.check(lib.gkb_set_philox_noise(self._h, noise.seed, 0))
        self.Noise = noise

    def SetNoise(self, n):
        self._apply_noise(n)

    def Reset(self):
        _lib.check(_lib.load().gkb_reset(self._h))
        self.Noise.Reset()
        if isinstance(self.Noise, AWGN):  # Reset() re-seeds an AWGN (noise.go:145-159): re-arm the handle with the new stream
            _lib.check(_lib.load().gkb_set_philox_noise(self._h, self.Noise.seed, 0))

    def __str__(self):
        return "F=%s\nG=%s\nH=%s\n%s" % (self.F, self.G, self.H, self.Noise)

    # -- Update
    def _innov_len(self):
        return self._m

    def Update(self, measurement, control=None):
        """LDKF.Update(measurement, control) (kalman.go:36): one step of every filter in the batch.
        measurement: [m] (shared) or [m, n_filters]; control: [c]."""
        y = _arr(measurement)
        if y.shape[0] != self._m:  # checkMatDims(measurement, H, rows2rows) vanilla.go:133-135
            raise GkbError(-1, "measurement (y)(%dx...) H(%dx...)" % (y.shape[0], self._m))
        u = None if control is None else _arr(control)
        if self.needCtrl and (u is None or u.shape[0] != self._c):  # vanilla.go:129-131
            raise GkbError(-1, "control (u)(%sx...) G(...x%d)" % ("nil" if u is None else u.shape[0], self._c))
        if u is not None and u.shape[0] != self._c:
            u = None
        est = self.UpdateBatch(y[None], None if u is None else u[None], every_step=False)
        return est

    def UpdateBatch(self, measurements, controls=None, every_step=True, want=None):
        """Batched entry point: `steps` Updates inside one kernel launch.
        measurements: [steps, m] (shared by all filters) or [steps, m, n_filters]; controls [steps, c]."""
        lib = _lib.load()
        y = _arr(measurements)
        steps = y.shape[0]
        shared = 1 if y.ndim == 2 else 0
        if y.shape[1] != self._m:
            raise GkbError(-1, "measurement (y)(%dx...) H(%dx...)" % (y.shape[1], self._m))
        if not shared and y.shape[2] != self._nf:
            raise GkbError(-1, "measurements for %d filters given to a batch of %d" % (y.shape[2], self._nf))
        u = None if controls is None else _arr(controls).reshape(steps, -1)
        if self.needCtrl and (u is None or u.shape[1] != self._c):
            raise GkbError(-1, "control (u) G(...x%d)" % self._c)
        out, fields, status = self._alloc_out(steps, every_step, want or self._want, self._innov_len())
        if self._fm and not shared:
            y = np.ascontiguousarray(y.transpose(0, 2, 1))  # [steps, n_filters, m] at the C-ABI
        _lib.check(lib.gkb_update(self._h, steps, _ptr(y), shared, _ptr(u), _lib.HOST, C.byref(out)))
        self._raise_status(status)
        if self._fm:  # back to the [steps, component, n_filters] convention of Estimate
            fields = {k: np.ascontiguousarray(a.transpose(0, 2, 1)) for k, a in fields.items()}
        return Estimate(self._n, self._m, fields, status)


class Vanilla(_LDKF):
    _kind = _lib.VANILLA


class Information(_LDKF):
    _kind = _lib.INFORMATION
    _want = ("state", "meas", "innov", "covar", "pred_covar")

    def _innov_len(self):
        return self._n  # Innovation() returns the information state (information.go:272-274)

    def GetStateTransition(self):
        """*WARNING:* returns the INVERSE of F, like the reference (information.go:103-105)."""
        return np.linalg.inv(self.F)


class SquareRoot(_LDKF):
    _kind = _lib.SQRT


def NewVanilla(x0, Covar0, F, G, H, noise, n_filters=1, device=0):
    kf = Vanilla()
    kf._create(_lib.VANILLA, x0, Covar0, F, G, H, noise, n_filters, device)
    return kf, _est0(kf, x0, Covar0)


def NewPurePredictorVanilla(x0, Covar0, F, G, H, noise, n_filters=1, device=0):
    kf = Vanilla()
    kf._kind = _lib.PREDICTOR
    kf.predictionOnly = True
    kf._create(_lib.PREDICTOR, x0, Covar0, F, G, H, noise, n_filters, device)
    return kf, _est0(kf, x0, Covar0)


def NewInformation(i0, I0, F, G, H, noise, n_filters=1, device=0):
    kf = Information()
    kf._create(_lib.INFORMATION, i0, I0, F, G, H, noise, n_filters, device)
    return kf, None


def NewInformationFromState(x0, P0, F, G, H, noise, n_filters=1, device=0):
    kf = Information()
    kf._create(_lib.INFORMATION, x0, P0, F, G, H, noise, n_filters, device, from_state=True)
    return kf, None


def NewSquareRoot(x0, P0, F, G, H, noise, n_filters=1, device=0):
    kf = SquareRoot()
    kf._create(_lib.SQRT, x0, P0, F, G, H, noise, n_filters, device)
    return kf, _est0(kf, x0, P0)


def _est0(kf, x0, P0):
    """est0 = {x0, 0, 0, P0, 0, nil} (vanilla.go:34-37)"""
    n, m = kf._n, kf._m
    x0 = _arr(x0)
    if x0.ndim == 2:
        x0 = x0[:, 0]
    f = {"state": x0.reshape(1, n, 1), "meas": np.zeros((1, m, 1)), "innov": np.zeros((1, m, 1)),
         "covar": _arr(P0).reshape(1, n * n, 1), "pred_covar": np.zeros((1, n * n, 1))}
    return Estimate(n, m, f)


# --------------------------------------------------------------------------------------------------
# NLDKF: HybridKF / SRIF (kalman.go:51-60)
# --------------------------------------------------------------------------------------------------
class _NLDKF(_Filter):
    _want = ("state", "meas", "innov", "covar", "pred_covar", "gain", "obs_dev")

    def __init__(self):
        super().__init__()
        self._Phi = self._Ht = self._Gamma = None
        self.locked = True
        self.ekfMode = False
        self.sncEnabled = False

    def Prepare(self, Phi, Htilde):
        """hybrid.go:78-82 / srif.go:82-86"""
        self._Phi = _arr(Phi)
        self._Ht = None if Htilde is None else _mat(Htilde)
        self.locked = False

    def EKFEnabled(self):
        return self.ekfMode

    def _innov_len(self):
        return self._m

    def _one(self, has_meas, real, computed):
        if self.locked:
            raise GkbError(-4, "kf is locked (call Prepare() first)")  # hybrid.go:105-107
        flags = np.array([(_lib.F_MEAS if has_meas else 0) | (_lib.F_EKF if self.ekfMode else 0) |
                          (_lib.F_SNC if self.sncEnabled else 0)], dtype=np.uint8)
        if has_meas:
            real, computed = _arr(real), _arr(computed)
            if real.shape != computed.shape:  # checkMatDims(real, computed, rowsAndcols) hybrid.go:108-112
                raise GkbError(-1, "real observation %s computed observation %s" % (real.shape, computed.shape))
        est = self.RunBatch(flags, self._Phi[None], None if self._Ht is None else self._Ht[None],
                            None if real is None else real[None], None if computed is None else computed[None],
                            None if self._Gamma is None else self._Gamma[None], every_step=False)
        self.sncEnabled = False  # hybrid.go:140,201
        self.locked = True
        return est

    def Update(self, realObservation, computedObservation):
        return self._one(True, realObservation, computedObservation)

    def Predict(self):
        return self._one(False, None, None)

    def RunBatch(self, flags, Phi, Htilde, real_obs, computed_obs, Gamma=None, every_step=True, want=None, out_buffers=None):
        """Batched entry point: `steps` epochs of Prepare + Update/Predict in one kernel launch.
        Phi: [steps, n, n] (shared) or [steps, n*n, n_filters]; Htilde likewise; observations
        [steps, m] or [steps, m, n_filters]; flags uint8[steps]; Gamma [steps, n, q] (shared)."""
        lib = _lib.load()
        n, m, nf = self._n, self._m, self._nf
        flags = None if flags is None else np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
        Phi = _arr(Phi)
        steps = Phi.shape[0]
        # shared by all filters: [steps, n, n]; per filter: [steps, n, n, N] or [steps, n*n, N]
        phi_shared = 1 if (Phi.ndim == 3 and Phi.shape[1:] == (n, n)) else 0
        Phi = np.ascontiguousarray(Phi.reshape(steps, n * n) if phi_shared else Phi.reshape(steps, n * n, nf))
        h_shared = 1
        if Htilde is not None:
            Htilde = _arr(Htilde)
            h_shared = 1 if (Htilde.ndim == 3 and Htilde.shape[1:] == (m, n)) else 0
            Htilde = np.ascontiguousarray(Htilde.reshape(steps, m * n) if h_shared else Htilde.reshape(steps, m * n, nf))

        def obs(a):
            if a is None:
                return None
            a = _arr(a)
            if a.ndim == 2:
                a = np.repeat(a[:, :, None], nf, axis=2)
            return np.ascontiguousarray(a.reshape(steps, m, nf))
        real_obs, computed_obs = obs(real_obs), obs(computed_obs)
        if Gamma is not None:
            Gamma = np.ascontiguousarray(_arr(Gamma).reshape(steps, -1))
        out, fields, status = self._alloc_out(steps, every_step, want or self._want, self._innov_len(), out_buffers)
        _lib.check(lib.gkb_nl_run(self._h, steps, _ptr(flags), _ptr(Phi), phi_shared, _ptr(Htilde), h_shared,
                                  _ptr(real_obs), _ptr(computed_obs), _ptr(Gamma), _lib.HOST, C.byref(out)))
        self._raise_status(status)
        est = Estimate(n, m, fields, status)
        # what the reference's estimates carry for smoothing (hybrid.go:193-196: copies of Phi and Gamma)
        est._Phi, est._phi_shared, est._every_step = Phi, phi_shared, bool(every_step)
        est._snc = bool(flags is not None and np.any(flags & _lib.F_SNC) and Gamma is not None)
        self._steps_run = getattr(self, "_steps_run", 0) + steps
        return est

    def SmoothAll(self, estimates):
        """hybrid.go:209-238 / srif.go:165-192: smooths, in place, the estimates of every epoch run so far.
        `estimates` is the batched Estimate of ONE RunBatch(every_step=True) call covering all the filter's
        epochs, or the list of single-epoch estimates returned by Update()/Predict()."""
        lib = _lib.load()
        n, nf = self._n, self._nf
        if isinstance(estimates, (list, tuple)):
            parts = list(estimates)
            count = len(parts)
        else:
            parts = None
            count = estimates._f["state"].shape[0] if estimates._every_step else 1
        expected = _lib.load().gkb_step(self._h)
        if count != expected:  # hybrid.go:210-212
            raise GkbError(-10, "incorrect number of estimates provided: %d instead of expected %d" % (count, expected))
        if parts is not None:
            if any(e._snc for e in parts[1:]):
                raise NotImplementedError("not yet implemented")  # hybrid.go:234 panics
            shared = all(e._phi_shared for e in parts)
            Phi = np.ascontiguousarray(np.concatenate(
                [e._Phi if e._phi_shared == shared else np.repeat(e._Phi[:, :, None], nf, axis=2) for e in parts]))
            xs = np.ascontiguousarray(np.concatenate([e._f["state"] for e in parts]))
            Ps = np.ascontiguousarray(np.concatenate([e._f["covar"] for e in parts]))
        else:
            if estimates._snc:
                raise NotImplementedError("not yet implemented")
            Phi, shared = estimates._Phi, estimates._phi_shared
            xs, Ps = estimates._f["state"], estimates._f["covar"]
        status = np.zeros(nf, dtype=np.int32)
        _lib.check(lib.gkb_smooth_all(n, count, nf, self._device, _ptr(Phi), int(shared), _ptr(xs), _ptr(Ps), _lib.HOST,
                                      status.ctypes.data))
        if np.any(status != 0):
            raise GkbError(int(status[status != 0][0]), "provided STM is not invertible")
        if parts is not None:
            for k, e in enumerate(parts):
                e._f["state"][...] = xs[k:k + 1]
                e._f["covar"][...] = Ps[k:k + 1]
        return None


class HybridKF(_NLDKF):
    def EnableEKF(self):
        self.ekfMode = True

    def DisableEKF(self):
        self.ekfMode = False

    def PreparePNT(self, Gamma):
        """hybrid.go:86-89: enables the SNC for the next update only."""
        self._Gamma = _arr(Gamma)
        self.sncEnabled = True

    def SetStrict(self, on=True):
        """Reference-order arithmetic (gkb_set_strict): dense products in the written order, no FMA contraction,
        dense Joseph form -- bit-identical to the reference's formulas; the mode to use whenever the results have to be
        the reference's (ill-conditioned OD runs amplify the fast kernels' rounding differences to percent level).
        Default: on for a single-filter HybridKF (the reference-shaped use), off for batched handles."""
        _lib.check(_lib.load().gkb_set_strict(self._h, int(bool(on))))

    def RunOD(self, scenario, orbit0, sigma_range, sigma_rate, seed, flags=None, every_step=False, filter_offset=0,
              out_buffers=None):
        """The fused OD run (gkb_od_run): per epoch the reference orbit / STM / range + range-rate partials /
        observations of every filter are computed on the device (gokalman_b200.od.Scenario holds the per-epoch
        station and truth tables) and consumed by Prepare + Update / Predict in the same kernel.  orbit0:
        [6, n_filters] initial reference orbits, or None to continue from the orbits the handle holds.
        Returns the batched Estimate (State / Covariance, final or every epoch)."""
        steps = scenario.steps
        flags = scenario.flags if flags is None else np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
        cfg = scenario.config(orbit0, sigma_range, sigma_rate, seed, filter_offset)
        out, fields, status = self._alloc_out(steps, every_step, ("state", "covar"), self._m, out_buffers)
        _lib.check(_lib.load().gkb_od_run(self._h, C.byref(cfg), steps, flags.ctypes.data, C.byref(out)))
        self._raise_status(status)
        return Estimate(self._n, self._m, fields, status)

    def SetNoise(self, n):
        Q, R = _mat(n.ProcessMatrix()), _mat(n.MeasurementMatrix())
        _lib.check(_lib.load().gkb_set_noise(self._h, _ptr(Q), R.shape[0], _ptr(R)))
        self.Noise = n

    def GetNoise(self):
        return self.Noise

    def __str__(self):
        return "HybridKF [k=%d]\n%s" % (_lib.load().gkb_step(self._h), self.Noise)


class SRIF(_NLDKF):
    _want = ("state", "meas", "innov", "covar", "pred_covar", "obs_dev")

    def _innov_len(self):
        return self._n  # Innovation() returns b (srif.go:238-240)

    def EnableEKF(self):
        pass  # srif.go:66-72: no-ops

    def DisableEKF(self):
        pass

    def EKFEnabled(self):
        return False

    def PreparePNT(self, Gamma):
        pass  # srif.go:75

    def SetNoise(self, n):
        raise NotImplementedError("noise not yet supported for SRIF")  # srif.go:77-79 panics


def NewHybridKF(x0, P0, noise, measSize, n_filters=1, device=0):
    """hybrid.go:23-34.  n_filters = 1: the drop-in filter, reference-order arithmetic by default (its estimates equal
    the reference's bit for bit); n_filters > 1: a batch in lockstep, production (FMA) kernels by default -- SetStrict()."""
    lib = _lib.load()
    x0, P0 = _arr(x0), _arr(P0)
    n = x0.shape[0]
    if n != P0.shape[1]:  # hybrid.go:25-27
        raise GkbError(-1, "x0(%dx...) Covar0(...x%d)" % (n, P0.shape[1]))
    Q, R = noise.ProcessMatrix(), _mat(noise.MeasurementMatrix())
    Q = None if Q is None else _mat(Q)
    q = 0 if Q is None else Q.shape[0]
    if q > 3:  # only a q x q SNC matrix (q <= 3) is used as Gamma Q Gamma^T
        raise GkbError(-8, "process noise of size %d: the SNC path takes q <= 3" % q)
    kf = HybridKF()
    h = C.c_void_p()
    per_filter = 1 if x0.ndim == 2 else 0
    _lib.check(lib.gkb_create_hybrid(n, measSize, q, n_filters, device, _ptr(x0), per_filter, _ptr(P0), _ptr(Q), _ptr(R),
                                     C.byref(h)))
    kf._h, kf._n, kf._m, kf._nf, kf._device, kf.Noise = h, n, measSize, n_filters, device, noise
    f = {"state": (x0[:, 0] if x0.ndim == 2 else x0).reshape(1, n, 1), "meas": np.zeros((1, measSize, 1)),
         "innov": np.zeros((1, measSize, 1)), "obs_dev": np.zeros((1, measSize, 1)),
         "covar": P0.reshape(1, n * n, 1), "pred_covar": np.zeros((1, n * n, 1))}
    return kf, Estimate(n, measSize, f)


def NewSRIF(x0, P0, measSize, nonTriR, noise, n_filters=1, device=0):
    lib = _lib.load()
    x0, P0 = _arr(x0), _arr(P0)
    n = x0.shape[0]
    if n != P0.shape[1]:  # srif.go:16-18
        raise GkbError(-1, "x0(%dx...) P0(...x%d)" % (n, P0.shape[1]))
    R = _mat(noise.MeasurementMatrix())
    kf = SRIF()
    h = C.c_void_p()
    _lib.check(lib.gkb_create_srif(n, measSize, n_filters, device, _ptr(x0), 0, _ptr(P0), _ptr(R), int(bool(nonTriR)),
                                   C.byref(h)))
    kf._h, kf._n, kf._m, kf._nf, kf._device, kf.Noise = h, n, measSize, n_filters, device, noise
    # est0: R0 = chol(diag(1/P0_ii)), b0 = R0 x0; Covariance() = inv(R0) inv(R0)^T (srif.go:20-47)
    d = np.diag(P0).copy()
    f = {"state": x0.reshape(1, n, 1), "covar": np.diag(d).reshape(1, n * n, 1),
         "pred_covar": np.diag(d).reshape(1, n * n, 1), "innov": (x0 / np.sqrt(d)).reshape(1, n, 1)}
    return kf, Estimate(n, measSize, f)


# --------------------------------------------------------------------------------------------------
# Monte Carlo + chi-square (montecarlo.go, chisquare.go)
# --------------------------------------------------------------------------------------------------
class MonteCarloRuns:
    """montecarlo.go:12-16.  The reference stores every Estimate of every run; here the runs are a
    recipe (model, noise seed, controls) that the fused kernel regenerates on demand -- Philox is
    counter-based, so NewChiSquare sees exactly the trajectories Mean/StdDev/Truth describe."""

    def __init__(self, samples, steps, rowsH, controls, kf, trial_offset=0, devices=None, reduce="nccl"):
        self.runs, self.steps, self.rowsH = samples, steps, rowsH
        self.kf, self.controls, self.trial_offset = kf, controls, trial_offset
        # devices: the runs are sharded over these GPUs of this process (gkb_mc_chisquare_multi: contiguous trial
        # ranges, one collective -- NCCL all-reduce or the rank-ordered peer-memory sum -- inside the C-ABI)
        self.devices = None if devices is None else [int(d) for d in devices]
        self.reduce = {"nccl": _lib.REDUCE_NCCL, "peer": _lib.REDUCE_PEER}[reduce]
        self.noise = kf.Noise
        self._stats = None

    def _config(self, kind, tested, with_nees, with_nis, trials=None):
        """gkb_mc_config: the truth generator is the pure predictor `self.kf` with ITS noise (montecarlo.go:92);
        the tested filter -- any LDKF, chisquare.go:16 -- brings its OWN F / G / H and its own Noise's Q / R."""
        kf = self.kf
        cfg = _lib.McConfig()
        self._keep = keep = {}
        keep["F"], keep["H"] = _arr(kf.F), _mat(kf.H)
        keep["G"] = None if kf.G is None else _arr(kf.G)
        keep["Q"], keep["R"] = _mat(self.noise.ProcessMatrix()), _mat(self.noise.MeasurementMatrix())
        keep["x0t"] = _arr(kf._x0)
        keep["x0f"] = _arr(tested._x0) if tested is not None else _arr(kf._x0)
        keep["P0"] = _arr(tested._P0) if tested is not None else _arr(kf._P0)
        keep["u"] = None if self.controls is None else _arr(self.controls).reshape(self.steps, -1)
        cfg.kind, cfg.n, cfg.m, cfg.c = kind, kf._n, kf._m, kf._c
        if tested is not None:
            if (tested._n, tested._m) != (kf._n, kf._m):  # the reference's mat64 products would panic
                raise GkbError(-1, "tested filter is %dx%d, the Monte Carlo runs are %dx%d" % (tested._n, tested._m, kf._n, kf._m))
            if tested.G is not None and kf._c and tested._c != kf._c:
                raise GkbError(-1, "control (u)(%dx...) G(...x%d)" % (kf._c, tested._c))
            tn = tested.Noise
            keep["fF"], keep["fH"] = _arr(tested.F), _mat(tested.H)
            # a tested filter without G ignores the controls (needCtrl false): an all-zero G says exactly that
            keep["fG"] = (None if kf._c == 0 else np.zeros((kf._n, kf._c))) if tested.G is None else _arr(tested.G)
            keep["fQ"], keep["fR"] = _mat(tn.ProcessMatrix()), _mat(tn.MeasurementMatrix())
            for name, key in (("filter_F", "fF"), ("filter_G", "fG"), ("filter_H", "fH"), ("filter_Q", "fQ"), ("filter_R", "fR")):
                setattr(cfg, name, None if keep[key] is None else keep[key].ctypes.data)
        for name, key in (("F", "F"), ("G", "G"), ("H", "H"), ("Q", "Q"), ("R", "R"), ("x0_truth", "x0t"),
                          ("x0_filter", "x0f"), ("P0", "P0"), ("controls", "u")):
            setattr(cfg, name, None if keep[key] is None else keep[key].ctypes.data)
        cfg.trials = self.runs if trials is None else trials
        cfg.trial_offset, cfg.steps = self.trial_offset, self.steps
        cfg.with_nees, cfg.with_nis = int(with_nees), int(with_nis)
        cfg.info_raw_init = int(kind == _lib.INFORMATION and tested is not None and not tested._from_state)
        cfg.device = kf._device
        if isinstance(self.noise, ReplayNoise):
            cfg.noise_mode = _lib.NOISE_REPLAY
            w = _arr(self.noise.process).reshape(self.steps, kf._n, self.runs)
            v = _arr(self.noise.measurement).reshape(self.steps, kf._m, self.runs)
            keep["w"], keep["v"] = w, v
            cfg.w, cfg.v, cfg.noise_mem = w.ctypes.data, v.ctypes.data, _lib.HOST
        else:
            cfg.noise_mode, cfg.seed = _lib.NOISE_PHILOX, self.noise.seed
        return cfg

    def _run(self, cfg, want_stats=False, want_truth=False, want_noise=False, want_status=False, sums=False):
        n, m, steps, runs = cfg.n, cfg.m, cfg.steps, cfg.trials
        out = _lib.McOutputs()
        out.mem, out.sums_only = _lib.HOST, int(sums)
        res = {"NIS": np.zeros(steps), "NEES": np.zeros(steps)}
        out.nis, out.nees = res["NIS"].ctypes.data, res["NEES"].ctypes.data
        if want_stats:
            for key in ("sum_d", "sum_dd", "x_ref"):
                res[key] = np.zeros((steps, n))
                setattr(out, key, res[key].ctypes.data)
        if want_truth:
            res["truth_x"], res["truth_y"] = np.zeros((steps, n, runs)), np.zeros((steps, m, runs))
            out.truth_x, out.truth_y = res["truth_x"].ctypes.data, res["truth_y"].ctypes.data
        if want_noise:
            res["noise_w"], res["noise_v"] = np.zeros((steps, n, runs)), np.zeros((steps, m, runs))
            out.noise_w, out.noise_v = res["noise_w"].ctypes.data, res["noise_v"].ctypes.data
        res["first_error"] = np.zeros(1, dtype=np.int32)
        out.first_error = res["first_error"].ctypes.data
        if want_status:
            res["status"] = np.zeros(runs, dtype=np.int32)
            out.status = res["status"].ctypes.data
        if self.devices is not None:
            devs = (C.c_int * len(self.devices))(*self.devices)
            _lib.check(_lib.load().gkb_mc_chisquare_multi(C.byref(cfg), devs, len(self.devices), self.reduce, C.byref(out)))
        else:
            _lib.check(_lib.load().gkb_mc_chisquare(C.byref(cfg), C.byref(out)))
        return res

    def _get_stats(self):
        if self._stats is None:
            cfg = self._config(_lib.VANILLA, None, True, True)
            self._stats = self._run(cfg, want_stats=True)
        return self._stats

    def Mean(self, step):
        """montecarlo.go:18-37: mean of every state component over the runs at `step`."""
        s = self._get_stats()
        return s["x_ref"][step] + s["sum_d"][step] / self.runs

    def StdDev(self, step):
        """montecarlo.go:40-59: unbiased standard deviation over the runs at `step`."""
        s = self._get_stats()
        var = (s["sum_dd"][step] - s["sum_d"][step] ** 2 / self.runs) / (self.runs - 1)
        return np.sqrt(np.maximum(var, 0.0))

    def Truth(self, with_noise=False):
        """The (state, measurement) trajectories of every run: [steps, n, runs], [steps, m, runs]."""
        cfg = self._config(_lib.VANILLA, None, True, True)
        r = self._run(cfg, want_truth=True, want_noise=with_noise)
        if with_noise:
            return r["truth_x"], r["truth_y"], r["noise_w"], r["noise_v"]
        return r["truth_x"], r["truth_y"]

    def AsCSV(self, headers):
        """montecarlo.go:62-89"""
        tx, _ = self.Truth()
        n = tx.shape[1]
        out = []
        for i in range(n):
            header = headers[i]
            lines = ["".join("%s-%d," % (header, r) for r in range(self.runs)) + header + "-mean," + header + "-stddev"]
            for k in range(self.steps):
                mean, std = self.Mean(k), self.StdDev(k)
                lines.append("".join("%f," % tx[k, i, r] for r in range(self.runs)) + "%f,%f" % (mean[i], std[i]))
            out.append("\n".join(lines))
        return out


def _controls(controls, steps):
    """montecarlo.go:97-107 / chisquare.go:26-35: exactly `steps` vectors, or one vector => zeros."""
    if controls is None:
        return None
    controls = [np.asarray(c, dtype=np.float64).reshape(-1) for c in controls]
    if len(controls) == 1:
        return np.zeros((steps, controls[0].shape[0]))
    if len(controls) != steps:
        raise ValueError("must provide as much control vectors as steps, or just one control vector")
    return np.stack(controls)


def NewMonteCarloRuns(samples, steps, rowsH, controls, kf, trial_offset=0, devices=None, reduce="nccl"):
    """montecarlo.go:92-119.  `kf` must be a pure predictor whose noise is AWGN (or ReplayNoise).  devices: shard the
    runs over several GPUs of this process (Mean / StdDev / NewChiSquare then reduce across them inside the C-ABI)."""
    if not getattr(kf, "predictionOnly", False):
        raise ValueError("the Kalman filter needed for the Monte Carlo runs must be a pure predictor")
    return MonteCarloRuns(samples, steps, rowsH, _controls(controls, steps), kf, trial_offset, devices, reduce)


_KIND_OF = {Vanilla: _lib.VANILLA, Information: _lib.INFORMATION, SquareRoot: _lib.SQRT}


def NewChiSquare(kf, runs, controls, withNEES, withNIS, sums=False):
    """chisquare.go:16-95.  Returns (NISmeans, NEESmeans) -- in that order, like the reference.
    sums=True returns the per-step SUMS over this object's runs instead (for a multi-GPU all-reduce,
    see gokalman_b200.sharding)."""
    if not withNEES and not withNIS:
        raise GkbError(-10, "Chi Square requires either NEES or NIS or both")
    ctrl = _controls(controls, runs.steps)  # raises like chisquare.go:33-35 returns an error
    if ctrl is not None and kf.needCtrl and ctrl.shape[1] != kf._c:
        raise GkbError(-1, "control (u)(%dx...) G(...x%d)" % (ctrl.shape[1], kf._c))  # Update error -> panic, chisquare.go:40-42
    kind = _lib.VANILLA if isinstance(kf, Vanilla) else _KIND_OF[type(kf)]
    saved = runs.controls
    runs.controls = ctrl if ctrl is not None else saved
    try:
        cfg = runs._config(kind, kf, withNEES, withNIS)
        res = runs._run(cfg, sums=sums)
    finally:
        runs.controls = saved
    if res["first_error"][0] != 0:  # the reference panics (chisquare.go:40-42)
        raise GkbError(int(res["first_error"][0]), "a trial failed during Update")
    return res["NIS"], res["NEES"]


# --------------------------------------------------------------------------------------------------
# BatchKF (batch.go) and BatchGroundTruth (truth.go)
# --------------------------------------------------------------------------------------------------
class BatchKF:
    """batch.go:12-79.  SetNextMeasurement stores the measurements on the host; Solve() sends them to the
    device in one call (`gkb_batch_solve`: Lambda / N accumulation + inverse in one kernel).  SolveBatch is
    the batched entry point: N independent batch filters, each with its own measurement streams."""

    def __init__(self, numMeasurements, noise, device=0):
        self._cap, self.noise, self._device = int(numMeasurements), noise, device
        self.Measurements = []
        self.step = 0

    def SetNextMeasurement(self, realObs, computedObs, Phi, H):
        if self.step >= self._cap:
            raise IndexError("index out of range: BatchKF was created for %d measurements" % self._cap)  # Go slice panic
        real, comp, H = _arr(realObs), _arr(computedObs), _mat(H)
        self.Measurements.append(dict(RealObs=real, ComputedObs=comp, ObservationDev=real - comp, Phi=Phi, H=H))
        self.step += 1

    def Solve(self):
        """batch.go:64-79 -> (xHat0, P0); raises where the reference returns an error."""
        if not self.Measurements:
            raise GkbError(-10, "no measurement was set")
        H = np.stack([mm["H"] for mm in self.Measurements])
        real = np.stack([mm["RealObs"] for mm in self.Measurements])
        comp = np.stack([mm["ComputedObs"] for mm in self.Measurements])
        x, P, status = self.SolveBatch(H, real, comp, n_filters=1)
        if status[0] != 0:
            raise GkbError(int(status[0]), "could not invert the information matrix")
        return x[:, 0], P[:, :, 0]

    def SolveBatch(self, H, real_obs, computed_obs, n_filters):
        """H: [steps, m, n] (shared) or [steps, m, n, N]; observations [steps, m] or [steps, m, N].
        Returns xHat0 [n, N], P0 [n, n, N], status [N]."""
        lib = _lib.load()
        H = _arr(H)
        steps, m, n = H.shape[0], H.shape[1], H.shape[2]
        nf = int(n_filters)
        h_shared = 1 if H.ndim == 3 else 0
        H = np.ascontiguousarray(H.reshape(steps, m * n) if h_shared else H.reshape(steps, m * n, nf))

        def obs(a):
            a = _arr(a)
            if a.ndim == 2:
                a = np.repeat(a[:, :, None], nf, axis=2)
            return np.ascontiguousarray(a.reshape(steps, m, nf))
        real_obs, computed_obs = obs(real_obs), obs(computed_obs)
        R = _mat(self.noise.MeasurementMatrix())
        if R.shape[0] != m:
            raise GkbError(-1, "H(%dx...) R(%dx...)" % (m, R.shape[0]))
        x, P, status = np.zeros((n, nf)), np.zeros((n * n, nf)), np.zeros(nf, dtype=np.int32)
        _lib.check(lib.gkb_batch_solve(n, m, steps, nf, self._device, _ptr(R), _ptr(H), h_shared, _ptr(real_obs),
                                       _ptr(computed_obs), _lib.HOST, _ptr(x), _ptr(P), status.ctypes.data))
        return x, P.reshape(n, n, nf), status


def NewBatchKF(numMeasurements, noise, device=0):
    return BatchKF(numMeasurements, noise, device)


class ErrorEstimate(Estimate):
    """truth.go:68-70: a VanillaEstimate holding (estimate - truth)."""


class BatchGroundTruth:
    """truth.go:10-66 (host side, like exporter.go: it only subtracts stored vectors)."""

    def __init__(self, states, measurements):
        self.states, self.measurements = states, measurements

    def Error(self, k, est):
        return self.ErrorWithOffset(k, est, None)

    def ErrorWithOffset(self, k, est, offset):
        x = np.asarray(est.State(), dtype=np.float64)
        y = np.asarray(est.Measurement(), dtype=np.float64)
        es, em = np.zeros_like(x), np.zeros_like(y)
        if k >= 0:
            es = x.copy()
            if offset is not None:
                es = es + _arr(offset)
            if self.states is not None:
                t = self.states[k]
                if t is not None:
                    t = _arr(t)
                    if t.shape[0] != es.shape[0]:
                        raise ValueError("ground truth state size different from estimated state size (k=%d: %d != %d)"
                                         % (k, es.shape[0], t.shape[0]))
                    es = es - t
            em = y.copy()
            if self.states is not None:  # truth.go:52 tests t.states here too
                t = self.measurements[k]
                if t is not None:
                    t = _arr(t)
                    if t.shape[0] != em.shape[0]:
                        raise ValueError("ground truth measurement size different from estimated measurement size (k=%d)" % k)
                    em = em - t
        n, m = es.shape[0], em.shape[0]
        f = {"state": es.reshape(1, n, 1), "meas": em.reshape(1, m, 1),
             "covar": np.asarray(est.Covariance(), dtype=np.float64).reshape(1, n * n, 1)}
        return ErrorEstimate(n, m, f)


def NewBatchGroundTruth(states, measurements):
    return BatchGroundTruth(states, measurements)


# --------------------------------------------------------------------------------------------------
# c2d.go -- stays host-side (BASELINE.json north_star: "c2d.go and exporter.go stay host-side")
# --------------------------------------------------------------------------------------------------
def VanLoan(A, Gamma, W, dt):
    """c2d.go:13-75: (F, Q) of the discretised system from the continuous-time (A, Gamma, W) by Van Loan's
    matrix exponential.  Returns (F, Q, err) like the Go function: `err` is None or the Nyquist warning
    string (the reference still returns F and Q with it).  Quirk kept: the eigenvalue used for the test is
    the LAST one of the spectrum (the loop never updates its running maximum, c2d.go:19-24)."""
    from scipy.linalg import expm
    A, Gamma, W = _mat(A), _mat(Gamma), _mat(W)
    n = A.shape[0]
    lam = np.linalg.eigvals(A)
    err = None
    if 2.0 * abs(lam[-1]) * dt >= np.pi:
        err = "gokalman: Nyquist sampling criterion not fulfilled with Δt=%f" % dt
    GWG = (Gamma @ W) @ Gamma.T * dt
    Ap = A * dt
    M = np.zeros((2 * n, 2 * n))
    M[:n, :n] = -Ap
    M[n:, n:] = Ap.T
    M[:n, n:] = GWG
    E = expm(M)
    F = np.ascontiguousarray(E[n:, n:].T)
    Q = F @ E[:n, n:]
    Q = np.triu(Q) + np.triu(Q, 1).T  # AsSymDense: the upper triangle is kept
    return F, Q, err


def VanLoanBatch(A, Gamma, W, dt, device=0):
    """c2d.go:13-75 for a batch of systems on the device (`gkb_van_loan`): A [n, n] (shared) or [n, n, N]; Gamma [n, q]
    or [n, q, N]; W [q, q]; dt a scalar or [N].  Returns F [n, n, N], Q [n, n, N], status [N].  The Nyquist warning is
    the single-system host wrapper's (VanLoan above); the batch call never fails on it, like the reference still
    returns F and Q."""
    lib = _lib.load()
    A, Gamma, W = _arr(A), _arr(Gamma), _mat(W)
    n, q = A.shape[0], W.shape[0]
    dt = np.ascontiguousarray(np.atleast_1d(np.asarray(dt, dtype=np.float64)))
    a_shared, g_shared, dt_shared = int(A.ndim == 2), int(Gamma.reshape(n, q, -1).shape[2] == 1), int(dt.size == 1)
    count = max(1 if a_shared else A.shape[2], 1 if g_shared else Gamma.reshape(n, q, -1).shape[2], dt.size)
    A = np.ascontiguousarray(A.reshape(n * n) if a_shared else A.reshape(n * n, count))
    Gamma = np.ascontiguousarray(Gamma.reshape(n * q) if g_shared else Gamma.reshape(n * q, count))
    F, Q, status = np.zeros((n * n, count)), np.zeros((n * n, count)), np.zeros(count, dtype=np.int32)
    _lib.check(lib.gkb_van_loan(n, q, count, device, _ptr(A), a_shared, _ptr(Gamma), g_shared, _ptr(W), _ptr(dt), dt_shared,
                                _lib.HOST, _ptr(F), _ptr(Q), status.ctypes.data))
    return F.reshape(n, n, count), Q.reshape(n, n, count), status


# --------------------------------------------------------------------------------------------------
# helper.go: exported helpers
# --------------------------------------------------------------------------------------------------
def ScaledIdentity(n, s):
    """helper.go:13-23"""
    return s * np.eye(n)


def DenseIdentity(n):
    """helper.go:26-28"""
    return np.eye(n)


def ScaledDenseIdentity(n, s):
    """helper.go:31-41"""
    return s * np.eye(n)


def Identity(n):
    """helper.go:44-46"""
    return np.eye(n)


def IsNil(m):
    """helper.go:49-63: nil or all zeros"""
    return m is None or not np.any(np.asarray(m))


def AsSymDense(m):
    """helper.go:65-84: the upper triangle as a symmetric matrix; error when an off-diagonal pair differs by
    more than 1e-6 absolute AND more than 1e-2 relative."""
    m = _mat(m)
    r, c = m.shape
    if r != c:
        raise GkbError(-3, "matrix is not square")
    for i in range(r):
        for j in range(i):
            a, b = m[i, j], m[j, i]
            if abs(a - b) > 1e-6 and abs(a - b) > 1e-2 * max(abs(a), abs(b)):
                raise GkbError(-3, "matrix is not symmetric (%d, %d): %.30f != %.30f" % (i, j, a, b))
    return np.triu(m) + np.triu(m, 1).T


def Sign(v):
    """helper.go:133-138"""
    return 1.0 if abs(v) < 1e-12 or v > 0 else -1.0


def HouseholderTransf(A, n, m, device=0):
    """helper.go:142-172, in place like the Go function (and returns A).  A: [(n+m), (n+1)] or a batch
    [(n+m), (n+1), count]; runs on the GPU (`gkb_householder_transf`)."""
    lib = _lib.load()
    A = np.asarray(A, dtype=np.float64)
    if A.shape[0] != n + m or A.shape[1] != n + 1:
        raise GkbError(-1, "A(%dx%d) is not (n+m)x(n+1) = %dx%d" % (A.shape[0], A.shape[1], n + m, n + 1))
    count = 1 if A.ndim == 2 else A.shape[2]
    buf = np.ascontiguousarray(A.reshape((n + m) * (n + 1), count))
    _lib.check(lib.gkb_householder_transf(n, m, count, device, _ptr(buf), _lib.HOST))
    A[...] = buf.reshape(A.shape)
    return A
