"""GPU parity, NLDKF kinds (HybridKF CKF/EKF/SNC and SRIF) against the CPU oracle, 1e-10."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


def _od_streams(rng, n, m, nf, steps):
    """Synthetic per-filter STM-like Phi (near identity, well conditioned) and measurement partials."""
    Phi = np.eye(n)[None, :, :, None] + 0.02 * rng.standard_normal((steps, n, n, nf))
    Ht = rng.standard_normal((steps, m, n, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.05 * rng.standard_normal((steps, m, nf))
    return Phi, Ht, real, comp


def _oracle_run(o, flags, Phi, Ht, real, comp, Gamma, f, F_MEAS, F_EKF, F_SNC):
    ests = []
    for k in range(len(flags)):
        o.Prepare(Phi[k, :, :, f], Ht[k, :, :, f])
        if flags[k] & F_EKF:
            o.EnableEKF()
        else:
            o.DisableEKF()
        if flags[k] & F_SNC:
            o.PreparePNT(Gamma[k])
        ests.append(o.UpdateNL(real[k, :, f], comp[k, :, f]) if flags[k] & F_MEAS else o.Predict())
    return ests


def _check(est, refs, fields, tag):
    getters = {"state": "State", "meas": "Measurement", "innov": "Innovation", "covar": "Covariance",
               "pred_covar": "PredCovariance", "gain": "Gain", "obs_dev": "ObservationDev"}
    nf, steps = len(refs), len(refs[0])
    for fld in fields:
        g = getattr(est, getters[fld])()
        for f in range(nf):
            rows = []
            for k in range(steps):
                a = np.asarray(getattr(refs[f][k], getters[fld])())
                rows.append(a)
            width = max(r.size for r in rows)
            ref = np.stack([np.pad(r.reshape(-1), (0, width - r.size)) for r in rows])
            got = (g[..., f] if nf > 1 else g).reshape(steps, -1)[:, :width]
            err = fx.scaled_err_steps(got, ref)  # every step scaled by its own max-abs
            assert err <= TOL, (tag, fld, f, err)


@pytest.mark.parametrize("n,m,q", [(6, 2, 3), (4, 2, 2), (3, 1, 0), (6, 3, 3), (8, 3, 3), (7, 2, 2), (8, 1, 0), (8, 2, 3), (7, 3, 0), (7, 1, 1)])
def test_hybrid_ckf_ekf_snc_matches_oracle(oracle, n, m, q):
    """hybrid.go:104-204: Predict / CKF update / EKF update / SNC epochs mixed in one batched run
    with per-filter Phi, Htilde and observations (BASELINE config 4 shape is n=6, m=2, q=3)."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(42 + n + m)
    nf, steps = 33, 48
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 10.0), np.full(n // 2, 1.0)]))
    Q = np.diag(np.full(q, 1e-3)) if q else None
    R = np.diag(np.full(m, 1e-2))
    flags = np.zeros(steps, dtype=np.uint8)
    for k in range(steps):
        fl = F_MEAS if (k % 5 != 3) else 0        # every 5th epoch has no measurement -> Predict()
        if k >= 15:
            fl |= F_EKF                            # EKF after 15 epochs (hybrid_test.go:65,270-273)
        if q and (fl & F_MEAS) and k % 2 == 0:
            fl |= F_SNC
        flags[k] = fl
    Gamma = None
    if q:
        dt = 10.0
        Gamma = np.zeros((steps, n, q))
        for i in range(q):
            Gamma[:, i, i] = dt * dt / 2
            if n - n // 2 + i < n:
                Gamma[:, n - n // 2 + i, i] = dt
    class NoQ:  # a Noise without a process matrix (q = 0): SNC is never enabled
        def ProcessMatrix(self):
            return None

        def MeasurementMatrix(self):
            return R
    kf, est0 = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R) if q else NoQ(), m, n_filters=nf)
    est = kf.RunBatch(flags, Phi, Ht, real, comp, Gamma, every_step=True)
    assert np.all(est.status == 0)
    refs = []
    for f in range(nf):
        o = oracle.NewHybridKF(np.zeros(n), P0, Q, R, m)
        refs.append(_oracle_run(o, flags, Phi, Ht, real, comp, Gamma, f, F_MEAS, F_EKF, F_SNC))
    _check(est, refs, ["state", "covar", "pred_covar", "gain", "innov", "obs_dev"], "hybrid")


@pytest.mark.parametrize("n,m,nf", [(6, 2, 130), (6, 2, 2), (4, 2, 258), (6, 3, 128)])
def test_hybrid_tma_staged_path_matches_oracle(oracle, n, m, nf):
    """The production path of the hybrid filter (final-estimate outputs, per-filter streams, no SNC)
    runs the warp-private TMA kernel (cp.async.bulk.tensor boxes -> shared memory, mbarrier ring per
    warp).  Ragged last warp / CTA (130 = 128 + 2 filters), Predict epochs (fewer boxes per epoch) and
    the CKF -> EKF switch are all covered; the result must equal the oracle (1e-10) and the plain-load
    and bulk-row kernels bit for bit."""
    import os
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(900 + n + nf)
    steps = 37
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 10.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    Q = np.diag(np.full(3, 1e-3))
    flags = np.array([(F_MEAS if (k % 6 != 4) else 0) | (F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)

    def run(path):
        if path:
            os.environ["GKB_NL_PATH"] = path
        try:
            kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R), m, n_filters=nf)
            est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
            vec, mat = kf.GetState()
        finally:
            os.environ.pop("GKB_NL_PATH", None)
        return est, vec, mat
    est, vec, mat = run(None)
    assert np.all(est.status == 0)
    for path in ("plain", "bulk"):
        est2, vec2, mat2 = run(path)
        assert np.array_equal(vec, vec2) and np.array_equal(mat, mat2), path
        assert np.array_equal(np.asarray(est.State()), np.asarray(est2.State())), path
    # the persistent scheduler's chunked mode (epochs of a group handed from warp to warp through the state
    # arrays) is what a 10^5-filter run uses; forced here on a small batch: bit-identical again
    for chunks in ("2", "5"):
        os.environ["GKB_NL_CHUNKS"] = chunks
        try:
            est3, vec3, mat3 = run(None)
        finally:
            del os.environ["GKB_NL_CHUNKS"]
        assert np.array_equal(vec, vec3) and np.array_equal(mat, mat3), chunks
        assert np.array_equal(np.asarray(est.State()), np.asarray(est3.State())), chunks
        assert np.array_equal(np.asarray(est.Covariance()), np.asarray(est3.Covariance())), chunks
    # every-step state / covariance outputs (what SmoothAll consumes) are streamed by the same TMA kernels
    def run_all(path, chunks=None):
        if path:
            os.environ["GKB_NL_PATH"] = path
        if chunks:
            os.environ["GKB_NL_CHUNKS"] = chunks
        try:
            kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R), m, n_filters=nf)
            e = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True, want=("state", "covar"))
        finally:
            os.environ.pop("GKB_NL_PATH", None)
            os.environ.pop("GKB_NL_CHUNKS", None)
        return np.asarray(e.State()).copy(), np.asarray(e.Covariance()).copy()
    xs_a, Ps_a = run_all(None)
    xs_p, Ps_p = run_all("plain")
    xs_c, Ps_c = run_all(None, "3")
    assert np.array_equal(xs_a, xs_p) and np.array_equal(Ps_a, Ps_p)
    assert np.array_equal(xs_a, xs_c) and np.array_equal(Ps_a, Ps_c)
    assert np.array_equal(xs_a.reshape(steps, n, nf)[-1], np.asarray(est.State()).reshape(n, nf))
    for f in sorted(set([0, 1, nf // 2, nf - 2, nf - 1])):
        o = oracle.NewHybridKF(np.zeros(n), P0, Q, R, m)
        ref = _oracle_run(o, flags, Phi, Ht, real, comp, None, f, F_MEAS, F_EKF, F_SNC)[-1]
        xs = np.asarray(est.State()).reshape(n, nf)[:, f]
        Pc = np.asarray(est.Covariance()).reshape(n, n, nf)[:, :, f]
        assert fx.scaled_err(xs, ref.State()) <= TOL, (f, fx.scaled_err(xs, ref.State()))
        assert fx.scaled_err(Pc, ref.Covariance()) <= TOL
        assert fx.scaled_err(vec[:, f], ref.State()) <= TOL and fx.scaled_err(mat[:, :, f], ref.Covariance()) <= TOL


@pytest.mark.parametrize("n,m,nf", [(6, 2, 130), (4, 2, 34), (6, 3, 64)])
def test_srif_tma_staged_path_matches_oracle(oracle, n, m, nf):
    """SRIF on the production path (final read-outs only, per-filter streams): the warp-private TMA
    kernel, with Predict epochs (dense R afterwards: the general LU inverse) and Update epochs
    (upper-triangular R: the straight-line production epoch, which takes b-bar = b instead of forming
    R-bar Phi inv(R) b).  Equal to the literal plain-load kernel to rounding (1e-12: the b-bar shortcut), bit for bit
    across its own scheduling variants, and to the oracle to 1e-10: State(), Covariance(), raw b and R."""
    import os
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(1900 + n + nf)
    steps = 33
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 50.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    flags = np.array([F_MEAS if (k % 5 != 3) else 0 for k in range(steps)], dtype=np.uint8)

    def run(path):
        if path:
            os.environ["GKB_NL_PATH"] = path
        try:
            kf, _ = gk.NewSRIF(0.2 * np.ones(n), P0, m, False, gk.NewNoiseless(np.zeros((n, n)), R), n_filters=nf)
            est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
            vec, mat = kf.GetState()
        finally:
            os.environ.pop("GKB_NL_PATH", None)
        return est, vec, mat
    est, vec, mat = run(None)
    assert np.all(est.status == 0)
    est2, vec2, mat2 = run("plain")
    assert np.array_equal(mat, mat2)  # R never sees the shortcut
    assert fx.scaled_err(vec, vec2) <= 1e-12 and fx.scaled_err(np.asarray(est.State()), np.asarray(est2.State())) <= 1e-12
    assert np.array_equal(np.asarray(est.Covariance()), np.asarray(est2.Covariance()))
    os.environ["GKB_NL_CHUNKS"] = "4"  # chunked persistent scheduling, forced
    try:
        est3, vec3, mat3 = run(None)
    finally:
        del os.environ["GKB_NL_CHUNKS"]
    assert np.array_equal(vec, vec3) and np.array_equal(mat, mat3)
    assert np.array_equal(np.asarray(est.Covariance()), np.asarray(est3.Covariance()))
    for f in sorted(set([0, 1, nf // 2, nf - 2, nf - 1])):
        o = oracle.NewSRIF(0.2 * np.ones(n), P0, m, False, R)
        ref = _oracle_run(o, flags, Phi, Ht, real, comp, None, f, F_MEAS, F_EKF, F_SNC)[-1]
        xs = np.asarray(est.State()).reshape(n, nf)[:, f]
        Pc = np.asarray(est.Covariance()).reshape(n, n, nf)[:, :, f]
        assert fx.scaled_err(xs, ref.State()) <= TOL, (f, fx.scaled_err(xs, ref.State()))
        assert fx.scaled_err(Pc, ref.Covariance()) <= TOL, (f, fx.scaled_err(Pc, ref.Covariance()))
        rv, rm, _ = ref.raw()
        assert fx.scaled_err(vec[:, f], rv) <= TOL and fx.scaled_err(mat[:, :, f], np.asarray(rm).reshape(n, n)) <= TOL


@pytest.mark.parametrize("n,m,nf", [(6, 2, 70), (4, 1, 40)])
def test_srif_tma_speculative_epoch_falls_back(oracle, n, m, nf):
    """The production SRIF kernel runs measurement epochs speculatively without row interchanges in the LU of Phi
    (srif_step_tri) and must hand the epoch to the general step, from the untouched shared-memory stage, when some
    lane needs one.  Here a few filters get a Phi with two rows exchanged at some epochs (partial pivoting has to
    swap), one filter gets a singular Phi (srif.go:112-114 error).  Equal to the literal plain-load kernel to rounding
    (b: 1e-12, the production epoch's b-bar = b; R: bit for bit), bit-equal across scheduling variants, 1e-10 to the
    oracle, and the singular filter reports the error without disturbing its neighbours."""
    import os
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(77 + n + nf)
    steps = 29
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    swapped = {(3, 1), (3, 33), (11, 33), (12, 33), (20, nf - 1), (28, 0)}  # (epoch, filter)
    for k, f in swapped:
        Phi[k, [0, n - 1], :, f] = Phi[k, [n - 1, 0], :, f]
    bad = 35
    Phi[9, 1, :, bad] = Phi[9, 0, :, bad]  # two equal rows: exactly singular
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 50.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    flags = np.array([F_MEAS if (k % 7 != 5) else 0 for k in range(steps)], dtype=np.uint8)

    def run(path, chunks=None):
        if path:
            os.environ["GKB_NL_PATH"] = path
        if chunks:
            os.environ["GKB_NL_CHUNKS"] = chunks
        try:
            kf, _ = gk.NewSRIF(0.2 * np.ones(n), P0, m, False, gk.NewNoiseless(np.zeros((n, n)), R), n_filters=nf)
            est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
            vec, mat = kf.GetState()
        finally:
            os.environ.pop("GKB_NL_PATH", None)
            os.environ.pop("GKB_NL_CHUNKS", None)
        return est, vec, mat
    est, vec, mat = run(None)
    good = np.array([f for f in range(nf) if f != bad])
    assert np.all(est.status[good] == 0) and est.status[bad] != 0
    for literal, other in ((True, run("plain")), (False, run(None, "3"))):
        est2, vec2, mat2 = other
        assert np.array_equal(est.status, est2.status)
        # b: rounding level.  (After a fallback a warp runs the literal epoch for the rest of its TASK; with forced
        # chunking the tasks are shorter, so a few more epochs take the production epoch's b-bar = b.)
        assert fx.scaled_err(vec[:, good], vec2[:, good]) <= 1e-12, literal
        assert np.array_equal(mat[:, :, good], mat2[:, :, good])
        assert np.array_equal(np.asarray(est.Covariance())[..., good], np.asarray(est2.Covariance())[..., good])
    for f in sorted(set([0, 1, 2, 32, 33, 34, 36, nf - 1])):
        o = oracle.NewSRIF(0.2 * np.ones(n), P0, m, False, R)
        ref = _oracle_run(o, flags, Phi, Ht, real, comp, None, f, F_MEAS, F_EKF, F_SNC)[-1]
        xs = np.asarray(est.State()).reshape(n, nf)[:, f]
        Pc = np.asarray(est.Covariance()).reshape(n, n, nf)[:, :, f]
        assert fx.scaled_err(xs, ref.State()) <= TOL, (f, fx.scaled_err(xs, ref.State()))
        assert fx.scaled_err(Pc, ref.Covariance()) <= TOL, (f, fx.scaled_err(Pc, ref.Covariance()))
        rv, rm, _ = ref.raw()
        assert fx.scaled_err(vec[:, f], rv) <= TOL and fx.scaled_err(mat[:, :, f], np.asarray(rm).reshape(n, n)) <= TOL


@pytest.mark.parametrize("n,m", [(6, 2), (4, 2), (3, 1), (8, 2), (7, 3)])
def test_srif_matches_oracle(oracle, n, m):
    """srif.go:101-160 with per-filter Phi / Htilde; State(), Covariance(), PredCovariance() read-outs."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(77 + n)
    nf, steps = 21, 40
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 50.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    flags = np.array([F_MEAS if (k % 4 != 2) else 0 for k in range(steps)], dtype=np.uint8)
    kf, est0 = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(np.zeros((n, n)), R), n_filters=nf)
    est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True)
    assert np.all(est.status == 0)
    refs = []
    for f in range(nf):
        o = oracle.NewSRIF(np.zeros(n), P0, m, False, R)
        refs.append(_oracle_run(o, flags, Phi, Ht, real, comp, None, f, F_MEAS, F_EKF, F_SNC))
    _check(est, refs, ["state", "covar", "pred_covar", "innov", "obs_dev"], "srif")


def test_srif_kats_on_device(oracle):
    """srif_test.go:15-29 (est0 covariance == P0) and the one-step drop-in API."""
    gk = _gpu()
    x0 = np.array([0, 0.35, 0])
    P0 = 10.0 * np.eye(3)
    R = np.diag([(5e-3) ** 2, (5e-6) ** 2])
    kf, est0 = gk.NewSRIF(x0, P0, 2, True, gk.NewNoiseless(np.zeros((6, 6)), R))
    assert np.max(np.abs(est0.Covariance() - P0)) <= 1e-12
    vec, mat = kf.GetState()
    o = oracle.NewSRIF(x0, P0, 2, True, R)
    rv, rm, _ = o.InitialEstimate().raw()
    assert fx.scaled_err(vec[:, 0], rv) <= 1e-14 and fx.scaled_err(mat[:, :, 0], rm) <= 1e-14
    with pytest.raises(gk.GkbError) as ei:  # locked until Prepare() (srif.go:102-104)
        kf.Update(np.zeros(2), np.zeros(2))
    assert ei.value.code == -4
    with pytest.raises(NotImplementedError):
        kf.SetNoise(gk.NewNoiseless(np.zeros((3, 3)), R))


def test_hybrid_basic_lock_and_toggle(oracle):
    """hybrid_test.go:15-54 TestHybridBasic: lock, EKF toggle, one Predict and one Update."""
    gk = _gpu()
    rng = np.random.default_rng(5)
    n, m = 6, 2
    P0 = np.diag([10, 10, 10, 1, 1, 1.0])
    R = np.diag([1e-6, 1e-6])
    Q = np.diag([1e-12] * 3)
    kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R), m)
    with pytest.raises(gk.GkbError):
        kf.Update(np.zeros(2), np.zeros(2))
    assert not kf.EKFEnabled()
    kf.EnableEKF()
    assert kf.EKFEnabled()
    kf.DisableEKF()
    o = oracle.NewHybridKF(np.zeros(n), P0, Q, R, m)
    Phi = np.eye(n) + 0.01 * rng.standard_normal((n, n))
    Ht = rng.standard_normal((m, n))
    kf.Prepare(Phi, None)
    o.Prepare(Phi, None)
    eg, eo = kf.Predict(), o.Predict()
    assert fx.scaled_err(eg.Covariance(), eo.Covariance()) <= TOL
    with pytest.raises(gk.GkbError):  # locked again after the call
        kf.Predict()
    kf.Prepare(Phi, Ht)
    o.Prepare(Phi, Ht)
    real, comp = np.array([1.0, -2.0]), np.array([0.9, -2.1])
    eg, eo = kf.Update(real, comp), o.UpdateNL(real, comp)
    for a, b in ((eg.State(), eo.State()), (eg.Covariance(), eo.Covariance()), (eg.Gain(), eo.Gain()),
                 (eg.Innovation(), eo.Innovation()), (eg.ObservationDev(), eo.ObservationDev())):
        assert fx.scaled_err(a, b) <= TOL
    assert eg.IsWithinNσ(1e6)


@pytest.mark.parametrize("kind,n,m,nf,shared", [("hybrid", 6, 2, 37, False), ("hybrid", 4, 2, 5, True), ("srif", 6, 2, 9, False),
                                                 ("hybrid", 8, 2, 11, False), ("srif", 7, 2, 6, False)])
def test_smooth_all_matches_oracle(oracle, kind, n, m, nf, shared):
    """SmoothAll (hybrid.go:209-238, srif.go:165-192): backward sweep over the stored estimates of a
    batched run, every smoothed state / covariance of every step against the oracle's restatement."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS
    rng = np.random.default_rng(4242 + n + nf)
    steps = 23
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    if shared:
        Phi = np.ascontiguousarray(Phi[:, :, :, 0])
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 10.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    flags = np.array([F_MEAS if k % 5 != 3 else 0 for k in range(steps)], dtype=np.uint8)
    if kind == "hybrid":
        kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(np.diag(np.full(3, 1e-3)), R), m, n_filters=nf)
    else:
        kf, _ = gk.NewSRIF(0.1 * np.ones(n), P0, m, False, gk.NewNoiseless(np.eye(n), R), n_filters=nf)
    est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True, want=("state", "covar"))
    x_fwd = np.asarray(est.State()).reshape(steps, n, nf).copy()
    P_fwd = np.asarray(est.Covariance()).reshape(steps, n, n, nf).copy()
    with pytest.raises(gk.GkbError):  # hybrid.go:210-212: wrong number of estimates
        kf.SmoothAll([est])
    assert kf.SmoothAll(est) is None
    xs = np.asarray(est.State()).reshape(steps, n, nf)
    Ps = np.asarray(est.Covariance()).reshape(steps, n, n, nf)
    assert np.array_equal(xs[-1], x_fwd[-1]) and np.array_equal(Ps[-1], P_fwd[-1])  # the last estimate is untouched
    for f in range(nf):
        Phi_f = Phi if shared else Phi[:, :, :, f]
        xr, Pr = oracle.smooth_all(np.ascontiguousarray(Phi_f), x_fwd[:, :, f], P_fwd[:, :, :, f])
        for k in range(steps):
            assert fx.scaled_err(xs[k, :, f], xr[k]) <= TOL, (f, k)
            assert fx.scaled_err(Ps[k, :, :, f], Pr[k]) <= TOL, (f, k)


def test_householder_and_srif_update_kats_on_device(oracle):
    """The reference's own known-answer tests run on the GPU routine: helper_test.go:108-117
    (HouseholderTransf, 1e-15) and srif_test.go:31-56 (measurementSRIFUpdate: A = [[R b],[H y]], 1e-4 on the
    four printed digits), plus a random batch against the oracle."""
    gk = _gpu()
    A = np.array([[1, -2, -1], [2, -1, 1], [1, 1, 2.0]])
    want = np.array([[-2.449489742783178, 1.224744871391589, -1.2247448713915892],
                     [0, -2.121320343559643, -2.121320343559643], [0, 0, 0]])
    got = gk.HouseholderTransf(A.copy(), 2, 1)
    assert np.max(np.abs(got - want)) <= 1e-15
    R, H = 0.1 * np.eye(2) * 0 + np.array([[0.1, 0], [0, 0.1]]), np.array([[1, -2.0], [2, -1], [1, 1]])
    b, y = np.array([0.2, 0.2]), np.array([-1.1, 1.2, 1.8])
    A2 = np.zeros((5, 3))
    A2[:2, :2], A2[:2, 2], A2[2:, :2], A2[2:, 2] = R, b, H, y
    got2 = gk.HouseholderTransf(A2.copy(), 2, 3)
    assert np.max(np.abs(got2[:2, :2] - np.array([[-2.4515, 1.2237], [0, -2.1243]]))) <= 1e-4
    assert np.max(np.abs(got2[:2, 2] - np.array([-1.2727, -2.0607]))) <= 1e-4
    assert np.max(np.abs(got2[2:, 2] - np.array([-0.1319, 0.0871, -0.2810]))) <= 1e-4
    rng = np.random.default_rng(8)
    batch = rng.standard_normal((8, 7, 50))
    ref = np.stack([oracle.householder_transf(batch[:, :, j].copy(), 6, 2) for j in range(50)], axis=2)
    got3 = gk.HouseholderTransf(batch.copy(), 6, 2)
    assert fx.scaled_err(got3, ref) <= 1e-13


@pytest.mark.parametrize("kind", ["hybrid", "srif"])
def test_host_stream_pipeline_is_bit_identical(kind, monkeypatch):
    """gkb_nl_run with HOST streams cuts the epochs into chunks that travel through two staging sets on a copy
    stream while the previous chunk's kernels run (engine.cu).  Forced to 3- and 7-epoch chunks (ragged last chunk,
    a 1-epoch tail) the results -- final and every-step outputs -- equal the single-shot call bit for bit."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF
    rng = np.random.default_rng(77)
    n, m, nf, steps = 6, 2, 70, 22
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    flags = np.array([(F_MEAS if k % 6 != 4 else 0) | (F_EKF if k >= 9 else 0) for k in range(steps)], dtype=np.uint8)
    P0, R = np.diag([10, 10, 10, 1, 1, 1.0]), np.diag([1e-2, 1e-2])

    def run(every):
        if kind == "hybrid":
            kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(np.diag([1e-12] * 3), R), m, n_filters=nf)
        else:
            kf, _ = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(np.diag([1e-12] * 3), R), n_filters=nf)
        est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=every, want=("state", "covar"))
        return est.State().copy(), est.Covariance().copy(), kf.GetState()
    monkeypatch.setenv("GKB_NL_H2D_CHUNK", "1000")
    ref = {every: run(every) for every in (False, True)}
    for chunk in ("3", "7"):
        monkeypatch.setenv("GKB_NL_H2D_CHUNK", chunk)
        for every in (False, True):
            x, P, (vec, mat) = run(every)
            if kind == "srif":
                # a 1-epoch tail launch (22 = 3 x 7 + 1) runs the literal general epoch, the others the production
                # epoch with b-bar = b: equal to rounding in b / State(), bit for bit in R / Covariance()
                assert fx.scaled_err(x, ref[every][0]) <= 1e-12 and fx.scaled_err(vec, ref[every][2][0]) <= 1e-12
            else:
                assert np.array_equal(x, ref[every][0]) and np.array_equal(vec, ref[every][2][0]), (chunk, every)
            assert np.array_equal(P, ref[every][1]) and np.array_equal(mat, ref[every][2][1]), (chunk, every)


def test_failed_epoch_writes_nan_rows_and_keeps_previous_estimate(oracle):
    """A filter whose epoch fails (here: a singular Phi for one filter at one epoch of an SRIF run, srif.go:112-114
    returns an error) keeps its previous estimate, reports the error in status, and its output rows of that epoch
    are NaN -- not stale data from an earlier call; the other filters and the later epochs are unaffected."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF, F_SNC
    rng = np.random.default_rng(5)
    n, m, nf, steps = 6, 2, 9, 12
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    Phi[5, :, :, 2] = 0.0
    P0, R = np.diag([50, 50, 50, 1, 1, 1.0]), np.diag([1e-2, 1e-2])
    flags = np.full(steps, F_MEAS, dtype=np.uint8)
    kf, _ = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(np.zeros((n, n)), R), n_filters=nf)
    kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True)      # an earlier call fills the staging buffers
    from gokalman_b200 import _lib
    _lib.check(_lib.load().gkb_reset(kf._h))
    est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True)
    assert est.status[2] == -5 and np.all(np.delete(est.status, 2) == 0)
    xs, Ps = est.State(), est.Covariance()
    assert np.all(np.isnan(xs[5, :, 2])) and np.all(np.isnan(Ps[5, :, :, 2]))
    assert np.all(np.isfinite(np.delete(xs, 2, axis=2))) and np.all(np.isfinite(xs[[0, 4, 6, 11], :, 2]))
    # the oracle's filter also keeps its previous estimate through the failed call and carries on
    o = oracle.NewSRIF(np.zeros(n), P0, m, False, R)
    for k in range(steps):
        o.Prepare(Phi[k, :, :, 2], Ht[k, :, :, 2])
        try:
            e = o.UpdateNL(real[k, :, 2], comp[k, :, 2])
        except Exception:
            assert k == 5
            continue
        assert fx.scaled_err(xs[k, :, 2], e.State()) <= TOL, k


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,shared", [(6, 2, False), (8, 2, False), (8, 3, False), (7, 2, True)])
def test_hybrid_failed_epoch_keeps_previous_estimate(oracle, n, m, shared):
    """hybrid.go:104-204 returns (nil, err) from a failed Update and leaves the filter as it was.  One filter gets a NaN
    observation at one epoch: that epoch's rows are NaN, status names the error, and every later epoch equals an oracle
    filter that never saw the epoch.  n = 7, 8 run the shared-memory-column kernel (Phi / H-tilde staged by cp.async,
    P-bar written out before the Joseph form): the previous covariance must survive there too."""
    gk = _gpu()
    from gokalman_b200._lib import F_MEAS, F_EKF
    rng = np.random.default_rng(77 + n + m)
    nf, steps, bad_k, bad_f = 37, 14, 5, 3
    Phi, Ht, real, comp = _od_streams(rng, n, m, nf, steps)
    if shared:
        Phi = np.ascontiguousarray(Phi[:, :, :, 0])
        Ht = np.ascontiguousarray(Ht[:, :, :, 0])
    real[bad_k, 0, bad_f] = np.nan
    P0 = np.diag(np.concatenate([np.full(n - n // 2, 10.0), np.full(n // 2, 1.0)]))
    R = np.diag(np.full(m, 1e-2))
    flags = np.array([F_MEAS | (F_EKF if k >= 8 else 0) for k in range(steps)], dtype=np.uint8)
    kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(np.diag(np.full(3, 1e-9)), R), m, n_filters=nf)
    est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True)
    assert est.status[bad_f] != 0 and np.all(np.delete(est.status, bad_f) == 0)
    xs, Ps, Pp = est.State(), est.Covariance(), est.PredCovariance()
    assert np.all(np.isnan(xs[bad_k, :, bad_f])) and np.all(np.isnan(Ps[bad_k, :, :, bad_f]))
    assert np.all(np.isnan(Pp[bad_k, :, :, bad_f]))
    assert np.all(np.isfinite(np.delete(xs, bad_f, axis=2)))
    for f in (bad_f, bad_f + 1):
        o = oracle.NewHybridKF(np.zeros(n), P0, np.diag(np.full(3, 1e-9)), R, m)
        for k in range(steps):
            if f == bad_f and k == bad_k:
                continue
            o.Prepare(Phi[k] if shared else Phi[k, :, :, f], Ht[k] if shared else Ht[k, :, :, f])
            (o.EnableEKF if flags[k] & F_EKF else o.DisableEKF)()
            e = o.UpdateNL(real[k, :, f], comp[k, :, f])
            assert fx.scaled_err(xs[k, :, f], e.State()) <= TOL, (f, k)
            assert fx.scaled_err(Ps[k, :, :, f], e.Covariance()) <= TOL, (f, k)
            assert fx.scaled_err(Pp[k, :, :, f], e.PredCovariance()) <= TOL, (f, k)
