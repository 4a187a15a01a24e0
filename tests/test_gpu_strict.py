"""The north star's 1e-10 on the statOD configuration (BASELINE configs[3]: R = 1e-6 against P0 = diag(10,10,10,1,1,1)).

The production hybrid kernels use FMAs, a packed covariance and a restructured Joseph update; on this
configuration forming (I - K H) P-bar cancels to eps |P-bar|, so ANY change of rounding moves the result by
eps |P-bar| / |P+| -- more than 1e-10.  `gkb_set_strict` selects the reference-order twin (filters_strict.cuh:
dense products in the written order, no FMA contraction, dense Joseph form, AsSymDense).  Here:

  * strict vs the oracle at the PLAIN 1e-10, every Estimate field of every step, per-step scaling,
    on App. D constants -- small sizes with Predict / SNC epochs, and the full 10^5 x 200 run;
  * production vs strict on ALL 10^5 filters (GPU against GPU), explicitly labelled "production-vs-strict":
    its bar is the rounding sensitivity of the reference's own formulas, not the parity bar.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GETTERS = {"state": "State", "meas": "Measurement", "innov": "Innovation", "covar": "Covariance",
           "pred_covar": "PredCovariance", "gain": "Gain", "obs_dev": "ObservationDev"}
P0_APPD = np.diag([10, 10, 10, 1, 1, 1.0])   # hybrid_test.go:79
R_APPD = np.diag([1e-6, 1e-6])               # hybrid_test.go:77
Q_APPD = np.diag([1e-12] * 3)                # hybrid_test.go:76


def _oracle_run(oracle, flags, Phi, Ht, real, comp, Gamma, f, L):
    o = oracle.NewHybridKF(np.zeros(6), P0_APPD, Q_APPD, R_APPD, 2)
    ests = []
    for k in range(len(flags)):
        o.Prepare(Phi[k, :, :, f], Ht[k, :, :, f])
        (o.EnableEKF if flags[k] & L.F_EKF else o.DisableEKF)()
        if flags[k] & L.F_SNC:
            o.PreparePNT(Gamma[k])
        ests.append(o.UpdateNL(real[k, :, f], comp[k, :, f]) if flags[k] & L.F_MEAS else o.Predict())
    return ests


def _errs(est, refs, f, nf, fields):
    """worst per-step scaled error and worst strict per-entry relative error over the fields"""
    worst, worst_rel, crossings = 0.0, 0.0, []
    steps = len(refs)
    for fld in fields:
        g = getattr(est, GETTERS[fld])()
        rows = [np.asarray(getattr(refs[k], GETTERS[fld])()).reshape(-1) for k in range(steps)]
        width = max(r.size for r in rows)
        ref = np.stack([np.pad(r, (0, width - r.size)) for r in rows])
        got = (g[..., f] if nf > 1 else g).reshape(steps, -1)[:, :width]
        worst = max(worst, fx.scaled_err_steps(got, ref))
        rel, bad = fx.strict_rel_err(got, ref, TOL)
        worst_rel = max(worst_rel, rel)
        crossings += [(fld,) + b for b in bad]
    return worst, worst_rel, crossings


def test_strict_hybrid_appd_constants_every_step(oracle):
    """App. D constants (R = 1e-6, P0 = diag(10,10,10,1,1,1), Q = 1e-12 I3, Gamma = [dt^2/2 I; dt I]), the bench's
    stream construction, CKF for 15 measurement epochs then EKF (hybrid_test.go:65,270-273), some Predict() and
    SNC epochs: the strict path matches the oracle to the plain 1e-10 on every field of every step, with every
    step scaled by its own max-abs; the strict per-entry relative error is printed with its zero crossings."""
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench import np_od_streams
    nf, steps, n, m, dt = 40, 80, 6, 2, 10.0
    Phi, Ht, real, comp = np_od_streams(nf, steps, 4242)
    Phi, Ht = Phi.reshape(steps, n, n, nf), Ht.reshape(steps, m, n, nf)
    flags = np.zeros(steps, dtype=np.uint8)
    meas_seen = 0
    for k in range(steps):
        fl = 0 if k % 7 == 5 else L.F_MEAS
        if meas_seen >= 15:
            fl |= L.F_EKF
        if (fl & L.F_MEAS) and k % 3 == 1:
            fl |= L.F_SNC
        meas_seen += 1 if fl & L.F_MEAS else 0
        flags[k] = fl
    Gamma = np.zeros((steps, n, 3))
    Gamma[:, :3, :] = 0.5 * dt * dt * np.eye(3)
    Gamma[:, 3:, :] = dt * np.eye(3)
    fields = ("state", "innov", "obs_dev", "covar", "pred_covar", "gain")

    def run(strict):
        kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R_APPD), m, n_filters=nf)
        kf.SetStrict(strict)
        return kf.RunBatch(flags, Phi, Ht, real, comp, Gamma, every_step=True)
    est_s, est_p = run(True), run(False)
    assert np.all(est_s.status == 0) and np.all(est_p.status == 0)
    report = []
    for f in (0, 1, 17, 31, 32, nf - 1):
        refs = _oracle_run(oracle, flags, Phi, Ht, real, comp, Gamma, f, L)
        e_s, rel_s, cross = _errs(est_s, refs, f, nf, fields)
        e_p, _, _ = _errs(est_p, refs, f, nf, fields)
        report.append((f, e_s, rel_s, len(cross), e_p))
        assert e_s <= TOL, ("strict vs oracle", f, e_s)
    print("App. D small run (filter, strict scaled err, strict per-entry rel err, #entries > 1e-10, "
          "production-vs-oracle scaled err):", report)


def test_strict_hybrid_well_conditioned_matches_production(oracle):
    """On a well-conditioned run both paths sit at rounding level: strict vs oracle ~1e-15, production vs strict
    <= 1e-10 -- the two kernels compute the same filter."""
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    rng = np.random.default_rng(3)
    nf, steps, n, m = 33, 40, 6, 2
    Phi = np.eye(n)[None, :, :, None] + 0.02 * rng.standard_normal((steps, n, n, nf))
    Ht = rng.standard_normal((steps, m, n, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.05 * rng.standard_normal((steps, m, nf))
    flags = np.array([L.F_MEAS | (L.F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)
    R = np.diag([1e-2, 1e-2])

    def run(strict):
        kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R), m, n_filters=nf)
        kf.SetStrict(strict)
        return kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=True)
    es, ep = run(True), run(False)
    for name in ("State", "Covariance", "PredCovariance", "Gain"):
        a, b = np.asarray(getattr(es, name)()), np.asarray(getattr(ep, name)())
        for f in range(nf):
            assert fx.scaled_err_steps(b[..., f], a[..., f]) <= TOL, (name, f)
    o = oracle.NewHybridKF(np.zeros(n), P0_APPD, Q_APPD, R, m)
    for k in range(steps):
        o.Prepare(Phi[k, :, :, 5], Ht[k, :, :, 5])
        (o.EnableEKF if flags[k] & L.F_EKF else o.DisableEKF)()
        eo = o.UpdateNL(real[k, :, 5], comp[k, :, 5])
        assert fx.scaled_err(es.State()[k, :, 5], eo.State()) <= 1e-13
        assert fx.scaled_err(es.Covariance()[k, :, :, 5], eo.Covariance()) <= 1e-13


def test_strict_hybrid_full_size_plain_tolerance(oracle):
    """BASELINE configs[3] at full size: 10^5 filters x 200 epochs, CKF -> EKF after 15 epochs, statOD constants.
    (1) STRICT kernel vs the oracle on eight filters cut out of the batch (first / last warp, CTA boundaries, the
    ragged tail): plain 1e-10 on state and covariance -- the north star's bar, no sensitivity allowance.
    (2) production-vs-strict on ALL 10^5 filters, GPU against GPU: reported (median / 99th percentile / max of the
    per-filter scaled error) and bounded by PRODUCTION_VS_STRICT, which is NOT the parity bar: it is the spread the
    reference's own formulas show on these streams when a*b+c is merely fused (oracle built with -ffp-contract=fast
    against the same oracle unfused, measured below on the same eight filters)."""
    import torch
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench_hybrid import make_streams
    lib = gk.load()
    nf, steps, n, m = 100000, 200, 6, 2
    dev = torch.device("cuda", 0)
    Phi, Ht, real, comp = make_streams(torch, nf, steps, 99, dev)
    flags_np = np.array([L.F_MEAS | (L.F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)
    flags = torch.from_numpy(flags_np).to(dev)

    def run(strict):
        kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R_APPD), m, n_filters=nf)
        kf.SetStrict(strict)
        xs = torch.zeros(n, nf, dtype=torch.float64, device=dev)
        Ps = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
        st = torch.zeros(nf, dtype=torch.int32, device=dev)
        out = L.Outputs()
        out.mem, out.every_step = L.DEVICE, 0
        out.state, out.covar, out.status = xs.data_ptr(), Ps.data_ptr(), st.data_ptr()
        L.check(lib.gkb_nl_run(kf._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                               comp.data_ptr(), None, L.DEVICE, C.byref(out)))
        torch.cuda.synchronize()
        assert int((st != 0).sum().item()) == 0
        return xs, Ps, lib.gkb_last_kernel_ms()
    xs_s, Ps_s, ms_s = run(True)
    xs_p, Ps_p, ms_p = run(False)
    pick = [0, 1, 31, 32, 49999, 77777, 99998, 99999]
    idx = torch.tensor(pick, device=dev)
    hPhi, hHt = Phi[:, :, idx].cpu().numpy(), Ht[:, :, idx].cpu().numpy()
    hreal, hcomp = real[:, :, idx].cpu().numpy(), comp[:, :, idx].cpu().numpy()
    xr, Pr = oracle.run_nl_batch(oracle.HYBRID, np.zeros(n), P0_APPD, R_APPD, flags_np, hPhi, hHt, hreal, hcomp, threads=4)
    xf, Pf = oracle.run_nl_batch(oracle.HYBRID, np.zeros(n), P0_APPD, R_APPD, flags_np, hPhi, hHt, hreal, hcomp, threads=4, fma=True)
    gx, gP = xs_s[:, idx].cpu().numpy(), Ps_s[:, idx].cpu().numpy()
    report, sens_max = [], 0.0
    for j in range(len(pick)):
        ex, eP = fx.scaled_err(gx[:, j], xr[:, j]), fx.scaled_err(gP[:, j], Pr[:, j])
        relx, _ = fx.strict_rel_err(gx[:, j], xr[:, j])
        relP, _ = fx.strict_rel_err(gP[:, j], Pr[:, j])
        sens = max(fx.scaled_err(xf[:, j], xr[:, j]), fx.scaled_err(Pf[:, j], Pr[:, j]))
        sens_max = max(sens_max, sens)
        report.append((pick[j], ex, eP, relx, relP, sens))
        assert ex <= TOL and eP <= TOL, ("strict vs oracle", report)
    print("strict full size (filter, scaled err x, scaled err P, per-entry rel x, per-entry rel P, "
          "oracle fma-vs-unfused spread):", report)
    print("strict kernel %.2f ms, production kernel %.2f ms for 10^5 x 200" % (ms_s, ms_p))

    # (2) production-vs-strict, every filter
    def per_filter_err(a, b):
        floor = b.abs().amax(dim=0, keepdim=True)
        return ((a - b).abs() / torch.maximum(b.abs(), floor)).amax(dim=0)
    e = torch.maximum(per_filter_err(xs_p, xs_s), per_filter_err(Ps_p, Ps_s))
    q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.float64, device=dev)).cpu().numpy()
    emax = float(e.max().item())
    print("production-vs-strict over %d filters: median %.2e, p99 %.2e, max %.2e (reference formulas' own "
          "fma-vs-unfused spread on the 8 probe filters: up to %.2e)" % (nf, q[0], q[1], emax, sens_max))
    # labelled "production-vs-strict": the rounding sensitivity of this ill-conditioned run, NOT the parity bar
    # (measured: median 9.4e-10, p99 3.0e-7, max 1.2e-4 over 10^5 filters -- the covariance of the reference's own
    # conventional formulas reaches condition numbers of 1e12 on these streams)
    PRODUCTION_VS_STRICT_MEDIAN, PRODUCTION_VS_STRICT_P99, PRODUCTION_VS_STRICT_MAX = 1e-8, 1e-5, 1e-2
    assert q[0] <= PRODUCTION_VS_STRICT_MEDIAN and q[1] <= PRODUCTION_VS_STRICT_P99 and emax <= PRODUCTION_VS_STRICT_MAX


def test_strict_hybrid_bench_configuration_full_length(oracle):
    """The bench's default workload itself -- BASELINE configs[3] as `bench.py` runs it: 10^5 filters x 1000 epochs of the
    device-synthesised statOD streams (41.6 GB), CKF -> EKF after 15 epochs -- through the strict kernel (scheduler on: 3125
    groups over 1184 resident warps) and eight filters cut out of it replayed through the oracle on the SAME streams: plain
    1e-10 on the final state and covariance after 1000 epochs (measured: exactly 0)."""
    import torch
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench_hybrid import make_streams_od, SIGMA
    lib = gk.load()
    nf, steps, n, m = 100000, 1000, 6, 2
    dev = torch.device("cuda", 0)
    Phi, Ht, real, comp, scn, _ = make_streams_od(torch, L, lib, nf, steps, 1234, 0)
    flags_np = np.ascontiguousarray(scn.flags[:steps])
    flags = torch.from_numpy(flags_np).to(dev)
    R = np.diag([SIGMA ** 2, SIGMA ** 2])
    kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R), m, n_filters=nf)
    kf.SetStrict(True)
    xs = torch.zeros(n, nf, dtype=torch.float64, device=dev)
    Ps = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
    st = torch.zeros(nf, dtype=torch.int32, device=dev)
    out = L.Outputs()
    out.mem, out.every_step = L.DEVICE, 0
    out.state, out.covar, out.status = xs.data_ptr(), Ps.data_ptr(), st.data_ptr()
    L.check(lib.gkb_nl_run(kf._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                           comp.data_ptr(), None, L.DEVICE, C.byref(out)))
    torch.cuda.synchronize()
    assert int((st != 0).sum().item()) == 0
    pick = [0, 31, 32, 4095, 50001, 77777, 99967, 99999]
    idx = torch.tensor(pick, device=dev)
    hPhi, hHt = Phi[:, :, idx].cpu().numpy(), Ht[:, :, idx].cpu().numpy()
    hreal, hcomp = real[:, :, idx].cpu().numpy(), comp[:, :, idx].cpu().numpy()
    del Phi, Ht, real, comp
    xr, Pr = oracle.run_nl_batch(oracle.HYBRID, np.zeros(n), P0_APPD, R, flags_np, hPhi, hHt, hreal, hcomp, threads=4)
    gx, gP = xs[:, idx].cpu().numpy(), Ps[:, idx].cpu().numpy()
    report = [(pick[j], fx.scaled_err(gx[:, j], xr[:, j]), fx.scaled_err(gP[:, j], Pr[:, j])) for j in range(len(pick))]
    print("strict, bench configuration 10^5 x 1000 (filter, scaled err x, scaled err P): %s; kernel %.2f ms"
          % (report, lib.gkb_last_kernel_ms()))
    assert all(ex <= TOL and eP <= TOL for _, ex, eP in report), report


def test_strict_on_srif_selects_the_literal_epoch(oracle):
    """gkb_set_strict on a GKB_SRIF handle runs the literal epoch of srif.go:101-160 (x-bar = Phi inv(R) b and
    b-bar = R-bar x-bar formed explicitly): bit-identical to the general kernel, 1e-10 to the oracle; the production
    epoch (b-bar = b) differs from it at rounding level only."""
    import os
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    rng = np.random.default_rng(21)
    n, m, nf, steps = 6, 2, 66, 30
    Phi = np.eye(n)[None, :, :, None] + 0.02 * rng.standard_normal((steps, n, n, nf))
    Ht = rng.standard_normal((steps, m, n, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.05 * rng.standard_normal((steps, m, nf))
    flags = np.full(steps, L.F_MEAS, dtype=np.uint8)
    P0, R = np.diag([50, 50, 50, 1, 1, 1.0]), np.diag([1e-2, 1e-2])

    def run(strict, path=None):
        if path:
            os.environ["GKB_NL_PATH"] = path
        try:
            kf, _ = gk.NewSRIF(0.1 * np.ones(n), P0, m, False, gk.NewNoiseless(np.zeros((n, n)), R), n_filters=nf)
            _lib_check = L.check(L.load().gkb_set_strict(kf._h, int(strict)))
            est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
            return est, kf.GetState()
        finally:
            os.environ.pop("GKB_NL_PATH", None)
    es, (bs, Rs) = run(True)
    ep, (bp, Rp) = run(False)
    el, (bl, Rl) = run(False, "plain")
    assert np.array_equal(bs, bl) and np.array_equal(Rs, Rl) and np.array_equal(es.State(), el.State())
    assert np.array_equal(Rs, Rp) and fx.scaled_err(bp, bs) <= 1e-12 and not np.array_equal(bp, bs)
    o = oracle.NewSRIF(0.1 * np.ones(n), P0, m, False, R)
    for k in range(steps):
        o.Prepare(Phi[k, :, :, 7], Ht[k, :, :, 7])
        eo = o.UpdateNL(real[k, :, 7], comp[k, :, 7])
    assert fx.scaled_err(es.State()[:, 7], eo.State()) <= TOL and fx.scaled_err(es.Covariance()[:, :, 7], eo.Covariance()) <= TOL


def test_strict_scheduler_and_register_twin_are_bit_identical():
    """The shared-memory strict kernel under its (chunk, group) scheduler -- forced onto a small ragged batch with 1, 2
    and 5 chunks, Predict epochs on chunk boundaries, final and every-step outputs -- returns the bits of the default
    launch and of the fully unrolled register version (GKB_STRICT_PATH=regs), the first implementation of the same step."""
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench import np_od_streams
    nf, steps, n, m = 77, 41, 6, 2
    Phi, Ht, real, comp = np_od_streams(nf, steps, 99)
    Phi, Ht = Phi.reshape(steps, n, n, nf), Ht.reshape(steps, m, n, nf)
    flags = np.array([(0 if k in (8, 9, 20) else L.F_MEAS) | (L.F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)

    def run(every, env):
        saved = {k: os.environ.get(k) for k in ("GKB_NL_CHUNKS", "GKB_STRICT_PATH")}
        os.environ.update(env)
        try:
            kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R_APPD), m, n_filters=nf)
            kf.SetStrict(True)
            est = kf.RunBatch(flags, Phi, Ht, real, comp, None, every_step=every)
            assert np.all(est.status == 0)
            return [np.asarray(getattr(est, g)()) for g in ("State", "Covariance", "PredCovariance", "Gain", "Innovation")]
        finally:
            for k, v in saved.items():
                os.environ.pop(k, None)
                if v is not None:
                    os.environ[k] = v
    for every in (False, True):
        ref = run(every, {})
        for env in ({"GKB_NL_CHUNKS": "1"}, {"GKB_NL_CHUNKS": "2"}, {"GKB_NL_CHUNKS": "5"}, {"GKB_STRICT_PATH": "regs"}):
            got = run(every, env)
            for a, b in zip(ref, got):
                assert np.array_equal(a, b), (every, env)


def test_strict_shared_streams_match_per_filter_streams_and_oracle(oracle):
    """Phi / H-tilde shared by the batch (the [steps, n, n] call shape: uniform loads instead of the cp.async stages) with
    per-filter observations: same bits as the same run fed with per-filter copies of the matrices, and exactly the oracle."""
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    rng = np.random.default_rng(17)
    nf, steps, n, m = 45, 30, 6, 2
    Phi = np.eye(n)[None] + 0.02 * rng.standard_normal((steps, n, n))
    Ht = rng.standard_normal((steps, m, n))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.05 * rng.standard_normal((steps, m, nf))
    flags = np.array([(0 if k == 11 else L.F_MEAS) | (L.F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)
    R = np.diag([1e-2, 1e-2])

    def run(Phi_, Ht_):
        kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R), m, n_filters=nf)
        kf.SetStrict(True)
        est = kf.RunBatch(flags, Phi_, Ht_, real, comp, None, every_step=True)
        assert np.all(est.status == 0)
        return est
    a = run(Phi, Ht)
    b = run(np.repeat(Phi[..., None], nf, axis=3), np.repeat(Ht[..., None], nf, axis=3))
    for g in ("State", "Covariance", "PredCovariance", "Gain", "Innovation"):
        assert np.array_equal(np.asarray(getattr(a, g)()), np.asarray(getattr(b, g)())), g
    for f in (0, 31, 32, nf - 1):
        o = oracle.NewHybridKF(np.zeros(n), P0_APPD, Q_APPD, R, m)
        for k in range(steps):
            o.Prepare(Phi[k], Ht[k])
            (o.EnableEKF if flags[k] & L.F_EKF else o.DisableEKF)()
            eo = o.UpdateNL(real[k, :, f], comp[k, :, f]) if flags[k] & L.F_MEAS else o.Predict()
            assert np.array_equal(a.State()[k, :, f], eo.State()), (f, k)
            assert np.array_equal(a.Covariance()[k, :, :, f], eo.Covariance()), (f, k)


def test_single_filter_handle_defaults_to_reference_arithmetic(oracle):
    """The drop-in use -- ONE HybridKF driven epoch by epoch through Prepare / Update / Predict, nothing set -- runs in
    reference-order arithmetic by default: on the bench's orbit scenario (cond(P) ~ 1e13, where the FMA kernel would be
    percent off after a few dozen epochs) every Estimate of every epoch equals the oracle's EXACTLY.  A batched handle
    keeps the production default (and differs); SetStrict(False) on the single handle selects it too."""
    import gokalman_b200 as gk
    from gokalman_b200 import od
    gk.load()
    steps = 80
    scn = od.Scenario(steps, 10.0, od.leo_truth0(), always_track=True, theta0=2.5)
    orbit0 = od.perturbed_orbits(od.leo_truth0(), 1, sigma_r=1.0, sigma_v=1e-3, seed=3)
    Phi, Ht, real, comp, _ = od.synthesize(scn, orbit0, 1e-3, 1e-3, seed=3)
    n, m = 6, 2
    o = oracle.NewHybridKF(np.zeros(n), P0_APPD, Q_APPD, R_APPD, m)

    def drive(kf):
        xs, Ps = [], []
        for k in range(steps):
            kf.Prepare(Phi[k, :, 0].reshape(n, n), Ht[k, :, 0].reshape(m, n))
            (kf.EnableEKF if k >= 15 else kf.DisableEKF)()
            e = kf.Update(real[k, :, 0], comp[k, :, 0])
            if isinstance(e, tuple):
                assert e[1] is None, e[1]
                e = e[0]
            xs.append(np.array(e.State()).reshape(-1))
            Ps.append(np.array(e.Covariance()).reshape(n, n))
        return np.array(xs), np.array(Ps)
    xr, Pr = [], []
    for k in range(steps):
        o.Prepare(Phi[k, :, 0].reshape(n, n), Ht[k, :, 0].reshape(m, n))
        (o.EnableEKF if k >= 15 else o.DisableEKF)()
        e = o.UpdateNL(real[k, :, 0], comp[k, :, 0])
        xr.append(np.array(e.State()).reshape(-1))
        Pr.append(np.array(e.Covariance()).reshape(n, n))
    xr, Pr = np.array(xr), np.array(Pr)
    kf, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R_APPD), m)
    xs, Ps = drive(kf)
    assert np.array_equal(xs, xr) and np.array_equal(Ps, Pr)
    kf2, _ = gk.NewHybridKF(np.zeros(n), P0_APPD, gk.NewNoiseless(Q_APPD, R_APPD), m)
    kf2.SetStrict(False)
    xp, Pp = drive(kf2)
    assert not np.array_equal(Pp, Pr)  # the production step is a different rounding sequence ...
    print("single filter, production vs reference after %d epochs: state %.2e, covariance %.2e (scaled)"
          % (steps, fx.scaled_err(xp[-1], xr[-1]), fx.scaled_err(Pp[-1], Pr[-1])))
