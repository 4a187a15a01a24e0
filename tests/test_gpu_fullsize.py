"""GPU checks at BASELINE.json's FULL sizes (the oracle cannot run these in seconds): size-independent
properties plus an exact (1e-10) oracle comparison on a handful of filters cut out of the full batch.

  configs[1]  Monte Carlo 10^6 trials x 1000 steps: the per-step sums over [0, 10^6) equal the sums over
              [0, 5*10^5) plus those over [5*10^5, 10^6) (Philox is keyed by the global trial index).
  configs[3]  hybrid CKF->EKF, 10^5 filters x 200 epochs streamed from HBM (device-resident C-ABI call).
              SRIF, same size, through the speculative straight-line epoch.
  configs[4]  32-state vanilla, 10^5 filters x 200 steps.
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


def test_mc_full_size_shard_sums_add_up():
    import gokalman_b200 as gk
    f = fx.jerk3()
    trials, steps = 1000000, 1000

    def sums(n_trials, offset):
        mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=0x5EED))
        kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
        runs = gk.NewMonteCarloRuns(n_trials, steps, 1, [np.zeros(1)], mckf, trial_offset=offset)
        return gk.NewChiSquare(kf, runs, [np.zeros(1)], True, True, sums=True)
    nis, nees = sums(trials, 0)
    nis_a, nees_a = sums(trials // 2, 0)
    nis_b, nees_b = sums(trials // 2, trials // 2)
    assert np.all(np.isfinite(nis)) and np.all(np.isfinite(nees))
    assert np.max(np.abs(nis_a + nis_b - nis) / nis) <= 1e-12
    assert np.max(np.abs(nees_a + nees_b - nees) / nees) <= 1e-12
    # chi-square sanity of the means (not exact: the reference's truth measurement lags the state by one step)
    assert 0.9 < nis.mean() / trials < 1.1 and 2.5 < nees.mean() / trials < 3.5


def test_hybrid_full_size_production_vs_oracle_labelled_bar(oracle):
    """PRODUCTION kernel (FMA, packed P, restructured Joseph) on the full-size statOD run against the oracle.  The
    parity bar proper (plain 1e-10) is held by the STRICT kernel in test_gpu_strict.py; the bar here is explicitly a
    "production-vs-reference-rounding" bar: max(1e-10, 8 x the spread the reference's own formulas show on the same
    filter between unfused and fused evaluation)."""
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
    P0 = np.diag([10, 10, 10, 1, 1, 1.0])
    R = np.diag([1e-6, 1e-6])
    kf, _ = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(np.diag([1e-12] * 3), R), m, n_filters=nf)
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
    Pm = Ps.reshape(n, n, nf)
    assert bool(torch.equal(Pm, Pm.transpose(0, 1)))              # stored symmetric
    assert bool((torch.diagonal(Pm, dim1=0, dim2=1) > 0).all())   # positive variances everywhere
    pick = [0, 1, 31, 32, 49999, 77777, 99998, 99999]             # first / last warp, CTA boundaries, the ragged tail
    idx = torch.tensor(pick, device=dev)
    hPhi, hHt = Phi[:, :, idx].cpu().numpy(), Ht[:, :, idx].cpu().numpy()
    hreal, hcomp = real[:, :, idx].cpu().numpy(), comp[:, :, idx].cpu().numpy()
    xr, Pr = oracle.run_nl_batch(oracle.HYBRID, np.zeros(n), P0, R, flags_np, hPhi, hHt, hreal, hcomp, threads=4)
    # This workload is ill-conditioned by construction (statOD: R = 1e-6 against P0 = 10, the covariance collapses
    # by seven orders of magnitude in the first updates): the reference's OWN formulas move by 1e-11 ... 1e-6 when
    # a*b+c is merely fused (Go on arm64 does that), so 1e-10 is not attainable by any FMA arithmetic here.  The
    # bar is therefore 1e-10 or 8x the reference formulas' measured FMA sensitivity on the same filter, whichever
    # is larger (the well-conditioned parity tests in test_gpu_parity_nl.py keep the plain 1e-10).
    xf, Pf = oracle.run_nl_batch(oracle.HYBRID, np.zeros(n), P0, R, flags_np, hPhi, hHt, hreal, hcomp, threads=4, fma=True)
    got_x, got_P = xs[:, idx].cpu().numpy(), Ps[:, idx].cpu().numpy()
    report = []
    for j in range(len(pick)):
        sens = max(fx.scaled_err(xf[:, j], xr[:, j]), fx.scaled_err(Pf[:, j], Pr[:, j]))
        ex, eP = fx.scaled_err(got_x[:, j], xr[:, j]), fx.scaled_err(got_P[:, j], Pr[:, j])
        report.append((pick[j], ex, eP, sens))
        assert ex <= max(TOL, 8 * sens), report
        assert eP <= max(TOL, 8 * sens), report
    print("hybrid full size (filter, err x, err P, reference FMA sensitivity):", report)


def test_srif_full_size_subset_matches_oracle(oracle):
    """configs[3], SRIF arm: 10^5 filters x 200 measurement epochs through the production kernel (the speculative
    straight-line epoch on the packed triangular R, srif_step_tri), eight filters cut out and replayed through the
    oracle at the PLAIN 1e-10 (the square-root form does not share the hybrid's cancellation: observed <= 2e-14); the
    reference formulas' own fused-vs-unfused spread is printed next to it."""
    import torch
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench_hybrid import make_streams
    lib = gk.load()
    nf, steps, n, m = 100000, 200, 6, 2
    dev = torch.device("cuda", 0)
    Phi, Ht, real, comp = make_streams(torch, nf, steps, 77, dev)
    flags_np = np.full(steps, L.F_MEAS, dtype=np.uint8)
    flags = torch.from_numpy(flags_np).to(dev)
    P0 = np.diag([50, 50, 50, 1, 1, 1.0])
    R = np.diag([1e-6, 1e-6])
    kf, _ = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(np.diag([1e-12] * 3), R), n_filters=nf)
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
    Pm = Ps.reshape(n, n, nf)
    assert bool(torch.equal(Pm, Pm.transpose(0, 1)))              # Covariance() = inv(R) inv(R)^T, mirrored
    assert bool((torch.diagonal(Pm, dim1=0, dim2=1) > 0).all())
    pick = [0, 1, 31, 32, 49999, 77777, 99998, 99999]
    idx = torch.tensor(pick, device=dev)
    hPhi, hHt = Phi[:, :, idx].cpu().numpy(), Ht[:, :, idx].cpu().numpy()
    hreal, hcomp = real[:, :, idx].cpu().numpy(), comp[:, :, idx].cpu().numpy()
    xr, Pr = oracle.run_nl_batch(oracle.SRIF, np.zeros(n), P0, R, flags_np, hPhi, hHt, hreal, hcomp, threads=4)
    xf, Pf = oracle.run_nl_batch(oracle.SRIF, np.zeros(n), P0, R, flags_np, hPhi, hHt, hreal, hcomp, threads=4, fma=True)
    got_x, got_P = xs[:, idx].cpu().numpy(), Ps[:, idx].cpu().numpy()
    report = []
    for j in range(len(pick)):
        sens = max(fx.scaled_err(xf[:, j], xr[:, j]), fx.scaled_err(Pf[:, j], Pr[:, j]))
        ex, eP = fx.scaled_err(got_x[:, j], xr[:, j]), fx.scaled_err(got_P[:, j], Pr[:, j])
        report.append((pick[j], ex, eP, sens))
        assert ex <= TOL, report
        assert eP <= TOL, report
    print("srif full size (filter, err x, err P, reference FMA sensitivity):", report)
    # production-vs-literal over ALL filters: the production epoch takes b-bar = b where the reference (and the general
    # kernel, GKB_NL_PATH=plain) forms R-bar Phi inv(R) b; R is untouched by that, b / State() move at rounding level
    os.environ["GKB_NL_PATH"] = "plain"
    try:
        kf2, _ = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(np.diag([1e-12] * 3), R), n_filters=nf)
        xs2 = torch.zeros(n, nf, dtype=torch.float64, device=dev)
        Ps2 = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
        out.state, out.covar = xs2.data_ptr(), Ps2.data_ptr()
        L.check(lib.gkb_nl_run(kf2._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                               comp.data_ptr(), None, L.DEVICE, C.byref(out)))
        torch.cuda.synchronize()
    finally:
        os.environ.pop("GKB_NL_PATH", None)
    assert bool(torch.equal(Ps, Ps2))
    ex_all = ((xs - xs2).abs() / torch.maximum(xs2.abs(), xs2.abs().amax(dim=0, keepdim=True))).amax()
    print("srif production-vs-literal over %d filters: covariance bit-identical, state max scaled diff %.2e" % (nf, float(ex_all)))
    assert float(ex_all) <= TOL


def test_vanilla32_full_size_subset_matches_oracle(oracle):
    import torch
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    lib = gk.load()
    nf, steps, n, m = 100000, 200, 32, 8
    dev = torch.device("cuda", 0)
    f = fx.synth_lti(n, m, seed=5)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    y = torch.randn(steps, nf, m, dtype=torch.float64, device=dev, generator=g)
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], None, f["H"], gk.NewNoiseless(f["Q"], f["R"]), n_filters=nf)
    xs = torch.zeros(nf, n, dtype=torch.float64, device=dev)
    Ps = torch.zeros(nf, n * n, dtype=torch.float64, device=dev)
    st = torch.zeros(nf, dtype=torch.int32, device=dev)
    out = L.Outputs()
    out.mem, out.every_step = L.DEVICE, 0
    out.state, out.covar, out.status = xs.data_ptr(), Ps.data_ptr(), st.data_ptr()
    L.check(lib.gkb_update(kf._h, steps, y.data_ptr(), 0, None, L.DEVICE, C.byref(out)))
    torch.cuda.synchronize()
    assert int((st != 0).sum().item()) == 0
    # the covariance recursion does not depend on the measurements: every filter must hold the same P, bit for bit
    assert bool((Ps == Ps[0:1]).all())
    pick = [0, 7, 8, 1183, 1184, 50000, 99999]  # first CTA, the first persistent-grid wrap-around, the tail
    idx = torch.tensor(pick, device=dev)
    hy = np.ascontiguousarray(y[:, idx, :].cpu().numpy())
    xr, Pr = oracle.run_vanilla_batch(f["x0"], f["P0"], f["F"], f["H"], f["Q"], f["R"], hy, threads=4)
    got_x, got_P = xs[idx].cpu().numpy(), Ps[idx].cpu().numpy()
    for j in range(len(pick)):
        assert fx.scaled_err(got_x[j], xr[j]) <= TOL, (pick[j], fx.scaled_err(got_x[j], xr[j]))
        assert fx.scaled_err(got_P[j], Pr[j]) <= TOL, (pick[j], fx.scaled_err(got_P[j], Pr[j]))
