"""Device-side orbit-determination inputs (gkb_od_synthesize) and the fused OD run (gkb_od_run).

  * the streams the device produces (closed-form RK4 STM, partials, observations) equal the oracle's GENERIC RK4 on
    the 6 + 36 state / variational equations to 1e-10, per step, on App. D's scenario (LEO a = 7000 km, i = 30 deg,
    three stations, 10 degree mask, sigma^2 = 1e-6);
  * the STM is a real STM: Phi maps a small initial perturbation onto the propagated difference of two orbits;
  * the fused run is BIT-IDENTICAL to synthesise-then-gkb_nl_run (production and strict kernels), hence inherits
    the strict kernel's 1e-10 parity against the oracle's hybrid filter run on the oracle's own streams.
"""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10
P0 = np.diag([10, 10, 10, 1, 1, 1.0])
R = np.diag([1e-6, 1e-6])
Q = np.diag([1e-12] * 3)


def _setup(nf, steps, dt=10.0, always=False):
    import gokalman_b200 as gk
    from gokalman_b200 import od
    gk.load()
    scn = od.Scenario(steps, dt, od.leo_truth0(), always_track=always, theta0=2.5)  # first pass at epoch 29
    orbit0 = od.perturbed_orbits(od.leo_truth0(), nf, seed=11)
    return gk, od, scn, orbit0


def test_od_streams_match_generic_rk4_oracle(oracle):
    gk, od, scn, orbit0 = _setup(nf=70, steps=120)
    Phi, Ht, real, comp, orb = od.synthesize(scn, orbit0, 1e-3, 1e-3, seed=5)
    rPhi, rHt, rreal, rcomp, rorb = oracle.od_synth(scn.mu, scn.j2, scn.re, scn.dt, orbit0, scn.station, scn.truth_obs,
                                                    1e-3, 1e-3, 5)
    for f in (0, 31, 32, 69):
        # the four 3 x 3 blocks of the STM have very different magnitudes (1, h, G h, 1): each is held to 1e-10 of
        # its own scale at every step
        P, rP = Phi[:, :, f].reshape(-1, 6, 6), rPhi[:, :, f].reshape(-1, 6, 6)
        for a, b in ((slice(0, 3), slice(0, 3)), (slice(0, 3), slice(3, 6)), (slice(3, 6), slice(0, 3)), (slice(3, 6), slice(3, 6))):
            assert fx.scaled_err_steps(P[:, a, b], rP[:, a, b]) <= TOL, (f, a, b)
        H, rH = Ht[:, :, f].reshape(-1, 2, 6), rHt[:, :, f].reshape(-1, 2, 6)
        assert fx.scaled_err_steps(H[:, 0, :3], rH[:, 0, :3]) <= TOL
        assert fx.scaled_err_steps(H[:, 1, :3], rH[:, 1, :3]) <= TOL
        assert np.array_equal(H[:, 1, 3:], H[:, 0, :3]) and not np.any(H[:, 0, 3:])
        assert fx.scaled_err_steps(real[:, :, f], rreal[:, :, f]) <= 1e-13
        assert fx.scaled_err_steps(comp[:, :, f], rcomp[:, :, f]) <= TOL
        assert fx.scaled_err(orb[:, f], rorb[:, f]) <= 1e-12
    # same epochs, same truth: the measurement noise differs per filter and has the requested spread
    noise = (real - scn.truth_obs[:, :, None])
    assert abs(noise.std() - 1e-3) < 5e-5 and abs(noise.mean()) < 5e-5


def test_od_stm_maps_perturbations():
    """Phi(t_{k+1}, t_k) dx_k ~= dx_{k+1} for two neighbouring reference orbits (first order in the 1 m offset)."""
    gk, od, scn, _ = _setup(nf=2, steps=50)
    x0 = od.leo_truth0()
    orbit0 = np.stack([x0, x0 + np.array([1e-3, -2e-3, 1.5e-3, 1e-6, 2e-6, -1e-6])], axis=1)
    _, _, _, _, orb_end = od.synthesize(scn, orbit0, 0.0, 0.0, seed=1)
    Phi, _, _, _, _ = od.synthesize(scn, orbit0, 0.0, 0.0, seed=1)
    dx = orbit0[:, 1] - orbit0[:, 0]
    for k in range(scn.steps):
        dx = Phi[k, :, 0].reshape(6, 6) @ dx
    d_end = orb_end[:, 1] - orb_end[:, 0]
    assert np.max(np.abs(dx - d_end) / np.max(np.abs(d_end))) < 1e-5
    # and the truth table agrees with the device propagation of the unperturbed orbit (same RK4, host numpy)
    assert fx.scaled_err(orb_end[:, 0], scn.truth[-1]) <= 1e-12


@pytest.mark.parametrize("strict", [False, True])
def test_od_fused_run_is_bit_identical_to_streams(oracle, strict):
    """gkb_od_run == gkb_od_synthesize + gkb_nl_run, bit for bit (both kernels call the same od_step), with the 10
    degree mask (Predict epochs between passes) and CKF -> EKF; and, for the strict kernel, 1e-10 against the oracle's
    hybrid filter fed with the ORACLE's streams."""
    gk, od, scn, orbit0 = _setup(nf=66, steps=400)
    assert 0 < np.count_nonzero(scn.flags & 1) < scn.steps  # passes and gaps
    Phi, Ht, real, comp, _ = od.synthesize(scn, orbit0, 1e-3, 1e-3, seed=9)
    nf = orbit0.shape[1]

    def make():
        kf, _ = gk.NewHybridKF(np.zeros(6), P0, gk.NewNoiseless(Q, R), 2, n_filters=nf)
        kf.SetStrict(strict)
        return kf
    a = make().RunBatch(scn.flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
    b = make().RunOD(scn, orbit0, 1e-3, 1e-3, seed=9)
    assert np.all(a.status == 0) and np.all(b.status == 0)
    assert np.array_equal(a.State(), b.State())
    assert np.array_equal(a.Covariance(), b.Covariance())
    # the (chunk, group) scheduler of the fused kernel, forced onto this small batch: chunk boundaries inside passes and gaps
    import os
    for chunks in ("1", "3", "7"):
        os.environ["GKB_NL_CHUNKS"] = chunks
        try:
            c = make().RunOD(scn, orbit0, 1e-3, 1e-3, seed=9)
        finally:
            del os.environ["GKB_NL_CHUNKS"]
        assert np.all(c.status == 0)
        assert np.array_equal(b.State(), c.State()) and np.array_equal(b.Covariance(), c.Covariance()), chunks
    # final outputs into PINNED caller buffers are written by the kernel itself (mapped host memory); same bits as staged
    import torch
    pinned = {"state": torch.zeros(6 * nf, dtype=torch.float64).pin_memory().numpy(),
              "covar": torch.zeros(36 * nf, dtype=torch.float64).pin_memory().numpy(),
              "status": torch.zeros(nf, dtype=torch.int32).pin_memory().numpy()}
    d = make().RunOD(scn, orbit0, 1e-3, 1e-3, seed=9, out_buffers=pinned)
    assert np.all(d.status == 0)
    assert np.array_equal(b.State(), d.State()) and np.array_equal(b.Covariance(), d.Covariance())
    os.environ["GKB_OD_STAGED_OUTPUTS"] = "1"
    try:
        e = make().RunOD(scn, orbit0, 1e-3, 1e-3, seed=9, out_buffers=pinned)
    finally:
        del os.environ["GKB_OD_STAGED_OUTPUTS"]
    assert np.array_equal(b.State(), e.State()) and np.array_equal(b.Covariance(), e.Covariance())
    if strict:
        rPhi, rHt, rreal, rcomp, _ = oracle.od_synth(scn.mu, scn.j2, scn.re, scn.dt, orbit0[:, :4], scn.station, scn.truth_obs,
                                                     1e-3, 1e-3, 9)
        xr, Pr = oracle.run_nl_batch(oracle.HYBRID, np.zeros(6), P0, R, scn.flags, rPhi, rHt, rreal, rcomp, threads=2)
        xf, Pf = oracle.run_nl_batch(oracle.HYBRID, np.zeros(6), P0, R, scn.flags, np.ascontiguousarray(Phi[:, :, :4]),
                                     np.ascontiguousarray(Ht[:, :, :4]), np.ascontiguousarray(real[:, :, :4]),
                                     np.ascontiguousarray(comp[:, :, :4]), threads=2)
        for f in range(4):
            # the filter arithmetic: GPU strict vs the oracle on the SAME (device-made) streams -- the parity bar
            assert fx.scaled_err(b.State()[:, f], xf[:, f]) <= TOL
            assert fx.scaled_err(b.Covariance()[:, :, f].reshape(-1), Pf[:, f]) <= TOL
        print("end-to-end (oracle streams + oracle filter) vs fused device run, x / P:",
              [(fx.scaled_err(b.State()[:, f], xr[:, f]), fx.scaled_err(b.Covariance()[:, :, f].reshape(-1), Pr[:, f])) for f in range(4)])


def test_od_run_converges_on_the_truth():
    """Physics sanity of the whole fused path (CKF, un-rectified reference orbits 50 m / 5 cm/s off the truth): after
    2000 epochs with six station passes the reference orbits have drifted kilometres from the truth, and the
    estimated deviation x-hat brings them back to within metres (same figures as the CPU oracle: median 3 m)."""
    import gokalman_b200 as gk
    from gokalman_b200 import od
    gk.load()
    nf, steps = 64, 2000
    scn = od.Scenario(steps, 10.0, od.leo_truth0(), theta0=2.5)
    orbit0 = od.perturbed_orbits(od.leo_truth0(), nf, sigma_r=0.05, sigma_v=5e-5, seed=11)
    kf, _ = gk.NewHybridKF(np.zeros(6), P0, gk.NewNoiseless(Q, R), 2, n_filters=nf)
    flags = (scn.flags & ~np.uint8(2))  # CKF throughout: x-hat accumulates the deviation of the un-rectified reference
    est = kf.RunOD(scn, orbit0, 1e-3, 1e-3, seed=3, flags=flags)
    assert np.all(est.status == 0)
    _, _, _, _, orb_end = od.synthesize(scn, orbit0, 1e-3, 1e-3, seed=3)
    drift = np.linalg.norm(orb_end[:3] - scn.truth[-1][:3, None], axis=0)
    err = np.linalg.norm((orb_end + est.State())[:3] - scn.truth[-1][:3, None], axis=0)
    assert np.median(drift) > 1.0 and np.median(err) < 0.02, (np.median(drift), np.median(err))


def test_bench_scenario_strict_is_exact_and_production_is_within_the_reference_fma_spread(oracle):
    """The bench's own configuration (bench_hybrid.od_scenario: every epoch a measurement, sigma = 1e-3 against P0 = 10,
    CKF -> EKF after 15 epochs) drives the conventional covariance form to cond(P) ~ 1e13: the answer is not determined
    to better than a few percent by ANY rounding sequence but the reference's own.  On 512 filters x 200 epochs:
      (1) strict kernel == oracle on the same streams, EXACTLY (the parity bar of the bench headline);
      (2) labelled "production-vs-strict": the production (FMA) kernel moves the result no further than the reference's
          own formulas move when gcc merely contracts a*b+c (oracle FMA build vs oracle unfused build, same streams,
          same per-filter scaled metric) -- median within 4x of that spread.  This is a sensitivity statement, not parity."""
    import gokalman_b200 as gk
    from gokalman_b200 import od
    gk.load()
    nf, steps = 512, 200
    scn = od.Scenario(steps, 10.0, od.leo_truth0(), always_track=True, theta0=2.5)
    orbit0 = od.perturbed_orbits(od.leo_truth0(), nf, sigma_r=1.0, sigma_v=1e-3, seed=1234)
    Phi, Ht, real, comp, _ = od.synthesize(scn, orbit0, 1e-3, 1e-3, seed=1234)

    def run(strict):
        kf, _ = gk.NewHybridKF(np.zeros(6), P0, gk.NewNoiseless(Q, R), 2, n_filters=nf)
        kf.SetStrict(strict)
        e = kf.RunBatch(scn.flags, Phi, Ht, real, comp, None, every_step=False, want=("state", "covar"))
        assert np.all(e.status == 0)
        return e.State(), e.Covariance().reshape(36, nf)
    xs, Ps = run(True)
    xp, Pp = run(False)
    xr, Pr = oracle.run_nl_batch(oracle.HYBRID, np.zeros(6), P0, R, scn.flags, Phi, Ht, real, comp, threads=4)
    xf, Pf = oracle.run_nl_batch(oracle.HYBRID, np.zeros(6), P0, R, scn.flags, Phi, Ht, real, comp, threads=4, fma=True)
    assert np.array_equal(xs, xr) and np.array_equal(Ps, Pr)  # (1): bit for bit

    def spread(a, b):
        return np.abs(a - b).max(axis=0) / np.abs(b).max(axis=0)
    prod = np.maximum(spread(xp, xs), spread(Pp, Ps))
    ref = np.maximum(spread(xf, xr), spread(Pf, Pr))
    cond = np.median([np.linalg.cond(Pr[:, j].reshape(6, 6)) for j in range(32)])
    print("bench scenario, %d filters x %d epochs: cond(P) median %.1e; production-vs-strict median %.2e p99 %.2e; "
          "reference formulas FMA-vs-unfused median %.2e p99 %.2e" % (nf, steps, cond, np.median(prod), np.quantile(prod, 0.99),
                                                                     np.median(ref), np.quantile(ref, 0.99)))
    assert cond > 1e10                                  # the premise: this run IS ill-conditioned
    assert np.median(ref) > 1e-6                        # ... and the reference's own formulas are rounding-sensitive on it
    assert np.median(prod) <= 4.0 * np.median(ref)      # (2)
