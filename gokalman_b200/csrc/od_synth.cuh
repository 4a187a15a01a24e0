// od_synth.cuh -- per-filter orbit-determination inputs computed ON THE DEVICE: the step BEFORE the hybrid filter.
//
// In the reference's OD workflow (hybrid_test.go:159-294, examples/statOD5044) every epoch of every filter needs the
// state-transition matrix Phi of its reference orbit, the range / range-rate partials Htilde and the computed
// observation; the reference gets them from the `smd` propagator on the CPU, one filter at a time, and hands them
// to Prepare(Phi, Htilde) / Update(real, computed).  A batch of 10^5 filters streams 416 B per filter-epoch of such
// inputs; fed from the host that is PCIe-bound at ~1 % of the kernel's rate.  This file produces the same inputs
// from 48 B per filter (the initial reference orbit) plus small per-epoch tables shared by the batch:
//
//   dynamics     two-body + J2:  a(r) = -mu r/|r|^3 - k_J2 [x f, y f, z g],  f = 1/r^5 - 5 z^2/r^7, g = 3/r^5 - 5 z^2/r^7
//   propagation  ONE classical RK4 step of length h per epoch on the state AND on the variational equations
//                Phi' = [[0, I], [G(r), 0]] Phi, Phi(t_k) = I, G = da/dr (symmetric 3x3).  With Phi(t_k) = I the four
//                RK4 stages of the 36 STM entries collapse to closed forms in the stage gradients G0..G3:
//                  Phi_rr = I + h^2/6 (G0+G1+G2) + h^4/24 G2 G0        Phi_rv = h I + h^3/12 (G1+G2)
//                  Phi_vr = h/6 (G0+2G1+2G2+G3) + h^3/12 (G2 G0 + G3 G1)  Phi_vv = I + h^2/6 (G1+G2+G3) + h^4/24 G3 G1
//                (exactly what the generic RK4 on 42 equations computes, without its 84 temporaries)
//   measurement  range rho = |r - rs| and range-rate rho' = (r - rs).(v - vs)/rho to the epoch's tracking station
//                (ECI position / velocity rs, vs from a per-epoch table shared by the batch: hybrid_test.go:71-75),
//                Htilde = d(rho, rho')/d(r, v);  computed observation = (rho, rho') of the reference orbit;
//                real observation = the truth's noise-free (rho, rho') of the epoch (table) + sigma * N(0, 1) drawn
//                from Philox keyed by (seed, global filter index, epoch)  (hybrid_test.go:206: sigma^2 = 1e-6... here a parameter)
//
// od_step is inlined into the stream generator (od_synth_kernel) and into the fused filter kernel (od_run_kernel); it is
// written with explicit fma() for every multiply-add and plain single operations otherwise, so that no contraction
// choice is left to the compiler and the two kernels agree bit for bit ("synthesise, store, run gkb_nl_run on the
// streams" == "fused run": asserted in tests/test_gpu_od.py and in every bench run).
#pragma once
#include "engine_internal.h"
#include "philox.cuh"
#include "smallmat.cuh"

namespace gkb {

namespace od {

struct Grad { double xx, yy, zz, xy, xz, yz; };

// acceleration and its gradient at r
GKB_DEV void accel_grad(const OdParams& c, double x, double y, double z, double (&a)[3], Grad& G) {
  const double r2 = fma(x, x, fma(y, y, z * z));
  const double ir = rcp_nr(sqrt_nr(r2));
  const double ir2 = ir * ir, ir3 = ir2 * ir, ir5 = ir3 * ir2, ir7 = ir5 * ir2, ir9 = ir7 * ir2;
  const double z2 = z * z;
  const double f = fma(-5.0 * z2, ir7, ir5);
  const double g = fma(-5.0 * z2, ir7, 3.0 * ir5);
  const double tb = -c.mu * ir3;
  a[0] = fma(tb, x, -c.kj2 * x * f);
  a[1] = fma(tb, y, -c.kj2 * y * f);
  a[2] = fma(tb, z, -c.kj2 * z * g);
  const double q1 = fma(35.0 * z2, ir9, -5.0 * ir7);
  const double q2 = fma(35.0 * z2, ir9, -15.0 * ir7);
  const double q3 = fma(35.0 * z2, ir9, -25.0 * ir7);
  const double m5 = 3.0 * c.mu * ir5, m3 = c.mu * ir3;
  // (every a*b+c below is an explicit fma: the result must not depend on how a compiler contracts the expression,
  // because od_step is inlined into two different kernels that have to agree bit for bit)
  G.xx = fma(-c.kj2, fma(x * x, q1, f), fma(m5, x * x, -m3));
  G.yy = fma(-c.kj2, fma(y * y, q1, f), fma(m5, y * y, -m3));
  G.zz = fma(-c.kj2, fma(z * z, q3, g), fma(m5, z * z, -m3));
  G.xy = (x * y) * fma(-c.kj2, q1, m5);
  G.xz = (x * z) * fma(-c.kj2, q2, m5);
  G.yz = (y * z) * fma(-c.kj2, q2, m5);
}

GKB_DEV void grad_full(double (&M)[9], const Grad& G) {
  M[0] = G.xx; M[1] = G.xy; M[2] = G.xz;
  M[3] = G.xy; M[4] = G.yy; M[5] = G.yz;
  M[6] = G.xz; M[7] = G.yz; M[8] = G.zz;
}

}  // namespace od

// One epoch of one filter.  X[6]: reference orbit (r, v) at t_k in, at t_{k+1} out.  st[6]: station position /
// velocity (ECI) at t_{k+1}; tobs[2]: the truth's noise-free (range, range-rate) at t_{k+1}; z[2]: this filter's
// standard-normal draws for the epoch.  out[52]: Phi (36, row-major), Htilde (12), real (2), computed (2).
__device__ __forceinline__ void od_step(const OdParams& c, double* __restrict__ X, const double* __restrict__ st,
                                     const double* __restrict__ tobs, double z0, double z1, double* __restrict__ out) {
  const double h = c.h, hh = 0.5 * h;
  const double r0[3] = {X[0], X[1], X[2]}, v0[3] = {X[3], X[4], X[5]};
  double a0[3], a1[3], a2[3], a3[3];
  od::Grad G0, G1, G2, G3;
  od::accel_grad(c, r0[0], r0[1], r0[2], a0, G0);
  double r[3], v1[3], v2[3], v3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { r[i] = fma(hh, v0[i], r0[i]); v1[i] = fma(hh, a0[i], v0[i]); }
  od::accel_grad(c, r[0], r[1], r[2], a1, G1);
#pragma unroll
  for (int i = 0; i < 3; ++i) { r[i] = fma(hh, v1[i], r0[i]); v2[i] = fma(hh, a1[i], v0[i]); }
  od::accel_grad(c, r[0], r[1], r[2], a2, G2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { r[i] = fma(h, v2[i], r0[i]); v3[i] = fma(h, a2[i], v0[i]); }
  od::accel_grad(c, r[0], r[1], r[2], a3, G3);
  const double h6 = h * (1.0 / 6.0);
  double rn[3], vn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rn[i] = fma(h6, fma(2.0, v1[i] + v2[i], v0[i] + v3[i]), r0[i]);
    vn[i] = fma(h6, fma(2.0, a1[i] + a2[i], a0[i] + a3[i]), v0[i]);
    X[i] = rn[i];
    X[3 + i] = vn[i];
  }
  // STM over the epoch (closed form of RK4 on the variational equations with Phi(t_k) = I)
  double M0[9], M1[9], M2[9], M3[9], P20[9], P31[9];
  od::grad_full(M0, G0); od::grad_full(M1, G1); od::grad_full(M2, G2); od::grad_full(M3, G3);
  mul<3, 3, 3>(P20, M2, M0);
  mul<3, 3, 3>(P31, M3, M1);
  const double h2_6 = h * h * (1.0 / 6.0), h4_24 = h * h * h * h * (1.0 / 24.0), h3_12 = h * h * h * (1.0 / 12.0);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double id = (i == j) ? 1.0 : 0.0;
      const int e = i * 3 + j;
      out[kOdPhi + i * 6 + j] = fma(h4_24, P20[e], fma(h2_6, (M0[e] + M1[e]) + M2[e], id));             // Phi_rr
      out[kOdPhi + i * 6 + 3 + j] = fma(h3_12, M1[e] + M2[e], id * h);                                      // Phi_rv
      out[kOdPhi + (3 + i) * 6 + j] = fma(h3_12, P20[e] + P31[e], h6 * fma(2.0, M1[e] + M2[e], M0[e] + M3[e]));  // Phi_vr
      out[kOdPhi + (3 + i) * 6 + 3 + j] = fma(h4_24, P31[e], fma(h2_6, (M1[e] + M2[e]) + M3[e], id));   // Phi_vv
    }
  // range / range-rate to the epoch's station, their partials, the observations
  const double dx = rn[0] - st[0], dy = rn[1] - st[1], dz = rn[2] - st[2];
  const double dvx = vn[0] - st[3], dvy = vn[1] - st[4], dvz = vn[2] - st[5];
  const double rho2 = fma(dx, dx, fma(dy, dy, dz * dz));
  const double rho = sqrt_nr(rho2), irho = rcp_nr(rho);
  const double rdot = fma(dx, dvx, fma(dy, dvy, dz * dvz)) * irho;
  const double ux = dx * irho, uy = dy * irho, uz = dz * irho;
  out[kOdH + 0] = ux; out[kOdH + 1] = uy; out[kOdH + 2] = uz;
  out[kOdH + 3] = 0.0; out[kOdH + 4] = 0.0; out[kOdH + 5] = 0.0;
  out[kOdH + 6] = fma(-rdot, ux, dvx) * irho;
  out[kOdH + 7] = fma(-rdot, uy, dvy) * irho;
  out[kOdH + 8] = fma(-rdot, uz, dvz) * irho;
  out[kOdH + 9] = ux; out[kOdH + 10] = uy; out[kOdH + 11] = uz;
  out[kOdReal + 0] = fma(c.sigma[0], z0, tobs[0]);
  out[kOdReal + 1] = fma(c.sigma[1], z1, tobs[1]);
  out[kOdComp + 0] = rho;
  out[kOdComp + 1] = rdot;
}

}  // namespace gkb
