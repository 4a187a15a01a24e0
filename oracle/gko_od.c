/* gko_od.c -- CPU restatement of the engine's orbit-determination input synthesis (gkb_od_synthesize).
 * TEST INFRASTRUCTURE ONLY: nothing under gokalman_b200/ may include, link or call this file.
 *
 * The reference's OD callers obtain, for every epoch, the reference orbit, its state-transition matrix, the
 * range / range-rate partials and the computed observations from the external `smd` propagator
 * (hybrid_test.go:159-294); that code is not part of /root/reference, so there is no reference line to follow here.
 * This file states the engine's documented algorithm (include/gokalman_b200.h, "Orbit-determination inputs")
 * INDEPENDENTLY of the CUDA kernel's closed forms: the dynamics and the variational equations are integrated as the
 * GENERIC classical RK4 on all 6 + 36 equations (Phi' = A(r) Phi, Phi(t_k) = I), with plain unfused arithmetic.
 * "parity unpinned" in the sense of the task statement (no reference vectors exist for this step); what it pins is
 * the kernel's closed-form STM and partials against the textbook formulation. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gko.h"

static void od_accel(double mu, double kj2, const double* r, double* a, double* G /* 3x3 or NULL */) {
  double x = r[0], y = r[1], z = r[2];
  double r2 = x * x + y * y + z * z;
  double rn = sqrt(r2);
  double r3 = rn * r2, r5 = r3 * r2, r7 = r5 * r2, r9 = r7 * r2;
  double f = 1.0 / r5 - 5.0 * z * z / r7;
  double g = 3.0 / r5 - 5.0 * z * z / r7;
  a[0] = -mu * x / r3 - kj2 * x * f;
  a[1] = -mu * y / r3 - kj2 * y * f;
  a[2] = -mu * z / r3 - kj2 * z * g;
  if (!G) return;
  /* gradients of f and g */
  double dfdx = -5.0 * x / r7 + 35.0 * z * z * x / r9;
  double dfdy = -5.0 * y / r7 + 35.0 * z * z * y / r9;
  double dfdz = -15.0 * z / r7 + 35.0 * z * z * z / r9;
  double dgdx = -15.0 * x / r7 + 35.0 * z * z * x / r9;
  double dgdy = -15.0 * y / r7 + 35.0 * z * z * y / r9;
  double dgdz = -25.0 * z / r7 + 35.0 * z * z * z / r9;
  double rr[3] = {x, y, z};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[i * 3 + j] = mu * (3.0 * rr[i] * rr[j] / r5 - (i == j ? 1.0 / r3 : 0.0));
  G[0] += -kj2 * (f + x * dfdx); G[1] += -kj2 * x * dfdy; G[2] += -kj2 * x * dfdz;
  G[3] += -kj2 * y * dfdx; G[4] += -kj2 * (f + y * dfdy); G[5] += -kj2 * y * dfdz;
  G[6] += -kj2 * z * dgdx; G[7] += -kj2 * z * dgdy; G[8] += -kj2 * (g + z * dgdz);
}

/* derivative of y = [r(3), v(3), Phi(36 row-major)] */
static void od_deriv(double mu, double kj2, const double* y, double* dy) {
  double G[9];
  od_accel(mu, kj2, y, dy + 3, G);
  dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];
  const double* Phi = y + 6;
  double* dP = dy + 6;
  for (int j = 0; j < 6; ++j) {
    for (int i = 0; i < 3; ++i) dP[i * 6 + j] = Phi[(3 + i) * 6 + j]; /* top rows: d/dt Phi_r. = Phi_v. */
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
      for (int l = 0; l < 3; ++l) s += G[i * 3 + l] * Phi[l * 6 + j];
      dP[(3 + i) * 6 + j] = s;
    }
  }
}

int gko_od_synth(double mu, double j2, double re, double dt, int64_t nf, int steps, const double* orbit0,
                 const double* station, const double* truth_obs, double sigma_range, double sigma_rate, uint64_t seed,
                 int64_t filter_offset, double* Phi, double* Ht, double* real_obs, double* comp_obs, double* orbit_out) {
  const double kj2 = 1.5 * j2 * mu * re * re;
  for (int64_t f = 0; f < nf; ++f) {
    double y[42], k1[42], k2[42], k3[42], k4[42], t[42];
    for (int i = 0; i < 6; ++i) y[i] = orbit0[(int64_t)i * nf + f];
    for (int k = 0; k < steps; ++k) {
      for (int i = 0; i < 36; ++i) y[6 + i] = (i / 6 == i % 6) ? 1.0 : 0.0;
      od_deriv(mu, kj2, y, k1);
      for (int i = 0; i < 42; ++i) t[i] = y[i] + 0.5 * dt * k1[i];
      od_deriv(mu, kj2, t, k2);
      for (int i = 0; i < 42; ++i) t[i] = y[i] + 0.5 * dt * k2[i];
      od_deriv(mu, kj2, t, k3);
      for (int i = 0; i < 42; ++i) t[i] = y[i] + dt * k3[i];
      od_deriv(mu, kj2, t, k4);
      for (int i = 0; i < 42; ++i) y[i] = y[i] + dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
      for (int i = 0; i < 36; ++i) Phi[((int64_t)k * 36 + i) * nf + f] = y[6 + i];
      const double* st = station + (size_t)k * 6;
      double d[3], dv[3];
      for (int i = 0; i < 3; ++i) { d[i] = y[i] - st[i]; dv[i] = y[3 + i] - st[3 + i]; }
      double rho = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      double rdot = (d[0] * dv[0] + d[1] * dv[1] + d[2] * dv[2]) / rho;
      double H[12];
      for (int i = 0; i < 3; ++i) {
        H[i] = d[i] / rho;
        H[3 + i] = 0.0;
        H[6 + i] = dv[i] / rho - rdot * d[i] / (rho * rho);
        H[9 + i] = d[i] / rho;
      }
      for (int i = 0; i < 12; ++i) Ht[((int64_t)k * 12 + i) * nf + f] = H[i];
      double z[2];
      gko_philox_normals(seed, (uint64_t)(filter_offset + f), (uint32_t)k, 2, z);
      real_obs[((int64_t)k * 2) * nf + f] = truth_obs[(size_t)k * 2] + sigma_range * z[0];
      real_obs[((int64_t)k * 2 + 1) * nf + f] = truth_obs[(size_t)k * 2 + 1] + sigma_rate * z[1];
      comp_obs[((int64_t)k * 2) * nf + f] = rho;
      comp_obs[((int64_t)k * 2 + 1) * nf + f] = rdot;
    }
    if (orbit_out)
      for (int i = 0; i < 6; ++i) orbit_out[(int64_t)i * nf + f] = y[i];
  }
  return 0;
}
