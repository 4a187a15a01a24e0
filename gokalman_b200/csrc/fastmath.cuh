// fastmath.cuh -- branch-free FP64 building blocks for the filter kernels (sm_100a).
//
// The CUDA libm entry points (1.0/x, sqrt, log, sincospi) carry slow-path calls for denormals /
// infinities / huge arguments and load their polynomial coefficients from global tables
// (LDG.CONSTANT); inside a fully unrolled one-filter-per-thread loop that costs issue slots and
// instruction-cache space the FP64 pipe should be getting.  Every routine here is restricted to the
// argument range the kernels actually produce, uses the MUFU seed + Newton steps on the FP64 pipe and
// keeps its coefficients as literals (constant-bank operands of the DFMAs).
// Coefficients: tools/gen_math_coeffs.py (Chebyshev fits in 60-digit arithmetic); worst-case fit
// errors: sin 2.5e-18, cos 4.7e-17, log series 1.6e-16 * w (w <= 0.0295).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gkb {

// Accuracy of the MUFU seeds and of the refinements below was measured on B200 over 1e8 arguments
// log-uniform in [1e-10, 1e4] (tools/acc_probe): seeds ~1e-6 relative; rcp_nr and sqrt_nr return the
// correctly rounded IEEE result for every sample; rcp_fast is within 1 ulp.

// 1/x for normal, non-zero x: MUFU.RCP64H seed + two Newton steps.
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// 1/x to 1 ulp with one cubic step: y (1 + e + e^2).
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// sqrt(x) for normal positive x: MUFU.RSQ64H seed, one coupled Newton (Goldschmidt) step and the
// residual correction (itself a Newton step): 1e-6 -> 1e-12 -> below half an ulp.
__device__ __forceinline__ double sqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y;       // ~ sqrt(x)
  double h = 0.5 * y;     // ~ 1 / (2 sqrt(x))
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  return fma(d, h, g);
}

// exact conversion of a 32-bit unsigned integer scaled by 2^-shift, shift in {0, 30}: the "2^52 + k"
// construction, one DADD on the FP64 pipe instead of an I2F on the conversion unit.
__device__ __forceinline__ double u32_to_double(uint32_t k) {
  return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0;  // 2^52
}

// Polynomial coefficients of box_muller_fast.  A kernel may keep one copy in registers for its whole
// time loop (BmCoef::load_opaque) instead of re-materialising every literal (two UMOVs each) per use.
struct BmCoef {
  double L[7], S[7], C[6];
  double ln2_hi_m2, ln2_lo_m2;
  __device__ __forceinline__ void load() {
    L[0] = 0x1.5555555555558p-2; L[1] = 0x1.99999999952d7p-3; L[2] = 0x1.2492492df281ap-3;
    L[3] = 0x1.c71c62e3f11e6p-4; L[4] = 0x1.7462b51cb66b1p-4; L[5] = 0x1.39fe51a7c18f9p-4;
    L[6] = 0x1.2b5900de53b32p-4;
    S[0] = 0x1.921fb54442d18p-1; S[1] = -0x1.4abbce625be41p-4; S[2] = 0x1.466bc677587f8p-9;
    S[3] = -0x1.32d2cce2e5b19p-15; S[4] = 0x1.50782fda12d96p-22; S[5] = -0x1.e30071afc3e59p-30;
    S[6] = 0x1.e3f38399551bfp-38;
    C[0] = -0x1.3bd3cc9be458bp-2; C[1] = 0x1.03c1f081b0780p-6; C[2] = -0x1.55d3c7dbfd139p-12;
    C[3] = 0x1.e1f4fb60281f6p-19; C[4] = -0x1.a6c9c1be9eb49p-26; C[5] = 0x1.f3dbcea61b1a4p-34;
    ln2_hi_m2 = -2.0 * 0x1.62e42fefa3800p-1;
    ln2_lo_m2 = -2.0 * 0x1.ef35793c76730p-45;
  }
  // same values, but opaque to the compiler: they stay in registers across the caller's loop
  __device__ __forceinline__ void load_opaque() {
    load();
#pragma unroll
    for (int i = 0; i < 7; ++i) { asm volatile("mov.b64 %0, %0;" : "+d"(L[i])); asm volatile("mov.b64 %0, %0;" : "+d"(S[i])); }
#pragma unroll
    for (int i = 0; i < 6; ++i) asm volatile("mov.b64 %0, %0;" : "+d"(C[i]));
    asm volatile("mov.b64 %0, %0;" : "+d"(ln2_hi_m2));
    asm volatile("mov.b64 %0, %0;" : "+d"(ln2_lo_m2));
  }
};

// (z0, z1) = sqrt(-2 ln u0) * (cos, sin)(2 pi u1), u = (word + 0.5) * 2^-32: Box-Muller on two
// Philox words, no special cases (u0 in [2^-33, 1 - 2^-33], so -2 ln u0 in [2.3e-10, 45.8]).
__device__ __forceinline__ void box_muller_fast(const BmCoef& cf, uint32_t a, uint32_t b, double& z0, double& z1) {
  // ---- radius.  A = 2a + 1 is a 33-bit odd integer, u0 = A * 2^-33.
  const int hiA = 0x43300000 | (int)(a >> 31);
  const int loA = (int)((a << 1) | 1u);
  const double dA = __hiloint2double(hiA, loA) - 4503599627370496.0;  // exact
  int hi = __double2hiint(dA);
  const int lo = __double2loint(dA);
  int e = (hi >> 20) - 1023 - 33;  // u0 = m * 2^e, m in [1, 2)
  int mant = hi & 0x000fffff;
  const bool big = mant >= 0x6a09f;  // m > sqrt(2): use m / 2, e + 1, so that m in [0.7071, 1.4143]
  hi = (mant | 0x3ff00000) - (big ? 0x00100000 : 0);
  e += big ? 1 : 0;
  const double m = __hiloint2double(hi, lo);
  const double ef = __hiloint2double(0x43300000, e + 64) - 4503599627370560.0;  // (double)e, exact
  const double s = (m - 1.0) * rcp_fast(m + 1.0);
  const double w = s * s;
  double L = cf.L[6];
#pragma unroll
  for (int i = 5; i >= 0; --i) L = fma(L, w, cf.L[i]);
  const double half_lnm = fma(s * w, L, s);  // ln(m) / 2 = atanh(s)
  // x = -2 ln u0 = -2 e ln2 - 4 atanh(s), ln2 split so that e * ln2_hi is exact
  double x = fma(ef, cf.ln2_lo_m2, -4.0 * half_lnm);
  x = fma(ef, cf.ln2_hi_m2, x);
  const double r = sqrt_nr(x);
  // ---- angle.  theta = (b + 0.5) * 2^-32 turns; octant q, position inside the octant as the odd
  //      integer F2 in (0, 2^30): f = F2 * 2^-30 in (0, 1), measured from the nearer axis.
  const uint32_t q = b >> 29;
  const uint32_t rr2 = ((b & 0x1fffffffu) << 1) | 1u;
  const uint32_t F2 = (q & 1u) ? (0x40000000u - rr2) : rr2;
  const double f = __hiloint2double(0x41500000, (int)F2) - 4194304.0;  // F2 * 2^-30, exact (2^22 + .)
  const double t = f * f;
  double S = cf.S[6];
#pragma unroll
  for (int i = 5; i >= 0; --i) S = fma(S, t, cf.S[i]);
  double Cc = cf.C[5];
#pragma unroll
  for (int i = 4; i >= 0; --i) Cc = fma(Cc, t, cf.C[i]);
  Cc = fma(Cc, t, 1.0);
  const double sa = f * S;  // sin(f pi/4)
  const double ca = Cc;     // cos(f pi/4)
  const bool swap = ((q + 1u) >> 1) & 1u;           // octants 1, 2, 5, 6
  const uint32_t sneg = (q >> 2) & 1u;              // sin < 0 in octants 4..7
  const uint32_t cneg = ((q + 2u) >> 2) & 1u;       // cos < 0 in octants 2..5
  double sv = swap ? ca : sa;
  double cv = swap ? sa : ca;
  sv = __hiloint2double(__double2hiint(sv) ^ (int)(sneg << 31), __double2loint(sv));
  cv = __hiloint2double(__double2hiint(cv) ^ (int)(cneg << 31), __double2loint(cv));
  z0 = r * cv;
  z1 = r * sv;
}

}  // namespace gkb
