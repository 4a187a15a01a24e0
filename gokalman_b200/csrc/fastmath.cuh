// fastmath.cuh -- branch-free FP64 building blocks for the filter kernels (sm_100a).
//
// The CUDA libm entry points (1.0/x, sqrt) carry slow-path calls for denormals /
// infinities / huge arguments and load their polynomial coefficients from global tables
// (LDG.CONSTANT); inside a fully unrolled one-filter-per-thread loop that costs issue slots and
// instruction-cache space the FP64 pipe should be getting.  Every routine here is restricted to the
// argument range the kernels actually produce and uses the MUFU seed + Newton steps on the FP64 pipe.
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

}  // namespace gkb
