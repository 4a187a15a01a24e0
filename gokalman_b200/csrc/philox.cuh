// philox.cuh -- counter-based Philox4x32-10 (Salmon et al., SC'11) + piecewise-quintic inverse normal CDF.
//
// Replaces noise.go AWGN (noise.go:109-159), which draws from a time-seeded math/rand stream and
// is irreproducible by design.  Here the standard normals of (trial, step) are a pure function of
// (seed, global trial index, step): normal 4b + i is word i of the Philox block
//     counter = (trial_lo, trial_hi, step, b),  key = (seed_lo, seed_hi)
// pushed through the inverse normal CDF:  u = (word + 0.5) 2^-32,  z = Phi^-1(u), evaluated as a quintic in the
// low 27 bits of the normalised tail probability on one of 512 segments (16 per binary octave of min(u, 1 - u);
// table and construction: include/gokalman_b200_icdf.inc, tools/gen_icdf_table.py; max error 1.7e-12 absolute,
// |z| <= 6.34).  Per normal that is a handful of integer instructions, three 16-byte shared-memory loads and six
// FP64 instructions -- against ~43 FP64 instructions per Box-Muller pair (log, sqrt, sincos), which was 30 % of the
// Monte Carlo kernel's issue slots.  The CPU oracle (test infrastructure) includes the same table and evaluates the
// same fused Horner form, so its samples are bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gkb {

constexpr int kIcdfSegments = 512, kIcdfCoefs = 6;
// The table in global memory; kernels copy it into shared memory once per CTA (icdf_load).
__device__ const double kIcdfTable[kIcdfSegments * kIcdfCoefs] = {
#include "../../include/gokalman_b200_icdf.inc"
};

// Cooperative copy of the table into `tab` (shared memory, 16-byte aligned, kIcdfSegments * kIcdfCoefs doubles).
// The caller synchronises the CTA afterwards.
__device__ __forceinline__ void icdf_load(double* tab) {
  for (int i = threadIdx.x; i < kIcdfSegments * kIcdfCoefs; i += blockDim.x) tab[i] = kIcdfTable[i];
}

// One standard normal from one 32-bit word.
__device__ __forceinline__ double icdf_normal(uint32_t k, const double* __restrict__ tab) {
  const uint32_t upper = k >> 31;               // u > 1/2: z > 0
  const uint32_t j = upper ? ~k : k;            // 31 bits: tail probability p = (j + 0.5) 2^-32
  const uint32_t J = (j << 1) | 1u;             // p = J 2^-33, J odd, never zero
  const int lz = __clz((int)J);
  const uint32_t Jn = J << lz;                  // leading one at bit 31
  const uint32_t seg = ((31u - (uint32_t)lz) << 4) | ((Jn >> 27) & 15u);
  const double v = __hiloint2double(0x43300000, (int)(Jn & 0x07ffffffu)) - 4503599627370496.0;  // exact (2^52 + v)
  const double2* c = reinterpret_cast<const double2*>(tab + seg * kIcdfCoefs);
  const double2 c01 = c[0], c23 = c[1], c45 = c[2];
  double g = fma(c45.y, v, c45.x);
  g = fma(g, v, c23.y);
  g = fma(g, v, c23.x);
  g = fma(g, v, c01.y);
  g = fma(g, v, c01.x);                         // g = -Phi^-1(p) > 0
  return __hiloint2double(__double2hiint(g) ^ (int)((upper ^ 1u) << 31), __double2loint(g));
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The part of a Philox block that depends only on (seed, trial): with counter = (trial_lo, trial_hi, step, 0) the
// first round's products of c0 = trial_lo and its xor terms are the same for every step of a trial, so a kernel
// that walks the steps of one trial computes them once.  The three words are pinned in registers (the compiler
// would otherwise re-derive them from the thread index inside the loop).
struct PhiloxTrial {
  uint32_t lo0;   // lo(M0 * trial_lo): becomes c3 after round 1
  uint32_t n2;    // hi(M0 * trial_lo) ^ c3 (= block index 0) ^ k1: becomes c2 after round 1
  uint32_t a1;    // trial_hi ^ k0: xor-ed with hi(M1 * step) to give c0 after round 1
  __device__ __forceinline__ void init(uint64_t seed, uint64_t trial) {
    const uint32_t t0 = (uint32_t)trial;
    lo0 = 0xD2511F53u * t0;
    n2 = __umulhi(0xD2511F53u, t0) ^ (uint32_t)(seed >> 32);
    a1 = (uint32_t)(trial >> 32) ^ (uint32_t)seed;
    asm volatile("mov.b32 %0, %0;" : "+r"(lo0));
    asm volatile("mov.b32 %0, %0;" : "+r"(n2));
    asm volatile("mov.b32 %0, %0;" : "+r"(a1));
  }
};

// Block 0 of (seed, trial, step) = philox4x32_10(trial_lo, trial_hi, step, 0; seed) with round 1 taken from `pt`.
__device__ __forceinline__ void philox4x32_10_trial(const PhiloxTrial& pt, uint32_t step, uint32_t k0, uint32_t k1,
                                                    uint32_t (&out)[4]) {
  uint32_t c0 = __umulhi(0xCD9E8D57u, step) ^ pt.a1, c1 = 0xCD9E8D57u * step, c2 = pt.n2, c3 = pt.lo0;
  k0 += 0x9E3779B9u;
  k1 += 0xBB67AE85u;
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// COUNT <= 4 standard normals of (seed, trial, step): same values as philox_normals, round 1 hoisted.
template <int COUNT>
__device__ __forceinline__ void philox_normals_trial(const double* __restrict__ tab, const PhiloxTrial& pt, uint64_t seed,
                                                     uint32_t step, double (&z)[COUNT]) {
  static_assert(COUNT <= 4, "one Philox block");
  uint32_t o[4];
  philox4x32_10_trial(pt, step, (uint32_t)seed, (uint32_t)(seed >> 32), o);
#pragma unroll
  for (int i = 0; i < COUNT; ++i) z[i] = icdf_normal(o[i], tab);
}

// COUNT standard normals for (seed, trial, step).
template <int COUNT>
__device__ __forceinline__ void philox_normals(const double* __restrict__ tab, uint64_t seed, uint64_t trial, uint32_t step,
                                               double (&z)[COUNT]) {
  constexpr int BLOCKS = (COUNT + 3) / 4;
#pragma unroll
  for (int b = 0; b < BLOCKS; ++b) {
    uint32_t o[4];
    philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), step, (uint32_t)b, (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * b + i < COUNT) z[4 * b + i] = icdf_normal(o[i], tab);
  }
}

}  // namespace gkb
