// philox.cuh -- counter-based Philox4x32-10 (Salmon et al., SC'11) + Box-Muller, device side.
//
// Replaces noise.go AWGN (noise.go:109-159), which draws from a time-seeded math/rand stream and
// is irreproducible by design.  Here the standard normals of (trial, step) are a pure function of
// (seed, global trial index, step): normals 4b..4b+3 come from Philox block
//     counter = (trial_lo, trial_hi, step, b),  key = (seed_lo, seed_hi)
// with u = (word + 0.5) * 2^-32 and (z0, z1) = sqrt(-2 ln u0) * (cos, sin)(2 pi u1), evaluated by the
// branch-free box_muller_fast (fastmath.cuh, ~1.5 ulp).
// The CPU oracle (test infrastructure) restates the same stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fastmath.cuh"

namespace gkb {

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
// that walks the steps of one trial computes them once.  pin() keeps the three words in registers (the compiler
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
__device__ __forceinline__ void philox_normals_trial(const BmCoef& cf, const PhiloxTrial& pt, uint64_t seed, uint32_t step,
                                                     double (&z)[COUNT]) {
  static_assert(COUNT <= 4, "one Philox block");
  uint32_t o[4];
  philox4x32_10_trial(pt, step, (uint32_t)seed, (uint32_t)(seed >> 32), o);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (2 * p < COUNT) {
      double z0, z1;
      box_muller_fast(cf, o[2 * p], o[2 * p + 1], z0, z1);
      z[2 * p] = z0;
      if (2 * p + 1 < COUNT) z[2 * p + 1] = z1;
    }
  }
}

// COUNT standard normals for (seed, trial, step).
template <int COUNT>
__device__ __forceinline__ void philox_normals(const BmCoef& cf, uint64_t seed, uint64_t trial, uint32_t step,
                                               double (&z)[COUNT]) {
  constexpr int BLOCKS = (COUNT + 3) / 4;
#pragma unroll
  for (int b = 0; b < BLOCKS; ++b) {
    uint32_t o[4];
    philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), step, (uint32_t)b, (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (4 * b + 2 * p < COUNT) {
        double z0, z1;
        box_muller_fast(cf, o[2 * p], o[2 * p + 1], z0, z1);
        z[4 * b + 2 * p] = z0;
        if (4 * b + 2 * p + 1 < COUNT) z[4 * b + 2 * p + 1] = z1;
      }
    }
  }
}

}  // namespace gkb
