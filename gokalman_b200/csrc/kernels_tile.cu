// kernels_tile.cu -- large-state Vanilla.Update (vanilla.go:128-220) for n = 16, 24, ... 64 (multiples of 8), m <= 8:
// one WARP per filter up to n = 32, a warp PAIR per filter above (TileShape, tile_run), covariance resident in
// shared memory for all the steps of a call, every dense product on the FP64 tensor-core path (mma.sync m8n8k4
// f64, "DMMA").  The stage table below is written for n = 32; the DMMA counts scale as in bench_tile.py:
// dmma_per_update (2048 at n = 64).
//
// BASELINE configs[4]: synthetic 32-state filters with a shared LTI model and a
// per-filter measurement stream.  Per update the reference does 8n^3 + ... = 331 k flop at n = 32, m = 8
// (SURVEY App. B); the only HBM traffic is the 8 m bytes of measurement per step, so the path is bound
// by the FP64 pipe.  Layout of one step of one warp (g = lane / 4, t = lane % 4: the DMMA fragment
// coordinates; every shared-memory matrix is XOR-swizzled so that A / B fragment loads and C-fragment
// loads / stores are all bank-conflict free; symmetric matrices keep only their upper triangle and the
// lower one is read through the transposed address):
//
//   T    = F P                   4x4 tiles x 8 k-steps = 128 DMMA   (bufA -> bufB); F x rides on the A fragments
//   P-   = T F^T + Q             upper 10 tiles x 8     =  80 DMMA   (bufB -> bufA)
//   PHt  = P- H^T                4 tiles x 8            =  32 DMMA   (H x, H x- ride on the B fragments)
//   S    = H PHt + R             1 tile x 8             =   8 DMMA   -> registers (C fragment)
//   S^-1                         Gauss-Jordan on the C fragment, warp shuffles only (S is SPD: no pivoting; cond <= 1e16 test)
//   K    = PHt S^-1              4 tiles x 2            =   8 DMMA   (S^-1 reaches the B fragment by shuffles)
//   T2   = P- - K PHt^T          upper 10 tiles x 2     =  20 DMMA   (Joseph form with the products by I removed,
//   V    = T2 H^T - K R          4 tiles x (8 + 2)      =  40 DMMA    as in filters.cuh: vanilla_step)
//   P+   = T2 - V K^T            upper 10 tiles x 2     =  20 DMMA   (-> bufA)
//
// 336 DMMA = 86 k FMA per update instead of the 166 k of the literal sequence.  m < 8 is handled by
// padding H with zero rows and R with a unit diagonal (the padded block of S is I and contributes nothing).
// Large-state handles use FILTER-MAJOR arrays (each filter's vector / matrix contiguous), see the header.
#include <cstdlib>

#include "engine_internal.h"
#include "fastmath.cuh"

namespace gkb {

namespace {

constexpr int kMP = 8;  // padded measurement size (one DMMA tile)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double v) {  // sum over the 4 lanes of a quad (t = 0..3)
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// ---- shared-memory layouts ------------------------------------------------------------------------------
// Every matrix is XOR-swizzled so that all three DMMA access patterns are bank-conflict free:
//   A pattern  (lane -> row g, col k0 + t)         LDS.64, 16 lanes per wavefront = 4 rows x 4 columns
//   B pattern  (lane -> row k0 + t, col n0 + g)    LDS.64, 4 rows x 4 columns
//   C pattern  (lane -> row g, cols n0 + 2t, +1)   LDS/STS.128, 8 lanes per wavefront = 2 rows x 8 columns
// n x n matrices: pitch LDB (multiple of 16), column c of row r lives at c ^ {0, 8, 4, 12}[r & 3];
// n x 8 matrices: pitch 8, column c of row r lives at c ^ (4 * ((r >> 1) & 1)).
__device__ __forceinline__ int swz_big(int r) { return ((r & 1) << 3) | ((r & 2) << 1); }
__device__ __forceinline__ int swz_small(int r) { return (r & 2) << 1; }
template <int LDB>
__device__ __forceinline__ int idx_big(int r, int c) { return r * LDB + (c ^ swz_big(r)); }
__device__ __forceinline__ int idx_small(int r, int c) { return r * 8 + (c ^ swz_small(r)); }

// Per-lane constants of the fragment addressing (g = lane / 4, t = lane % 4).
template <int LDB>
struct Frag {
  int g, t, swA, swB, sA8, sB8;
  bool up_a0, up_a1, up_b0, up_b1;  // diagonal tiles of an upper-stored symmetric matrix: is (row, col) upper?
  __device__ __forceinline__ explicit Frag(int lane) {
    g = lane >> 2;
    t = lane & 3;
    swA = swz_big(g);
    swB = swz_big(t);
    sA8 = swz_small(g);
    sB8 = swz_small(t);
    up_a0 = g <= t;       // A pattern, even k-step: local (row g, col t)
    up_a1 = g <= 4 + t;   //            odd k-step:  local (row g, col 4 + t)
    up_b0 = t <= g;       // B pattern, even k-step: local (row t, col g)
    up_b1 = 4 + t <= g;   //            odd k-step:  local (row 4 + t, col g)
  }
  // n x n matrices
  __device__ __forceinline__ int a(int rt, int ks) const { return (rt * 8 + g) * LDB + ((ks * 4) ^ swA) + t; }
  __device__ __forceinline__ int b(int ks, int ct) const { return (ks * 4 + t) * LDB + ((ct * 8 + g) ^ swB); }
  __device__ __forceinline__ int c(int rt, int ct) const { return (rt * 8 + g) * LDB + ((ct * 8 + 2 * t) ^ swA); }
  // symmetric n x n matrix of which only the upper triangle (tiles rt <= ct, and within the diagonal tiles
  // row <= col) is stored: element (row, col) with row > col is read at (col, row)
  __device__ __forceinline__ int sym_a(int rt, int ks) const {
    const int kt = ks >> 1;
    if (rt < kt) return a(rt, ks);
    if (rt > kt) return b(ks, rt);
    return ((ks & 1) ? up_a1 : up_a0) ? a(rt, ks) : b(ks, rt);
  }
  __device__ __forceinline__ int sym_b(int ks, int ct) const {
    const int kt = ks >> 1;
    if (kt < ct) return b(ks, ct);
    if (kt > ct) return a(ct, ks);
    return ((ks & 1) ? up_b1 : up_b0) ? b(ks, ct) : a(ct, ks);
  }
  // symmetric n x n matrix stored as its upper 8 x 8 TILES only, packed (tile (ti, tj), ti <= tj, at
  // 64 * pt(ti, tj); inside a tile the n x 8 layout): A / B / C patterns, the lower triangle through the transpose
  template <int TM>
  static __device__ __forceinline__ constexpr int pt(int ti, int tj) { return ti * TM - ti * (ti - 1) / 2 + (tj - ti); }
  template <int TM>
  __device__ __forceinline__ int p_c(int rt, int ct) const { return pt<TM>(rt, ct) * 64 + g * 8 + ((2 * t) ^ sA8); }
  template <int TM>
  __device__ __forceinline__ int p_a(int rt, int ks) const {  // element (rt*8 + g, ks*4 + t)
    const int kt = ks >> 1, kl = (ks & 1) * 4;
    const int direct = pt<TM>(rt < kt ? rt : kt, rt < kt ? kt : rt) * 64 + g * 8 + (kl ^ sA8) + t;
    const int transp = pt<TM>(rt < kt ? rt : kt, rt < kt ? kt : rt) * 64 + (kl + t) * 8 + (g ^ sB8);
    if (rt < kt) return direct;
    if (rt > kt) return transp;
    return ((ks & 1) ? up_a1 : up_a0) ? direct : transp;
  }
  template <int TM>
  __device__ __forceinline__ int p_b(int ks, int ct) const {  // element (ks*4 + t, ct*8 + g)
    const int kt = ks >> 1, kl = (ks & 1) * 4;
    const int direct = pt<TM>(kt < ct ? kt : ct, kt < ct ? ct : kt) * 64 + (kl + t) * 8 + (g ^ sB8);
    const int transp = pt<TM>(kt < ct ? kt : ct, kt < ct ? ct : kt) * 64 + g * 8 + (kl ^ sA8) + t;
    if (kt < ct) return direct;
    if (kt > ct) return transp;
    return ((ks & 1) ? up_b1 : up_b0) ? direct : transp;
  }
  // n x 8 matrices
  __device__ __forceinline__ int a8(int rt, int ks) const { return (rt * 8 + g) * 8 + ((ks * 4) ^ sA8) + t; }
  __device__ __forceinline__ int b8(int ks) const { return (ks * 4 + t) * 8 + (g ^ sB8); }
  __device__ __forceinline__ int c8(int rt) const { return (rt * 8 + g) * 8 + ((2 * t) ^ sA8); }
};

// element (r, c) of a tile-packed symmetric matrix (any r, c)
template <int TM>
__device__ __forceinline__ int idx_packed(int r, int c) {
  const int lo = r < c ? r : c, hi = r < c ? c : r;
  const int ti = lo >> 3, tj = hi >> 3;
  return (ti * TM - ti * (ti - 1) / 2 + (tj - ti)) * 64 + idx_small(lo & 7, hi & 7);
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, const double (&v)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

}  // namespace

// Warps per filter: one up to n = 32; a PAIR for n = 48 / 64, where a filter's matrices take 40-60 KB of shared memory
// and only three filters fit an SM -- one warp each would leave a scheduler idle.  The two warps of a pair split every
// n x n product by row-tile blocks (tile_owner: balanced for the full and for the upper-triangular products), the
// n x 8 products by row tiles, compute the 8 x 8 S and its inverse redundantly, and meet at a 64-thread named barrier
// wherever the single-warp version has a __syncwarp.
template <int N>
struct TileShape {
  static constexpr int TM = N / 8, KS = N / 4, LDB = (N + 15) / 16 * 16;
  static constexpr int PW = N <= 32 ? 1 : 2;
  // The n x n products are accumulated TB row-tiles at a time: all of them for n <= 32 (the accumulators of the whole
  // matrix fit the registers, and T2 stays there between the two Joseph stages), two (n = 64) or one (n = 48) otherwise.
  static constexpr int TB = TM <= 4 ? TM : (TM % 4 == 0 ? 2 : 1);
  static constexpr int kPacked = TM * (TM + 1) / 2 * 64;  // upper tiles of a symmetric n x n matrix
  // Pair mode keeps T = F P in REGISTERS across the barrier that ends the stage (each warp holds the row blocks it
  // owns) and passes it to the next stage one row block at a time through a warp-private scratch of TB row tiles; T2
  // and P+ then overwrite P- in place, so the full-size scratch matrix of the single-warp version disappears
  // (n = 64: 60.5 -> 44.2 KB per filter, four filters per SM instead of three).
  static constexpr bool TREG = PW == 2;
  static constexpr int kScratch = TREG ? PW * TB * 8 * LDB : N * LDB;
  static constexpr int kPerFilter = kPacked + kScratch + 2 * N * 8 + 2 * N + 2 * kMP;
  static_assert(N % 8 == 0 && N <= 64 && TM % TB == 0, "tile kernel shapes");
};
// owner of row-tile block b among the two warps of a pair: 0 1 1 0 0 1 1 0 ...
__host__ __device__ constexpr int tile_owner(int b) { return ((b & 3) == 0 || (b & 3) == 3) ? 0 : 1; }
// how many of the blocks before b belong to the same warp as b: the index of b among that warp's blocks
__host__ __device__ constexpr int tile_owned_index(int b) {
  int cnt = 0;
  for (int i = 0; i < b; ++i) cnt += tile_owner(i) == tile_owner(b) ? 1 : 0;
  return cnt;
}

template <int N, int HALF>
__device__ __forceinline__ void tile_run(const TileIo& io, const double* sF, const double* sH, const double* sR,
                                         double* fbase, int slot, int slots, int lane) {
  using SH = TileShape<N>;
  constexpr int TM = SH::TM, KS = SH::KS, LDB = SH::LDB, TB = SH::TB, PW = SH::PW, kPacked = SH::kPacked;
  constexpr bool TREG = SH::TREG;
  constexpr int NBM = TREG ? (TM / TB + 1) / 2 : 1;  // row blocks a warp of a pair owns at most
  const Frag<LDB> fr(lane);
  const int g = fr.g, t = fr.t;
  double* bufA = fbase;                            // P, then P-, then P+  (symmetric: packed upper tiles)
  double* bufB = bufA + kPacked;                   // T (full), then T2 (upper);  pair mode: the two row-block scratches
  double* sPH = bufB + SH::kScratch;               // P- H^T, later V
  double* wscr = bufB + (TREG ? HALF * TB * 8 * LDB : 0);  // pair mode: this warp's row block of T
  double* sK = sPH + N * 8;                        // gain
  double* xs = sK + N * 8;                         // posterior state
  double* xms = xs + N;                            // predicted state
  double* sinn = xms + N;                          // innovation (8), y-hat (8)
  // the warps of this filter meet here (named barrier 1 + slot; barrier 0 is __syncthreads)
  auto psync = [&]() {
    if constexpr (PW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(32 * PW) : "memory");
  };
  auto mine = [](int row_tile) { return PW == 1 || tile_owner(row_tile / TB) == HALF; };
  const int plane = lane + 32 * HALF;  // this thread among the filter's threads
  constexpr int PT = 32 * PW;

  const int m = io.m;
  for (int64_t f = (int64_t)blockIdx.x * slots + slot; f < io.nf; f += (int64_t)gridDim.x * slots) {
    // ---- state in: x -> xs, P -> bufA (coalesced: a filter's matrix is contiguous)
    for (int idx = plane; idx < N; idx += PT) xs[idx] = io.x[f * N + idx];
    {
      const double* Pg = io.P + f * (int64_t)(N * N);
      for (int idx = plane; idx < N * N; idx += PT)
        if (idx / N <= idx % N) bufA[idx_packed<TM>(idx / N, idx % N)] = Pg[idx];
    }
    psync();
    int status = 0;
    for (int k = 0; k < io.steps; ++k) {
      // this step's measurement: requested now, consumed after the gain (HBM latency hidden by the DMMA stages)
      double yv = 0.0;
      if (t == 0 && g < m) yv = io.y_shared ? __ldg(io.y + (int64_t)k * m + g) : __ldg(io.y + ((int64_t)k * io.nf + f) * m + g);
      // this lane's slice of the previous posterior, x[ks*4 + t]
      double xq[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) xq[ks] = xs[ks * 4 + t];
      double c[TB * TM][2];
      double tacc[TREG ? NBM * TB * TM : 1][2];  // pair mode: this warp's row blocks of T = F P
      // ---- T = F P (vanilla.go:149-150) and x- = F x (138-146; Noiseless, no control)
#pragma unroll
      for (int t0 = 0; t0 < TM; t0 += TB) {
        if (!mine(t0)) continue;
        double xpart[TB];
        double(*acc)[2] = TREG ? &tacc[tile_owned_index(t0 / TB) * TB * TM] : &c[0];
#pragma unroll
        for (int i = 0; i < TB * TM; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
        for (int i = 0; i < TB; ++i) xpart[i] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          double a[TB], b[TM];
#pragma unroll
          for (int ti = 0; ti < TB; ++ti) a[ti] = sF[fr.a(t0 + ti, ks)];
#pragma unroll
          for (int tj = 0; tj < TM; ++tj) b[tj] = bufA[fr.template p_b<TM>(ks, tj)];
#pragma unroll
          for (int ti = 0; ti < TB; ++ti) {
            xpart[ti] = fma(a[ti], xq[ks], xpart[ti]);
#pragma unroll
            for (int tj = 0; tj < TM; ++tj) dmma(acc[ti * TM + tj], a[ti], b[tj]);
          }
        }
#pragma unroll
        for (int ti = 0; ti < TB; ++ti) {
          const double s = quad_sum(xpart[ti]);  // + G u (vanilla.go:140-143), the same vector for every filter
          if (t == 0) xms[(t0 + ti) * 8 + g] = io.gu != nullptr ? s + __ldg(io.gu + (int64_t)k * N + (t0 + ti) * 8 + g) : s;
        }
        if constexpr (!TREG) {
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = 0; tj < TM; ++tj) sts2(bufB + fr.c(t0 + ti, tj), c[ti * TM + tj]);
        }
      }
      psync();  // pair mode: every read of P is done -- P- may overwrite it
      // ---- P- = T F^T + Q (150-152): upper tiles only (AsSymDense keeps the upper triangle, helper.go:65-84)
#pragma unroll
      for (int t0 = 0; t0 < TM; t0 += TB) {
        if (!mine(t0)) continue;
#pragma unroll
        for (int ti = 0; ti < TB; ++ti)
#pragma unroll
          for (int tj = t0 + ti; tj < TM; ++tj) {
            const double2 q = __ldg(reinterpret_cast<const double2*>(io.Q + ((t0 + ti) * 8 + g) * N + tj * 8 + 2 * t));
            c[ti * TM + tj][0] = q.x;
            c[ti * TM + tj][1] = q.y;
          }
        if constexpr (TREG) {  // this block of T: registers -> the warp's scratch (C layout), read back as A fragments
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = 0; tj < TM; ++tj) sts2(wscr + fr.c(ti, tj), tacc[(tile_owned_index(t0 / TB) * TB + ti) * TM + tj]);
          __syncwarp();
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          double a[TB], b[TM];
#pragma unroll
          for (int ti = 0; ti < TB; ++ti) a[ti] = TREG ? wscr[fr.a(ti, ks)] : bufB[fr.a(t0 + ti, ks)];
#pragma unroll
          for (int tj = t0; tj < TM; ++tj) b[tj] = sF[fr.a(tj, ks)];  // B[k][j] = F[j][k]
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = t0 + ti; tj < TM; ++tj) dmma(c[ti * TM + tj], a[ti], b[tj]);
        }
#pragma unroll
        for (int ti = 0; ti < TB; ++ti)
#pragma unroll
          for (int tj = t0 + ti; tj < TM; ++tj) sts2(bufA + fr.template p_c<TM>(t0 + ti, tj), c[ti * TM + tj]);
        if constexpr (TREG) __syncwarp();  // the scratch is free for the warp's next block
      }
      psync();
      if (io.o_pred != nullptr && (io.every_step || k == io.steps - 1)) {
        double* dst = io.o_pred + ((io.every_step ? (int64_t)k * io.nf : 0) + f) * (N * N);
        for (int idx = plane; idx < N * N; idx += PT) {
          const int r = idx / N, cc = idx % N;
          dst[idx] = bufA[idx_packed<TM>(r, cc)];
        }
      }
      // ---- PHt = P- H^T (160-161), y-hat = H x (155-157), H x- for the innovation (182-184)
      double yhat, hxm;
      {
        double xmq[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) xmq[ks] = xms[ks * 4 + t];
        double ph[TM][2];
#pragma unroll
        for (int ti = 0; ti < TM; ++ti) ph[ti][0] = ph[ti][1] = 0.0;
        double hx = 0.0, hm = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double b = sH[fr.a(0, ks)];  // B[k][a] = H[a][k]
          hx = fma(b, xq[ks], hx);
          hm = fma(b, xmq[ks], hm);
#pragma unroll
          for (int ti = 0; ti < TM; ++ti)
            if (mine(ti)) dmma(ph[ti], bufA[fr.template p_a<TM>(ti, ks)], b);
        }
        yhat = quad_sum(hx);
        hxm = quad_sum(hm);
#pragma unroll
        for (int ti = 0; ti < TM; ++ti)
          if (mine(ti)) sts2(sPH + fr.c8(ti), ph[ti]);
      }
      psync();
      // ---- S = H PHt + R (162-163), in the C-fragment layout: this lane holds S[g][2t], S[g][2t+1]
      double s[2];
      {
        const double2 r = lds2(sR + fr.c8(0));
        s[0] = r.x;
        s[1] = r.y;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma(s, sH[fr.a(0, ks)], sPH[fr.b8(ks)]);
      }
      // ---- S <- inv(S) (164-167): in-place Gauss-Jordan; S is symmetric positive definite, so the
      //      diagonal pivots are safe; a non-positive pivot is the reference's singular-S error
      // mat64.Dense.Inverse also fails when ||S||inf ||inv(S)||inf > 1e16 (the oracle's exact form of gonum's condition
      // test): infinity norms over the true m x m block (the padding rows carry a unit diagonal and must not count)
      auto inf_norm = [&](const double (&e)[2]) {
        double a = (g < m) ? fabs(e[0]) + fabs(e[1]) : 0.0;
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, off));
        return a;
      };
      const double s_norm = inf_norm(s);
      bool bad = false;
#pragma unroll
      for (int p = 0; p < kMP; ++p) {
        const double mine = (p & 1) ? s[1] : s[0];
        const double fcol = __shfl_sync(0xffffffffu, mine, (lane & ~3) | (p >> 1));  // S[g][p]
        const double d = __shfl_sync(0xffffffffu, mine, p * 4 + (p >> 1));           // S[p][p]
        const double p0 = __shfl_sync(0xffffffffu, s[0], p * 4 + t);                 // S[p][2t]
        const double p1 = __shfl_sync(0xffffffffu, s[1], p * 4 + t);                 // S[p][2t+1]
        bad = bad || !(d > 0.0) || !(d < 1e300);
        const double dinv = rcp_nr(d);
        const double r0 = (2 * t == p) ? dinv : p0 * dinv;
        const double r1 = (2 * t + 1 == p) ? dinv : p1 * dinv;
        if (g == p) {
          s[0] = r0;
          s[1] = r1;
        } else {
          s[0] = fma(-fcol, r0, (2 * t == p) ? 0.0 : s[0]);
          s[1] = fma(-fcol, r1, (2 * t + 1 == p) ? 0.0 : s[1]);
        }
      }
      // (the verdict is taken after the gain and state update below: neither is committed on an error, and the norm
      // reduction then overlaps their tensor-core work instead of lengthening the serial inverse)
      const double sinv_norm = inf_norm(s);
      // ---- K = PHt inv(S) (168): inv(S) is symmetric, so B[k][a] = Sinv[a][k] sits in this lane's own row
      double kc[TM][2];
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) kc[ti][0] = kc[ti][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int col = ks * 4 + t;
        const double e0 = __shfl_sync(0xffffffffu, s[0], (lane & ~3) | (col >> 1));
        const double e1 = __shfl_sync(0xffffffffu, s[1], (lane & ~3) | (col >> 1));
        const double b = (col & 1) ? e1 : e0;
#pragma unroll
        for (int ti = 0; ti < TM; ++ti)
          if (mine(ti)) dmma(kc[ti], sPH[fr.a8(ti, ks)], b);
      }
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
        if (mine(ti)) sts2(sK + fr.c8(ti), kc[ti]);
      // ---- innovation nu = y - H x- (182-184) and x+ = x- + K nu (186-195)
      const double innov = (g < m) ? (yv - hxm) : 0.0;  // valid in the t == 0 lane of quad g
      {
        const double i0 = __shfl_sync(0xffffffffu, innov, (2 * t) * 4);
        const double i1 = __shfl_sync(0xffffffffu, innov, (2 * t + 1) * 4);
#pragma unroll
        for (int ti = 0; ti < TM; ++ti) {
          if (!mine(ti)) continue;
          const double dx = quad_sum(fma(kc[ti][0], i0, kc[ti][1] * i1));
          if (t == 0) xs[ti * 8 + g] = xms[ti * 8 + g] + dx;
        }
      }
      if (HALF == 0 && t == 0) {
        sinn[g] = innov;
        sinn[kMP + g] = (g < m) ? yhat : 0.0;
      }
      psync();
      bad = bad || !(s_norm * sinv_norm <= 1e16);
      if (bad) {  // warp-uniform (and identical in both warps of a pair): pivots and norms are the same in every lane
        status = GKB_ERR_SINGULAR_S;
        break;
      }
      // ---- Joseph form (197-205), restructured: T2 = P- - K PHt^T (symmetric: upper tiles)
#pragma unroll
      for (int t0 = 0; t0 < TM; t0 += TB) {
        if (!mine(t0)) continue;
#pragma unroll
        for (int ti = 0; ti < TB; ++ti)
#pragma unroll
          for (int tj = t0 + ti; tj < TM; ++tj) {
            const double2 v = lds2(bufA + fr.template p_c<TM>(t0 + ti, tj));
            c[ti * TM + tj][0] = v.x;
            c[ti * TM + tj][1] = v.y;
          }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          double a[TB], b[TM];
#pragma unroll
          for (int ti = 0; ti < TB; ++ti) a[ti] = -sK[fr.a8(t0 + ti, ks)];
#pragma unroll
          for (int tj = t0; tj < TM; ++tj) b[tj] = sPH[fr.a8(tj, ks)];  // B[k][j] = PHt[j][k]
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = t0 + ti; tj < TM; ++tj) dmma(c[ti * TM + tj], a[ti], b[tj]);
        }
#pragma unroll
        for (int ti = 0; ti < TB; ++ti)
#pragma unroll
          for (int tj = t0 + ti; tj < TM; ++tj) {
            if constexpr (TREG) sts2(bufA + fr.template p_c<TM>(t0 + ti, tj), c[ti * TM + tj]);  // in place over P-
            else sts2(bufB + fr.c(t0 + ti, tj), c[ti * TM + tj]);
          }
      }
      psync();
      // ---- V = T2 H^T - K R
      double vc[TM][2];
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) vc[ti][0] = vc[ti][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double b = sH[fr.a(0, ks)];
#pragma unroll
        for (int ti = 0; ti < TM; ++ti)
          if (mine(ti)) dmma(vc[ti], TREG ? bufA[fr.template p_a<TM>(ti, ks)] : bufB[fr.sym_a(ti, ks)], b);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const double b = sR[fr.b8(ks)];
#pragma unroll
        for (int ti = 0; ti < TM; ++ti)
          if (mine(ti)) dmma(vc[ti], -sK[fr.a8(ti, ks)], b);
      }
      // V takes the place of PHt (every thread of the filter is past its last read of sPH: the barrier above)
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
        if (mine(ti)) sts2(sPH + fr.c8(ti), vc[ti]);
      psync();
      // ---- P+ = T2 - V K^T, upper tiles (n <= 32: c still holds T2; larger n: T2 comes back from bufB)
#pragma unroll
      for (int t0 = 0; t0 < TM; t0 += TB) {
        if (!mine(t0)) continue;
        if constexpr (TB < TM) {
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = t0 + ti; tj < TM; ++tj) {
              const double2 v = TREG ? lds2(bufA + fr.template p_c<TM>(t0 + ti, tj)) : lds2(bufB + fr.c(t0 + ti, tj));
              c[ti * TM + tj][0] = v.x;
              c[ti * TM + tj][1] = v.y;
            }
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          double a[TB], b[TM];
#pragma unroll
          for (int ti = 0; ti < TB; ++ti) a[ti] = -sPH[fr.a8(t0 + ti, ks)];
#pragma unroll
          for (int tj = t0; tj < TM; ++tj) b[tj] = sK[fr.a8(tj, ks)];  // B[k][j] = K[j][k]
#pragma unroll
          for (int ti = 0; ti < TB; ++ti)
#pragma unroll
            for (int tj = t0 + ti; tj < TM; ++tj) dmma(c[ti * TM + tj], a[ti], b[tj]);
        }
#pragma unroll
        for (int ti = 0; ti < TB; ++ti)
#pragma unroll
          for (int tj = t0 + ti; tj < TM; ++tj) sts2(bufA + fr.template p_c<TM>(t0 + ti, tj), c[ti * TM + tj]);
      }
      psync();
      // ---- Estimate fields of this step
      if (io.every_step || k == io.steps - 1) {
        const int64_t row = (io.every_step ? (int64_t)k * io.nf : 0) + f;
        if (io.o_state != nullptr)
          for (int idx = plane; idx < N; idx += PT) io.o_state[row * N + idx] = xs[idx];
        if (io.o_innov != nullptr && plane < m) io.o_innov[row * m + plane] = sinn[plane];
        if (io.o_meas != nullptr && plane < m) io.o_meas[row * m + plane] = sinn[kMP + plane];
        if (io.o_gain != nullptr)
          for (int idx = plane; idx < N * m; idx += PT) io.o_gain[row * (int64_t)(N * m) + idx] = sK[idx_small(idx / m, idx % m)];
        if (io.o_covar != nullptr) {
          double* dst = io.o_covar + row * (int64_t)(N * N);
          for (int idx = plane; idx < N * N; idx += PT) {
            const int r = idx / N, cc = idx % N;
            dst[idx] = bufA[idx_packed<TM>(r, cc)];
          }
        }
      }
    }
    // ---- state out (a failed filter keeps the state it had when the call started)
    if (status == 0) {
      bool finite = true;
      for (int idx = lane; idx < N; idx += 32) finite = finite && isfinite(xs[idx]) && isfinite(bufA[idx_packed<TM>(idx, idx)]);
      finite = __all_sync(0xffffffffu, finite);
      if (!finite) status = GKB_ERR_NONFINITE;
    }
    if (status == 0) {
      for (int idx = plane; idx < N; idx += PT) io.x[f * N + idx] = xs[idx];
      double* Pg = io.P + f * (int64_t)(N * N);
      for (int idx = plane; idx < N * N; idx += PT) {
        const int r = idx / N, cc = idx % N;
        Pg[idx] = bufA[idx_packed<TM>(r, cc)];
      }
    } else if (plane == 0 && io.status != nullptr && io.status[f] == 0) {
      io.status[f] = status;
    }
    psync();
  }
}

template <int N>
__global__ void __launch_bounds__(N <= 16 ? 512 : (N <= 32 ? 384 : 256), 1) vanilla_tile_kernel(const __grid_constant__ TileIo io) {
  using SH = TileShape<N>;
  constexpr int LDB = SH::LDB, PW = SH::PW;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slots = (int)(blockDim.x >> 5) / PW, slot = warp / PW;  // filters in flight in this CTA, this warp's
  // CTA-shared model (Q is only ever a C-fragment initialiser: read straight from global / L2 once per step)
  double* sF = smem;                // [N][LDB]
  double* sH = sF + N * LDB;        // [8][LDB]  rows >= m are zero
  double* sR = sH + kMP * LDB;      // [8][8]    padded with a unit diagonal
  double* fbase = sR + kMP * 8 + (size_t)slot * SH::kPerFilter;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) sF[idx_big<LDB>(idx / N, idx % N)] = io.F[idx];
  for (int idx = threadIdx.x; idx < kMP * N; idx += blockDim.x) sH[idx_big<LDB>(idx / N, idx % N)] = io.H[idx];
  for (int idx = threadIdx.x; idx < kMP * kMP; idx += blockDim.x) sR[idx_small(idx / kMP, idx % kMP)] = io.R[idx];
  __syncthreads();
  if constexpr (PW == 1) {
    tile_run<N, 0>(io, sF, sH, sR, fbase, slot, slots, lane);
  } else {
    if ((warp & 1) == 0) tile_run<N, 0>(io, sF, sH, sR, fbase, slot, slots, lane);
    else tile_run<N, 1>(io, sF, sH, sR, fbase, slot, slots, lane);
  }
}

template <int N>
static int launch_tile_shape(const TileIo& io, int device, cudaStream_t s) {
  using SH = TileShape<N>;
  constexpr int LDB = SH::LDB, PW = SH::PW;
  constexpr size_t kShared = sizeof(double) * (N * LDB + kMP * LDB + kMP * 8);
  constexpr size_t kPerWarp = sizeof(double) * SH::kPerFilter;  // per filter in flight (one warp, or a pair for n > 32)
  static thread_local int cached_device = -1, sms = 148, max_smem = 227 * 1024;
  if (cached_device != device) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaFuncSetAttribute(vanilla_tile_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cached_device = device;
  }
  int warps = (int)(((size_t)max_smem - kShared) / kPerWarp);
  constexpr int kMaxWarps = N <= 16 ? 16 : (N <= 32 ? 12 : 4);  // filters in flight; register budget: 64 K / (32 x registers per thread)
  if (warps > kMaxWarps) warps = kMaxWarps;
  if (warps < 1) return GKB_ERR_UNSUPPORTED;
  int64_t ctas = (io.nf + warps - 1) / warps;
  if (ctas > sms) ctas = sms;  // persistent: one CTA per SM, warps stride over the filters
  if (const char* e = getenv("GKB_TILE_WARPS")) {  // profiling knob: fewer warps per CTA (occupancy sweeps in DESIGN.md)
    const int w = atoi(e);
    if (w >= 1 && w <= warps) warps = w;
  }
  vanilla_tile_kernel<N><<<(unsigned)ctas, warps * PW * 32, kShared + kPerWarp * warps, s>>>(io);
  return 0;
}

__global__ void tile_gu_kernel(const double* __restrict__ G, int n, int c, const double* __restrict__ u, int steps,
                               double* __restrict__ gu) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= steps * n) return;
  const int k = idx / n, i = idx % n;
  double s = 0.0;
  for (int j = 0; j < c; ++j) s = fma(G[i * c + j], u[(size_t)k * c + j], s);
  gu[idx] = s;
}

int launch_tile_gu(const double* G_dev, int n, int c, const double* u_dev, int steps, double* gu_dev, cudaStream_t s) {
  const int total = steps * n;
  tile_gu_kernel<<<(total + 127) / 128, 128, 0, s>>>(G_dev, n, c, u_dev, steps, gu_dev);
  return 0;
}

int tile_shape_supported(int n, int m) { return n >= 16 && n <= 64 && n % 8 == 0 && m >= 1 && m <= kMP; }

int launch_tile_update(const TileIo& io, int n, int device, cudaStream_t s) {
  switch (n) {
    case 16: return launch_tile_shape<16>(io, device, s);
    case 24: return launch_tile_shape<24>(io, device, s);
    case 32: return launch_tile_shape<32>(io, device, s);
    case 40: return launch_tile_shape<40>(io, device, s);
    case 48: return launch_tile_shape<48>(io, device, s);
    case 56: return launch_tile_shape<56>(io, device, s);
    case 64: return launch_tile_shape<64>(io, device, s);
    default: return GKB_ERR_UNSUPPORTED;
  }
}

}  // namespace gkb
