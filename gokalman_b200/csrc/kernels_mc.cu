// kernels_mc.cu -- instantiation and dispatch of the fused Monte Carlo / chi-square kernel
// (kernels_mc.cuh) over the compiled (n, m) shapes and the three tested LDKF kinds.
#include "kernels_mc.cuh"

namespace gkb {

#ifndef GKB_MC_PART
#error "compile with -DGKB_MC_PART=0..12 (see Makefile): 0 = dispatch + finish kernel; 1 + 3 g + (kind - 1) = the kernels of tested kind 1 / 2 / 3 (vanilla / information / sqrt) for shape group g = 0 (n <= 4), 1 (n = 5, 6), 2 (n = 7), 3 (n = 8)"
#endif
// The fused kernels are the heaviest templates of the library (the n = 8 information filter alone takes minutes):
// twelve parts build in parallel.  Groups 2 and 3 are the north star's "n <= 8": spilled, slower, same parity bar.
#if GKB_MC_PART > 0
#define GKB_MC_KIND ((GKB_MC_PART - 1) % 3 + 1)
#define GKB_MC_GROUP ((GKB_MC_PART - 1) / 3)
#if GKB_MC_GROUP == 0
#define GKB_MC_SHAPES(X) GKB_FOR_EACH_SHAPE_G0(X)
#define GKB_MC_NAME(base) base##_g0
#elif GKB_MC_GROUP == 1
#define GKB_MC_SHAPES(X) GKB_FOR_EACH_SHAPE_G1(X)
#define GKB_MC_NAME(base) base##_g1
#elif GKB_MC_GROUP == 2
#define GKB_MC_SHAPES(X) GKB_FOR_EACH_SHAPE_G2(X)
#define GKB_MC_NAME(base) base##_g2
#else
#define GKB_MC_SHAPES(X) GKB_FOR_EACH_SHAPE_G3(X)
#define GKB_MC_NAME(base) base##_g3
#endif
#else
#define GKB_MC_KIND 0
#endif
#if GKB_MC_PART == 0
// out[col][k] = scale * sum_b partial[b][k][col].  A CTA is 32 result slots x 8 row groups: thread (x, y) adds the CTA
// rows y, y + 8, ... of slot x in order, then the 8 group sums are added in group order -- a fixed order, so the
// result is bit-reproducible for a given grid, with 8x the memory parallelism of one thread per slot.
__global__ void __launch_bounds__(256) mc_finish_kernel(const double* __restrict__ partial, int grid, int steps, int cols,
                                                        double scale, double* __restrict__ out) {
  __shared__ double part[8][33];
  const int idx = blockIdx.x * 32 + threadIdx.x;
  const int total = steps * cols;
  double s = 0.0;
  if (idx < total)
    for (int b = threadIdx.y; b < grid; b += 8) s += partial[(size_t)b * total + idx];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || idx >= total) return;
#pragma unroll
  for (int g = 1; g < 8; ++g) s += part[g][threadIdx.x];
  const int k = idx / cols, col = idx % cols;
  out[(size_t)col * steps + k] = (col < kMcBaseCols) ? s * scale : s;  // only NIS / NEES become means
}
#endif  // GKB_MC_PART == 0

// mm: the truth generator (tm: F, G, H; io: chol(Q), chol(R), x0) plus the tested filter's initial estimate
template <int N, int M>
static void fill_mc_model(const HostModel& tm, const McIo& io, McModel<N, M>& mm, const double* A0) {
  for (int i = 0; i < N * N; ++i) { mm.F[i] = tm.F[i]; mm.LQ[i] = io.LQ[i]; mm.A0[i] = A0[i]; }
  for (int i = 0; i < M * M; ++i) mm.LR[i] = io.LR[i];
  for (int i = 0; i < N * GKB_MAX_C; ++i) mm.G[i] = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < tm.c; ++j) mm.G[i * tm.c + j] = tm.G[i * tm.c + j];
  for (int i = 0; i < M * N; ++i) mm.H[i] = tm.H[i];
  for (int i = 0; i < N; ++i) { mm.x0_truth[i] = io.x0_truth[i]; mm.x0_filter[i] = io.x0_filter[i]; }
  mm.c = tm.c;
  mm.need_ctrl = tm.need_ctrl;
}
// the tested filter's own G (only its shape and need_ctrl matter to the kernels: G u arrives precomputed in io.gu_f)
template <class Model, int N>
static void fill_tested_ctrl(const HostModel& hm, Model& md) {
  for (int i = 0; i < N * GKB_MAX_C; ++i) md.G[i] = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < hm.c; ++j) md.G[i * hm.c + j] = hm.G[i * hm.c + j];
  md.c = hm.c;
  md.need_ctrl = hm.need_ctrl;
}

template <class Kern>
static int pick_grid(Kern kern, size_t smem, int64_t trials, int device) {
  // per (kernel instantiation, shared-memory size): set the attribute and query occupancy once
  static thread_local Kern cached_kern = nullptr;
  static thread_local size_t cached_smem = (size_t)-1;
  static thread_local int cached_device = -1, cached_sms = 148, cached_per_sm = 1;
  if (cached_kern != kern || cached_smem != smem || cached_device != device) {
    cached_kern = kern;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_per_sm, kern, kThreads, smem);
    cached_smem = smem;
    cached_device = device;
  }
  int sms = cached_sms, per_sm = cached_per_sm;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > kMcMaxCtasPerSm) per_sm = kMcMaxCtasPerSm;
  // persistent CTAs: SM count x resident CTAs, no more than the work needs
  int64_t need = (trials + kThreads - 1) / kThreads;
  int64_t g = (int64_t)sms * per_sm;
  if (g > need) g = need;
  return (int)(g < 1 ? 1 : g);
}

#if GKB_MC_KIND == 1
template <int N, int M>
static int launch_mc_shape_vanilla(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
  const int cols = mc_cols(N, io.want_xstats);
  const size_t smem = sizeof(double) * (kIcdfSegments * kIcdfCoefs + kWarps * kChunk * cols);
  int grid = 1;
  const bool lean = io.noise_mode == GKB_NOISE_PHILOX && !io.want_xstats && !io.truth_x && !io.truth_y &&
                    !io.noise_w && !io.noise_v && !io.status && io.gu_f == io.gu;
  McModel<N, M> mm;
  fill_mc_model<N, M>(tm, io, mm, io.P0);
  {
      VanillaModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.Q[i] = hm.Q[i]; }
      for (int i = 0; i < M * M; ++i) md.R[i] = hm.R[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      fill_tested_ctrl<decltype(md), N>(hm, md);
      if (lean && io.gu == nullptr) {
        auto kern = mc_chisquare_kernel<N, M, VanillaTested<N, M>, true, true>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      } else if (lean) {
        auto kern = mc_chisquare_kernel<N, M, VanillaTested<N, M>, true>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      } else {
        auto kern = mc_chisquare_kernel<N, M, VanillaTested<N, M>, false>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      }
      return 0;
      }
}

int GKB_MC_NAME(launch_mc_vanilla)(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_mc_shape_vanilla<NN, MM>(tm, hm, io, device, grid_out, s);
  GKB_MC_SHAPES(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}
#endif

#if GKB_MC_KIND == 2
template <int N, int M>
static int launch_mc_shape_info(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
  const int cols = mc_cols(N, io.want_xstats);
  const size_t smem = sizeof(double) * (kIcdfSegments * kIcdfCoefs + kWarps * kChunk * cols);
  int grid = 1;
  const bool lean = io.noise_mode == GKB_NOISE_PHILOX && !io.want_xstats && !io.truth_x && !io.truth_y &&
                    !io.noise_w && !io.noise_v && !io.status && io.gu_f == io.gu;
  McModel<N, M> mm;
  fill_mc_model<N, M>(tm, io, mm, io.P0);
  {
      InfoModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.Finv[i] = hm.Finv[i]; md.Qinv[i] = hm.Qinv[i]; }
      for (int i = 0; i < M * M; ++i) md.Rinv[i] = hm.Rinv[i];
      md.rinv_dim = hm.rinv_dim;
      for (int i = 0; i < M * M; ++i) md.R[i] = hm.R[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      fill_tested_ctrl<decltype(md), N>(hm, md);
      if (lean) {
        auto kern = mc_chisquare_kernel<N, M, InfoTested<N, M>, true>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      } else {
        auto kern = mc_chisquare_kernel<N, M, InfoTested<N, M>, false>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      }
      return 0;
      }
}

int GKB_MC_NAME(launch_mc_info)(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_mc_shape_info<NN, MM>(tm, hm, io, device, grid_out, s);
  GKB_MC_SHAPES(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}
#endif

#if GKB_MC_KIND == 3
template <int N, int M>
static int launch_mc_shape_sqrt(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
  const int cols = mc_cols(N, io.want_xstats);
  const size_t smem = sizeof(double) * (kIcdfSegments * kIcdfCoefs + kWarps * kChunk * cols);
  int grid = 1;
  const bool lean = io.noise_mode == GKB_NOISE_PHILOX && !io.want_xstats && !io.truth_x && !io.truth_y &&
                    !io.noise_w && !io.noise_v && !io.status && io.gu_f == io.gu;
  McModel<N, M> mm;
  fill_mc_model<N, M>(tm, io, mm, io.P0);
  {
      SqrtModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.sqrtQ[i] = hm.sqrtQ[i]; }
      for (int i = 0; i < M * M; ++i) md.sqrtR[i] = hm.sqrtR[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      fill_tested_ctrl<decltype(md), N>(hm, md);
      if (lean) {
        auto kern = mc_chisquare_kernel<N, M, SqrtTested<N, M>, true>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      } else {
        auto kern = mc_chisquare_kernel<N, M, SqrtTested<N, M>, false>;
        *grid_out = grid = pick_grid(kern, smem, io.trials, device);
        kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      }
      return 0;
      }
}

int GKB_MC_NAME(launch_mc_sqrt)(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_mc_shape_sqrt<NN, MM>(tm, hm, io, device, grid_out, s);
  GKB_MC_SHAPES(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}
#endif

#if GKB_MC_PART == 0
int mc_max_grid(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * kMcMaxCtasPerSm;
}

#define GKB_MC_DECL(g) \
  int launch_mc_vanilla_g##g(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s); \
  int launch_mc_info_g##g(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s);    \
  int launch_mc_sqrt_g##g(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s);
GKB_MC_DECL(0) GKB_MC_DECL(1) GKB_MC_DECL(2) GKB_MC_DECL(3)
#undef GKB_MC_DECL

int launch_mc(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
  const int g = hm.n <= 4 ? 0 : hm.n <= 6 ? 1 : hm.n == 7 ? 2 : 3;
#define GKB_MC_CALL(fn) \
  (g == 0 ? fn##_g0(tm, hm, io, device, grid_out, s) : g == 1 ? fn##_g1(tm, hm, io, device, grid_out, s) : \
   g == 2 ? fn##_g2(tm, hm, io, device, grid_out, s) : fn##_g3(tm, hm, io, device, grid_out, s))
  switch (hm.kind) {
    case GKB_VANILLA: return GKB_MC_CALL(launch_mc_vanilla);
    case GKB_INFORMATION: return GKB_MC_CALL(launch_mc_info);
    case GKB_SQRT: return GKB_MC_CALL(launch_mc_sqrt);
    default: return GKB_ERR_UNSUPPORTED;
  }
#undef GKB_MC_CALL
}

int mc_shape_supported(int kind, int n, int m) {
  if (kind != GKB_VANILLA && kind != GKB_INFORMATION && kind != GKB_SQRT) return 0;
#define GKB_CASE(NN, MM) \
  if (n == NN && m == MM) return 1;
  GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)  // n <= 6 (parts 1-3) and n = 7, 8 (parts 4-6)
#undef GKB_CASE
  return 0;
}

int launch_mc_finish(const double* partial, int grid, int steps, int cols, double scale, double* out_cols,
                     cudaStream_t s) {
  const int total = steps * cols;
  mc_finish_kernel<<<(total + 31) / 32, dim3(32, 8), 0, s>>>(partial, grid, steps, cols, scale, out_cols);
  return 0;
}

#endif  // GKB_MC_PART == 0

}  // namespace gkb
