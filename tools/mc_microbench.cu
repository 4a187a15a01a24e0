// Standalone timing of the LEAN <3,1,Vanilla> Monte Carlo kernel for kernel-tuning experiments:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -o /tmp/mcb tools/mc_microbench.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include "../gokalman_b200/csrc/kernels_mc.cuh"
using namespace gkb;
#ifndef MB_CTAS_PER_SM
#define MB_CTAS_PER_SM 0
#endif
int main(int argc, char** argv) {
  const int64_t trials = argc > 1 ? atoll(argv[1]) : 1000000;
  const int steps = argc > 2 ? atoi(argv[2]) : 1000;
  constexpr int N = 3, M = 1;
  McModel<N, M> mm; VanillaModel<N, M> md; McIo io;
  memset(&mm, 0, sizeof mm); memset(&md, 0, sizeof md); memset(&io, 0, sizeof io);
  double F[9] = {1, 0.01, 5e-5, 0, 1, 0.01, 0, 0, 1}, G[3] = {5e-7 / 3, 5e-5, 0.01}, H[3] = {1, 0, 0};
  double Q[9] = {2.5e-15, 6.25e-13, 25e-11 / 3, 6.25e-13, 5e-7 / 3, 2.5e-8, 25e-11 / 3, 2.5e-8, 5e-6};
  double LQ[9] = {0}, x0[3] = {0, 0.35, 0};
  // chol(Q)
  for (int j = 0; j < 3; ++j) { double a = Q[j*3+j]; for (int l = 0; l < j; ++l) a -= LQ[j*3+l]*LQ[j*3+l]; a = sqrt(a); LQ[j*3+j] = a;
    for (int i = j+1; i < 3; ++i) { double s = Q[j*3+i]; for (int l = 0; l < j; ++l) s -= LQ[j*3+l]*LQ[i*3+l]; LQ[i*3+j] = s / a; } }
  for (int i = 0; i < 9; ++i) { mm.F[i] = md.F[i] = F[i]; md.Q[i] = Q[i]; mm.LQ[i] = LQ[i]; mm.A0[i] = (i % 4 == 0) ? 10.0 : 0.0; }
  for (int i = 0; i < 3; ++i) { mm.G[i] = md.G[i] = G[i]; mm.H[i] = md.H[i] = H[i]; mm.x0_truth[i] = mm.x0_filter[i] = x0[i]; }
  md.R[0] = 0.5; mm.LR[0] = sqrt(0.5); mm.c = md.c = 1; mm.need_ctrl = md.need_ctrl = 1;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto kern = mc_chisquare_kernel<N, M, VanillaTested<N, M>, true>;
  const int cols = 2; const size_t smem = sizeof(double) * (kIcdfSegments * kIcdfCoefs + kWarps * kChunk * cols);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (MB_CTAS_PER_SM > 0 && per_sm > MB_CTAS_PER_SM) per_sm = MB_CTAS_PER_SM;
  int grid = sms * per_sm; int64_t need = (trials + kThreads - 1) / kThreads; if (grid > need) grid = (int)need;
  double *partial, *u; cudaMalloc(&partial, sizeof(double) * (size_t)grid * steps * cols); cudaMalloc(&u, sizeof(double) * steps);
  cudaMemset(u, 0, sizeof(double) * steps);
  io.trials = trials; io.steps = steps; io.noise_mode = GKB_NOISE_PHILOX; io.seed = 0x5EED; io.with_nees = io.with_nis = 1; io.partial = partial;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemset(partial, 0, sizeof(double) * (size_t)grid * steps * cols);
    cudaEventRecord(e0); kern<<<grid, kThreads, smem>>>(mm, md, io); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  std::vector<double> h((size_t)grid * steps * cols); cudaMemcpy(h.data(), partial, h.size() * 8, cudaMemcpyDeviceToHost);
  double nis = 0, nees = 0; for (int b = 0; b < grid; ++b) for (int k = 0; k < steps; ++k) { nis += h[((size_t)b * steps + k) * 2]; nees += h[((size_t)b * steps + k) * 2 + 1]; }
  cudaError_t err = cudaGetLastError();
  printf("regs=%d ctas/sm=%d grid=%d  best %.3f ms  %.3e units/s  %.2f TF(537)  nis=%.9f nees=%.9f  %s\n", fa.numRegs, per_sm, grid, best,
         (double)trials * steps / (best * 1e-3), 537.0 * trials * steps / (best * 1e-3) / 1e12, nis / trials / steps, nees / trials / steps, cudaGetErrorString(err));
  return 0;
}
