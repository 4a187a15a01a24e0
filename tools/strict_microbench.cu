// Standalone timing of the reference-order ("strict") hybrid kernel <6,2> for kernel-tuning experiments (variants are
// selected with -D macros read by filters_strict.cuh / kernels_nl.cuh; every variant must print the same checksum):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -o tools/sb_base tools/strict_microbench.cu
// usage: sb_xxx [filters] [epochs] [9 = the register version]   (GKB_NL_CHUNKS=c forces the chunk count)
#include <cstdio>
#include <cstring>
#include <vector>
#include <random>
#include "../gokalman_b200/csrc/kernels_nl.cuh"
using namespace gkb;

__global__ void synth_kernel(int64_t nf, int steps, double* Phi, double* Ht, double* ro, double* co) {
  // near-identity transition matrices and O(1) measurement partials, different per (filter, epoch, entry)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nf) return;
  for (int k = 0; k < steps; ++k) {
    unsigned long long s = (unsigned long long)t * 0x9E3779B97F4A7C15ull + (unsigned long long)k * 0xBF58476D1CE4E5B9ull + 1;
    auto rnd = [&]() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return (double)((s * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) - 0.5; };
    for (int e = 0; e < 36; ++e) Phi[((int64_t)k * 36 + e) * nf + t] = ((e % 7 == 0) ? 1.0 : 0.0) + 0.02 * rnd();
    for (int e = 0; e < 12; ++e) Ht[((int64_t)k * 12 + e) * nf + t] = rnd();
    for (int e = 0; e < 2; ++e) { ro[((int64_t)k * 2 + e) * nf + t] = rnd(); co[((int64_t)k * 2 + e) * nf + t] = 0.1 * rnd(); }
  }
}

int main(int argc, char** argv) {
  const int64_t nf = argc > 1 ? atoll(argv[1]) : 100000;
  const int steps = argc > 2 ? atoi(argv[2]) : 200;
  const int pf = argc > 3 ? atoi(argv[3]) : 0;
  constexpr int N = 6, M = 2;
  NlModel<N, M> md; NlIo io;
  memset(&md, 0, sizeof md); memset(&io, 0, sizeof io);
  md.R[0] = md.R[3] = 1e-2; md.L[0] = md.L[3] = 0.1; md.q = 3;
  double *Phi, *Ht, *ro, *co, *vec, *mat, *ox, *oP; int32_t* status; uint8_t* flags;
  cudaMalloc(&Phi, 8ull * 36 * nf * steps); cudaMalloc(&Ht, 8ull * 12 * nf * steps);
  cudaMalloc(&ro, 8ull * 2 * nf * steps); cudaMalloc(&co, 8ull * 2 * nf * steps);
  cudaMalloc(&vec, 8ull * N * nf); cudaMalloc(&mat, 8ull * N * N * nf); cudaMalloc(&ox, 8ull * N * nf); cudaMalloc(&oP, 8ull * N * N * nf);
  cudaMalloc(&status, 4ull * nf); cudaMalloc(&flags, steps);
  synth_kernel<<<(unsigned)((nf + 127) / 128), 128>>>(nf, steps, Phi, Ht, ro, co);
  std::vector<uint8_t> hf(steps);
  for (int k = 0; k < steps; ++k) hf[k] = (uint8_t)(GKB_F_MEAS | (k >= 15 ? GKB_F_EKF : 0));
  cudaMemcpy(flags, hf.data(), steps, cudaMemcpyHostToDevice);
  std::vector<double> hx((size_t)N * nf, 0.0), hP((size_t)N * N * nf, 0.0);
  for (int64_t t = 0; t < nf; ++t) for (int i = 0; i < N; ++i) hP[(size_t)(i * N + i) * nf + t] = i < 3 ? 10.0 : 1.0;
  io.nf = nf; io.steps = steps; io.vec = vec; io.mat = mat; io.flags = flags; io.Phi = Phi; io.Htilde = Ht; io.real_obs = ro;
  io.computed_obs = co; io.o_state = ox; io.o_covar = oP; io.status = status; io.strict = 1;
  int* sched; cudaMalloc(&sched, 4ull * ((nf + 31) / 32 + 1)); io.sched = sched;
  HostModel hm; memset(&hm, 0, sizeof hm);
  hm.kind = GKB_HYBRID; hm.n = N; hm.m = M; hm.q = 3; hm.R[0] = hm.R[3] = 1e-2; hm.L[0] = hm.L[3] = 0.1;
  if (pf == 9) setenv("GKB_STRICT_PATH", "regs", 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f; int regs = 0;
  { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, hybrid_run_strict_kernel<N, M>); regs = fa.numRegs; }
  for (int rep = 0; rep < 5; ++rep) {
    cudaMemcpy(vec, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(mat, hP.data(), hP.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(status, 0, 4ull * nf);
    cudaEventRecord(e0);
    launch_nl_general<N, M>(hm, io, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  cudaMemcpy(hx.data(), ox, hx.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hP.data(), oP, hP.size() * 8, cudaMemcpyDeviceToHost);
  std::vector<int32_t> hs(nf); cudaMemcpy(hs.data(), status, 4ull * nf, cudaMemcpyDeviceToHost);
  unsigned long long ck = 0; int64_t bad = 0;
  for (double v : hx) { unsigned long long b; memcpy(&b, &v, 8); if (v == 0.0) b = 0; ck = ck * 1099511628211ull + b; }
  for (double v : hP) { unsigned long long b; memcpy(&b, &v, 8); if (v == 0.0) b = 0; ck = ck * 1099511628211ull + b; }
  for (int32_t v : hs) bad += v != 0;
  printf("pf=%d regs=%d  best %.3f ms  %.3e updates/s  checksum %016llx  failed %lld  P[0]=%.6e x[0]=%.6e  %s\n", pf, regs, best,
         (double)nf * steps / (best * 1e-3), ck, (long long)bad, hP[0], hx[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
