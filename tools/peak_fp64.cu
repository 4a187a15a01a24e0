// FP64 roofline denominators for B200 (sm_100a): sustained DFMA and DMMA (mma.sync m8n8k4 f64)
// throughput, measured with CUDA events.  MEASURED_PEAKS.json has no FP64 entry,
// so bench.py runs this probe on the box and uses its DFMA figure as the FP64 peak.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_fp64 tools/peak_fp64.cu
// Output: one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[2]), "+d"(c0[3]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c1[2]), "+d"(c1[3]) : "d"(a), "d"(b));
    }
  }
  double s = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 20000;
  const int ILP = 8;
  double best_dfma = 0, best_dmma = 0, sustained_dfma = 0;
  // burst: best of 10 short launches
  for (int rep = 0; rep < 12; ++rep) {
    CK(cudaEventRecord(e0));
    dfma_kernel<ILP><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * ILP * 8 * (double)iters * sms * 8 * 256;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best_dfma) best_dfma = tf;
  }
  // sustained: ~3 s of back-to-back launches
  {
    int n = 0;
    CK(cudaEventRecord(e0));
    for (; n < 150; ++n) dfma_kernel<ILP><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * ILP * 8 * (double)iters * sms * 8 * 256 * n;
    sustained_dfma = flops / (ms * 1e-3) / 1e12;
  }
  for (int rep = 0; rep < 12; ++rep) {
    CK(cudaEventRecord(e0));
    dmma_kernel<<<sms * 8, 256>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // m8n8k4: 8*8*4 MAC = 512 flop per warp-instruction; 4*8 instr per iter per warp
    double flops = 512.0 * 32 * (double)iters * (sms * 8 * 256 / 32);
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best_dmma) best_dmma = tf;
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, \"dfma_tflops_burst\": %.3f, "
         "\"dfma_tflops_sustained\": %.3f, \"dmma_m8n8k4_tflops_burst\": %.3f, "
         "\"how\": \"dfma: %d blocks x 256 thr, ILP %d, fma chain, best of 10 (burst) / 150 back-to-back launches (sustained); "
         "dmma: mma.sync.m8n8k4.f64, 4 independent accumulators\"}\n",
         p.name, sms, clk_khz / 1000.0, best_dfma, sustained_dfma, best_dmma, sms * 8, ILP);
  return 0;
}
