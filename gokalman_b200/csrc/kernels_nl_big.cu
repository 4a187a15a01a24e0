// kernels_nl_big.cu -- HybridKF / SRIF at n = 7 and 8 (the north star's "n <= 8"): the general one-filter-per-thread
// kernels of kernels_nl.cuh instantiated for the larger shapes.  The 8 x 8 intermediates no longer fit the register
// file (ptxas reports local memory: L1-resident spills), so these run slower than the n <= 6 shapes -- correct to the
// same parity bar; the TMA production path stays n <= 6.
#include "kernels_nl.cuh"

namespace gkb {

int launch_nl_run_big(const HostModel& hm, const NlIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_nl_general<NN, MM>(hm, io, s);
  GKB_FOR_EACH_BIG_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

}  // namespace gkb
