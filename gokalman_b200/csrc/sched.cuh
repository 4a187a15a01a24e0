// sched.cuh -- the (chunk of epochs, group of 32 filters) task scheduler shared by the one-filter-per-thread run kernels
// that are not on the TMA production path (the strict hybrid kernel, the fused OD run).
//
// More groups than resident warps leaves a tail when every warp owns whole groups (10^5 filters = 3125 groups over 1184
// resident warps: 3 rounds for 2.64 rounds of work).  Persistent warps instead claim TASKS from one atomic counter in
// chunk-major order (all groups' chunk 0, then chunk 1, ...); a group's state travels from the warp that ran chunk c to
// whichever warp claims chunk c + 1 through the handle's state arrays (L2) with one release / acquire flag per group.
// Chunk c + 1 of a group is claimed a whole pass over all groups after chunk c, so the acquire practically never spins.
// NlIo::sched: [0] task counter, [1 + g] chunks of group g already written back (zeroed by the launcher).
#pragma once
#include <cstdlib>

#include "engine_internal.h"

namespace gkb {

// Spin (lane 0) until the group's flag reaches `value`, then make the state written by the group's previous owner visible.
__device__ __forceinline__ void sched_acquire_group(const int* flag, int value, int lane) {
  if (lane == 0) {
    int seen;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
      if (seen < value) __nanosleep(200);
    } while (seen < value);
  }
  __syncwarp();
  __threadfence();
}
// Publish the group's state (written by every lane of the warp) to whichever warp claims its next chunk.
__device__ __forceinline__ void sched_release_group(int* flag, int value, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// Claim the next task (warp-uniform); false when none is left.
__device__ __forceinline__ bool sched_claim(int* counter, int n_tasks, int groups, int lane, int& chunk, int& group) {
  int task = 0;
  if (lane == 0) task = atomicAdd(counter, 1);
  task = __shfl_sync(0xffffffffu, task, 0);
  if (task >= n_tasks) return false;
  chunk = task / groups;
  group = task - chunk * groups;
  return true;
}

// Host side: chunk count for `groups` groups over `slots` resident warps -- about `tasks_per_warp` tasks per resident warp
// (finer tasks even out warps that run at different speeds), among the neighbouring counts the one that fills whole
// rounds best, chunks of at least 16 epochs; 1 = do not schedule.  GKB_NL_CHUNKS=c forces c (tests, sweeps).
inline int sched_pick_chunks(int64_t groups, int64_t slots, int steps, double tasks_per_warp, bool* forced) {
  int chunks = 1;
  if (groups > slots) {
    const double per_slot = (double)groups / (double)slots;
    const int base = (int)(tasks_per_warp / per_slot + 0.5);
    double best = -1.0;
    for (int c = base - 1; c <= base + 1; ++c) {
      if (c < 2 || steps / c < 16) continue;
      const double rounds = per_slot * c;
      const double eff = rounds / (double)(int64_t)(rounds + 0.999999);
      if (eff > best + 1e-9) { best = eff; chunks = c; }
    }
  }
  *forced = false;
  if (const char* e = getenv("GKB_NL_CHUNKS")) {
    const int c = atoi(e);
    if (c >= 1 && c <= steps) { chunks = c; *forced = true; }
  }
  return chunks;
}
// Fills NlIo::chunks / chunk_len for `chunks` chunks (0 chunks = unscheduled launch).
inline void sched_set_chunks(NlIo& io, int chunks, bool scheduled) {
  io.chunks = 0;
  io.chunk_len = io.steps;
  if (!scheduled) return;
  io.chunk_len = (io.steps + chunks - 1) / chunks;
  io.chunks = (io.steps + io.chunk_len - 1) / io.chunk_len;
}

}  // namespace gkb
