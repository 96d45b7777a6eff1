// grid_sync.cuh — grid-wide barrier + deterministic all-reduce for cooperative (co-resident) launches.
//
// torchdiffeq's adaptive controller needs ONE error norm over the whole batch per attempted step, so
// the fused dopri5 kernel pays a grid-wide reduction per attempt; its latency is on the critical path of a solve
// that is otherwise ~1 us per stage.  A counter barrier (fence + atomic + acquire-poll) followed by a re-read of
// per-CTA partials measured ~10k cycles per attempt on B200.  This version is an all-gather of tagged words:
//
//   * every CTA publishes ONE 64-bit word per value: {fp32 partial | epoch << 32}, a single st.relaxed.gpu (8-byte
//     naturally aligned scalar store = single-copy atomic, so data and flag can never be seen torn or out of order —
//     no fence, no atomic);
//   * warp 0 of every CTA polls all gridDim.x words (ld.relaxed.gpu, all lanes' loads in flight together) until each
//     carries the current epoch, adds the payloads in CTA order in fp64, and broadcasts through shared memory;
//   * slots are double-buffered by epoch parity: a CTA can only publish epoch e+2 into the buffer of epoch e after it
//     completed epoch e+1, which needs every CTA's e+1 word, which each CTA writes only after it finished reading
//     epoch e.  So no word is overwritten while someone still waits for it.
//
// Every CTA adds the same numbers in the same order => all threads of the grid get bit-identical totals, and the
// accept/reject branch that follows is uniform without any further broadcast.  Cost: one L2 store + ~1-2 L2 load
// round trips + two __syncthreads.
//
// The host must zero the slot array before the launch (epoch 0 = "nothing published") and launch cooperatively.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gode {

constexpr int kGsMaxVals = 4;

struct GridSyncWs {
  unsigned long long* slots;  // [2][kGsMaxVals][gridDim.x]
};

__host__ __device__ inline size_t grid_sync_bytes(int grid) { return sizeof(unsigned long long) * 2 * kGsMaxVals * (size_t)grid; }

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Sum NV per-thread floats over the whole grid.  `v` enters as this thread's contribution (already masked for padding
// lanes) and leaves as the grid total, as double, identical in every thread.  WARPS = warps per CTA.
// s_f: shared float[WARPS * NV]; s_d: shared double[NV].
template <int NV, int WARPS>
__device__ __forceinline__ void grid_allreduce_sum(double (&v)[NV], float* s_f, double* s_d, const GridSyncWs& ws,
                                                   unsigned int& epoch, int lane, int warp) {
  static_assert(NV <= kGsMaxVals, "raise kGsMaxVals");
  ++epoch;
  const int grid = gridDim.x;
  unsigned long long* buf = ws.slots + (size_t)(epoch & 1u) * kGsMaxVals * grid;
  // 1. warp partials (fp32 is ample: the result is rounded to fp32 anyway; CTA sums are combined in fp64)
  float w[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    w[k] = (float)v[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) s_f[warp * NV + k] = w[k];
  }
  __syncthreads();
  if (warp == 0) {
    // 2. publish this CTA's words
    if (lane < NV) {
      float s = s_f[lane];
      for (int q = 1; q < WARPS; ++q) s += s_f[q * NV + lane];
      st_relaxed_u64(buf + (size_t)lane * grid + blockIdx.x,
                     (unsigned long long)__float_as_uint(s) | ((unsigned long long)epoch << 32));
    }
    // 3. gather everyone's words: lane c owns CTAs c, c+32, ...
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      const unsigned long long* col = buf + (size_t)k * grid;
      for (int c0 = 0; c0 < grid; c0 += 32 * 4) {
        unsigned long long x[4];
        bool need[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          need[u] = c < grid;
          x[u] = need[u] ? ld_relaxed_u64(col + c) : ((unsigned long long)epoch << 32);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          while (need[u] && (unsigned int)(x[u] >> 32) != epoch) x[u] = ld_relaxed_u64(col + c);
          s += (double)__uint_as_float((unsigned int)x[u]);
        }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) s_d[k] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_d[k];
}

// Plain grid barrier with the same mechanism (no payload).  Also orders prior global writes of every CTA before later
// global reads (L2 path: __ldcg) of any CTA: bar.sync, then thread 0 fences and publishes the flag word; the polling
// warp fences after observing all flags, then bar.sync releases the rest of the CTA.
template <int WARPS>
__device__ __forceinline__ void grid_barrier(const GridSyncWs& ws, unsigned int& epoch, int lane, int warp) {
  ++epoch;
  const int grid = gridDim.x;
  unsigned long long* col = ws.slots + (size_t)(epoch & 1u) * kGsMaxVals * grid;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      __threadfence();  // release: cumulative over the CTA's writes ordered before the bar.sync above
      st_relaxed_u64(col + blockIdx.x, (unsigned long long)epoch << 32);
    }
    for (int c0 = 0; c0 < grid; c0 += 32 * 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        if (c < grid)
          while ((unsigned int)(ld_relaxed_u64(col + c) >> 32) != epoch) {}
      }
    }
    __threadfence();
  }
  __syncthreads();
}

}  // namespace gode
