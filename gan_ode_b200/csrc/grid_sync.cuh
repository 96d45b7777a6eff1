// grid_sync.cuh — grid-wide barrier + deterministic all-reduce for cooperative (co-resident) launches.
//
// torchdiffeq's adaptive controller needs ONE error norm over the whole batch per attempted step, so
// the fused dopri5 kernel pays a grid-wide reduction per attempt, and its latency sits on the critical path of a
// solve whose stages are ~0.3 us each.  Measured on B200 (scripts/sync_bench.cu, cycles per all-reduce, SM clock):
//
//     CTAs                              16     32     64    128    256
//     cooperative_groups grid.sync    4116   4031   4713   4206   4337    (+ re-read of per-CTA partials)
//     counter (red.release + poll 1)  3542   3516   3528   4179   4407    <- used for > 16 CTAs
//     tagged-word all-gather          1745   1817   3635   7641  11634    <- used for <= 16 CTAs (in the real kernel 32
//                                                                          polling CTAs were slower than the counter)
//     16-CTA clusters + DSMEM         2637   3326   3333   3585   4275    (not worth the launch constraints)
//
//   * tagged-word all-gather: every CTA publishes ONE 64-bit word per value, {fp32 partial | epoch << 32}, with a
//     single st.relaxed.gpu (8-byte naturally aligned scalar store = single-copy atomic: data and flag cannot be
//     seen torn or reordered, so no fence and no atomic); warp 0 of every CTA polls all words.  Two L2 round trips,
//     but the polling is O(CTAs^2) and collapses beyond ~48 CTAs.
//   * counter: the same tagged words, plus a relaxed red on one counter that one lane polls as a HINT that everyone
//     has published; then the tagged words are read with all loads in flight (re-polling any that lag).  No
//     release/acquire anywhere, so arrivals do not wait for the CTA's outstanding output stores.
//
// Both are double-buffered by epoch parity: a CTA can only publish epoch e+2 into the buffer of epoch e after it
// completed epoch e+1, which needs every CTA's arrival at e+1, which each CTA signals only after it finished reading
// epoch e — so nothing is overwritten while someone still needs it.  Every CTA adds the same numbers in the same
// (CTA-index) order in fp64, so all threads of the grid get bit-identical totals and the accept/reject branch that
// follows is uniform without any further broadcast.
//
// PERSISTENT workspace (round 2): no memset per launch.  The first kSyncRegionBytes of every workspace are the sync region
// [counter | epoch base | counter base | tagged-word slots]; the caller zero-fills it ONCE after allocation and then hands the
// same workspace to any number of launches that are ordered on one stream.  Tags keep counting ACROSS launches: every thread
// reads (epoch base, counter base) when the kernel starts, and thread 0 of CTA 0 stores the final values after the kernel's
// last grid-wide operation (by then every CTA has read them: it has arrived at that operation).  A stale word therefore always
// carries an OLDER tag than any tag a later launch waits for, whatever grid sizes the launches had, and the two
// cudaMemsetAsync nodes per step (forward + backward) are gone from the stream / the CUDA graph.  Tags and the counter are
// compared modulo 2^32 (equality / signed difference), so wrap-around is harmless.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gode.h"

namespace gode {

constexpr int kGsMaxVals = 4;
constexpr int kGsFlagMaxCtas = 16;  // measured inside dopri5_fwd_kernel (scripts/dp5_lane_sweep.py): beyond 16 CTAs the counter wins

struct GridSyncWs {
  unsigned int* counter;      // 256-byte header: [0] arrival counter, [1] epoch base, [2] counter base
  unsigned long long* slots;  // [2][kGsMaxVals][gridDim.x] tagged words
  unsigned long long* rows;   // tagged parameter-gradient rows [gridDim.x][P] (small_field.cuh::reduce_param_grads)
};

// Fixed size of the sync region at the front of every workspace (so that scratch data behind it can never be mistaken for
// a tagged word by a later launch with a larger grid): header + slots for up to kSyncMaxGrid CTAs.
constexpr size_t kSyncRegionBytes = GODE_SYNC_REGION_BYTES;
constexpr size_t kSyncSlotBytes = 256 * 1024;                       // header + all-reduce slots
constexpr size_t kSyncRowBytes = kSyncRegionBytes - kSyncSlotBytes;  // tagged gradient rows (only ever written as tagged words)
constexpr int kSyncMaxGrid = (int)((kSyncSlotBytes - 256) / (sizeof(unsigned long long) * 2 * kGsMaxVals));

__host__ __device__ inline size_t grid_sync_bytes(int /*grid*/) { return kSyncRegionBytes; }
__host__ __device__ inline void grid_sync_bind(GridSyncWs& ws, void* base) {
  ws.counter = reinterpret_cast<unsigned int*>(base);
  ws.slots = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + 256);
  ws.rows = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + kSyncSlotBytes);
}

// Programmatic dependent launch (PDL): a kernel launched with programmatic stream serialisation may start while the
// previous kernel on the stream is still running; griddep_wait() blocks until that kernel has completed and its memory is
// visible (a no-op for an ordinary launch).  griddep_launch_dependents() lets the NEXT kernel's CTAs be scheduled early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Per-thread view of the persistent counters: `epoch` is the running tag, `ctarget` the running counter target.
struct SyncState {
  unsigned int epoch, ctarget;
  __device__ __forceinline__ void begin(const GridSyncWs& ws) {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch) : "l"(ws.counter + 1) : "memory");
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ctarget) : "l"(ws.counter + 2) : "memory");
  }
  // thread 0 of CTA 0, after the kernel's LAST grid-wide operation
  __device__ __forceinline__ void finish(const GridSyncWs& ws) const {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ws.counter + 1), "r"(epoch) : "memory");
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ws.counter + 2), "r"(ctarget) : "memory");
  }
};

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void arrive_counter(unsigned int* c) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
}
__device__ __forceinline__ void wait_counter(const unsigned int* c, unsigned int target) {
  unsigned int seen;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(c) : "memory");
  } while ((int)(seen - target) < 0);
}

// Sum NV per-thread values over the whole grid.  `v` enters as this thread's contribution (already masked for padding
// lanes) and leaves as the grid total, identical in every thread.  s_f: shared float[WARPS*NV]; s_d: shared double[NV].
template <int NV, int WARPS>
__device__ __forceinline__ void grid_allreduce_sum(double (&v)[NV], float* s_f, double* s_d, const GridSyncWs& ws,
                                                   SyncState& ss, int lane, int warp) {
  static_assert(NV <= kGsMaxVals, "raise kGsMaxVals");
  const unsigned int epoch = ++ss.epoch;
  const int grid = gridDim.x;
  if (grid > kGsFlagMaxCtas) ss.ctarget += (unsigned int)grid;
  // warp partials (fp32 is ample: the result is rounded to fp32 anyway; CTA sums are combined in fp64)
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
    float mine = 0.f;
    if (lane < NV) {
      mine = s_f[lane];
      for (int q = 1; q < WARPS; ++q) mine += s_f[q * NV + lane];
    }
    if (grid <= kGsFlagMaxCtas) {
      unsigned long long* buf = ws.slots + (size_t)(epoch & 1u) * kGsMaxVals * grid;
      if (lane < NV)
        st_relaxed_u64(buf + (size_t)lane * grid + blockIdx.x,
                       (unsigned long long)__float_as_uint(mine) | ((unsigned long long)epoch << 32));
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const unsigned long long* col = buf + (size_t)k * grid;
        unsigned long long x[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = u * 32 + lane;
          x[u] = c < grid ? ld_relaxed_u64(col + c) : ((unsigned long long)epoch << 32);
        }
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = u * 32 + lane;
          while (c < grid && (unsigned int)(x[u] >> 32) != epoch) x[u] = ld_relaxed_u64(col + c);
          s += (double)__uint_as_float((unsigned int)x[u]);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_d[k] = s;
      }
    } else {
      // Same tagged words, but the O(CTAs^2) polling is replaced by ONE hot word: a relaxed counter that is only a
      // hint that everybody has probably published.  Correctness still rests on the tags alone, so nothing here is a
      // release/acquire and the arrival never waits for the CTA's outstanding trajectory/checkpoint stores to drain
      // (a red.release did, and cost ~3 us per attempt inside the solver).
      unsigned long long* buf = ws.slots + (size_t)(epoch & 1u) * kGsMaxVals * grid;
      if (lane < NV)
        st_relaxed_u64(buf + (size_t)lane * grid + blockIdx.x,
                       (unsigned long long)__float_as_uint(mine) | ((unsigned long long)epoch << 32));
      if (lane == 0) {
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ws.counter) : "memory");
        const unsigned int target = ss.ctarget;
        unsigned int seen;
        do {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ws.counter) : "memory");
        } while ((int)(seen - target) < 0);
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const unsigned long long* col = buf + (size_t)k * grid;
        double s = 0.0;
        for (int c0 = 0; c0 < grid; c0 += 256) {
          unsigned long long x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            x[u] = c < grid ? ld_relaxed_u64(col + c) : ((unsigned long long)epoch << 32);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 32 + lane;
            while (c < grid && (unsigned int)(x[u] >> 32) != epoch) x[u] = ld_relaxed_u64(col + c);
            s += (double)__uint_as_float((unsigned int)x[u]);
          }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_d[k] = s;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_d[k];
}

// Plain grid barrier.  Orders prior global writes of every CTA before later global reads (L2 path: __ldcg) of any
// CTA: bar.sync, then one thread does a gpu-scope release on the counter (cumulative over the CTA's writes ordered
// before the bar.sync), polls it with acquire loads, and a second bar.sync releases the rest of the CTA.
__device__ __forceinline__ void grid_barrier(const GridSyncWs& ws, SyncState& ss) {
  ss.ctarget += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    arrive_counter(ws.counter);
    wait_counter(ws.counter, ss.ctarget);
    __threadfence();
  }
  __syncthreads();
}

}  // namespace gode
