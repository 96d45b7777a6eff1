// mma_bench.cu — developer microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, K=16) as a function of N and of
// where A lives (shared memory descriptor vs TMEM), one CTA per SM, one elected lane issuing NI MMAs then one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../gan_ode_b200/csrc -o mma_bench.bin mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace gode;

template <int N, bool TS, int HAMMER>
__global__ void bench(long long* out, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t s_t;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) tc::tmem_alloc(&s_t, 512);
  if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, s_t, 0);
  constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, 128, N);
  constexpr int NI = 64;
  long long t0 = 0, t1 = 0;
  uint32_t phase = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    if (HAMMER && warp >= 4) {  // other warps: TMEM loads/stores (HAMMER=1) or MUFU (HAMMER=2) on columns [0,128) while the MMAs run
      uint32_t z[16]; uint32_t q[8] = {1, 2, 3, 4, 5, 6, 7, 8};
      const uint32_t a = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      float f = threadIdx.x;
      for (int it = 0; it < (HAMMER == 1 ? 60 : 12); ++it) {
        if (HAMMER == 1) {
        tc::tmem_ld16_nowait(a, z); tc::tmem_ld16_nowait(a + 16, z); tc::tmem_ld16_nowait(a + 32, z); tc::tmem_ld16_nowait(a + 48, z);
        tc::tmem_ld_wait();
        q[0] ^= z[3];
        tc::tmem_st8(a + 64, q); tc::tmem_st8(a + 72, q);
        tc::tmem_st_wait();
        } else {
#pragma unroll
          for (int u = 0; u < 64; ++u) f = tc::tanh_approx(f);
        }
      }
      if (f == 123.f) out[0] = q[0];
    }
    if (threadIdx.x == 0) t0 = clock64();
    if (warp == 0 && tc::elect_one()) {
      const uint64_t dA = tc::make_smem_desc(tc::smem_u32(smem), 128 * 16, 128);
      const uint64_t dB = tc::make_smem_desc(tc::smem_u32(smem + 16384), N * 16, 128);
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        if constexpr (TS) tc::mma_ts_bf16(tmem + 256, tmem + (j % 8) * 8, dB, idesc, j > 0);
        else tc::mma_ss<false>(tmem + 256, dA + (uint64_t)(((j % 4) * 2 * 128 * 16) >> 4), dB, idesc, j > 0);
      }
      tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, phase);
    phase ^= 1;
    if (threadIdx.x == 0) { t1 = clock64(); if (r == reps - 1) out[blockIdx.x] = (t1 - t0); }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N, bool TS, int HAMMER>
void run() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  auto k = bench<N, TS, HAMMER>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  k<<<148, HAMMER ? 384 : 128, 48 * 1024>>>(d, 20);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("[other warps: %s] M=128 N=%3d K=16 A from %s: %7.1f cycles per MMA (64 back to back + commit + wait)  -> %5.0f FLOP/clk/SM  (%s)\n", HAMMER == 0 ? "idle" : HAMMER == 1 ? "LDTM/STTM" : "MUFU", N,
         TS ? "TMEM" : "smem", mx / 64.0, 2.0 * 128 * N * 16 / (mx / 64.0), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, false, 0>(); run<64, false, 0>(); run<128, false, 0>(); run<256, false, 0>();
  run<32, true, 0>(); run<64, true, 0>(); run<128, true, 0>(); run<256, true, 0>();
  run<256, false, 1>(); run<32, true, 1>(); run<64, true, 1>();
  run<256, false, 2>(); run<32, true, 2>();
  return 0;
}
