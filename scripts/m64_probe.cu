// m64_probe.cu — developer probe: where do the 64 rows of an M=64 tcgen05.mma (cta_group::1, bf16, N=16, K=16) land in TMEM?
// A[m][k] = (k == 0) ? (m + 1) : 0, B[n][k] = (k == 0) ? 1 : 0  ->  D[m][n] = m + 1.  Every lane's first 16 columns are dumped.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace gode;
__global__ void k(float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar; __shared__ uint32_t s_t;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) tc::tmem_alloc(&s_t, 32);
  if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);          // [kc 2][row 64][8]
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 4096);   // [kc 2][n 16][8]
  if (threadIdx.x < 64) A[(0 * 64 + threadIdx.x) * 8 + 0] = __float2bfloat16((float)(threadIdx.x + 1));
  if (threadIdx.x < 16) B[(0 * 16 + threadIdx.x) * 8 + 0] = __float2bfloat16(1.f);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, s_t, 0);
  // zero the accumulator region for all 128 lanes first (so untouched lanes read 0)
  { uint32_t z[8] = {0,0,0,0,0,0,0,0}; const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16); tc::tmem_st8(a, z); tc::tmem_st8(a + 8, z); tc::tmem_st_wait(); }
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (warp == 0 && tc::elect_one()) {
    const uint64_t dA = tc::make_smem_desc(tc::smem_u32(A), 64 * 16, 128);
    const uint64_t dB = tc::make_smem_desc(tc::smem_u32(B), 16 * 16, 128);
    tc::mma_ss<false>(tmem, dA, dB, tc::make_idesc(tc::kFmtBF16, 64, 16), 0);
    tc::mma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0); tc::fence_after_sync();
  float v[16];
  tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  out[threadIdx.x * 2] = v[0]; out[threadIdx.x * 2 + 1] = v[15];
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}
int main() {
  float* d; cudaMalloc(&d, 128 * 2 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  k<<<1, 128, 8192>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  float h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s\nlane: D[lane][0]\n", cudaGetErrorString(e));
  for (int l = 0; l < 128; ++l) printf("%3d:%3.0f%s", l, h[2 * l], (l % 16 == 15) ? "\n" : " ");
  return 0;
}
