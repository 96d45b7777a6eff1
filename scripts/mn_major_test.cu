// mn_major_test.cu — developer check of the MN-major (transposed) no-swizzle shared-memory operand view used by the
// tensor-core adjoint: D[m][n] = sum_r X[r][m] * Y[r][n] with X (128 x 128) and Y (128 x 16) both stored row-tiles
// [chunk = col/8][row][8 cols] (the K-major layout of the forward), read here with MN-major descriptors
// (LBO = 128 B between 8-row groups = K direction, SBO = rows*16 B between 8-column chunks = M/N direction).
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace gode;

__global__ void k(const __nv_bfloat16* X, const __nv_bfloat16* Y, float* Dout) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t s_t;
  unsigned char* sX = smem;            // 16 chunks x 128 rows x 16 B = 32 KB
  unsigned char* sY = smem + 32768;    // 2 chunks x 128 rows x 16 B = 4 KB
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) tc::tmem_alloc(&s_t, 32);
  if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
  for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
    const int r = e / 128, c = e % 128;
    reinterpret_cast<__nv_bfloat16*>(sX)[((c / 8) * 128 + r) * 8 + c % 8] = X[e];
  }
  for (int e = threadIdx.x; e < 128 * 16; e += blockDim.x) {
    const int r = e / 16, c = e % 16;
    reinterpret_cast<__nv_bfloat16*>(sY)[((c / 8) * 128 + r) * 8 + c % 8] = Y[e];
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, s_t, 0);
  constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, 128, 16) | (1u << 15) | (1u << 16);  // A and B MN-major
  if (warp == 0 && tc::elect_one()) {
    const uint64_t dA = tc::make_smem_desc(tc::smem_u32(sX), 128, 128 * 16);
    const uint64_t dB = tc::make_smem_desc(tc::smem_u32(sY), 128, 128 * 16);
    for (int ks = 0; ks < 8; ++ks)  // 16 rows (K) per step = 2 core matrices along K = 256 B
      tc::mma_ss<false>(tmem, dA + (uint64_t)((ks * 256) >> 4), dB + (uint64_t)((ks * 256) >> 4), idesc, ks > 0);
    tc::mma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::fence_after_sync();
  float v[16];
  tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 16; ++i) Dout[(warp * 32 + (threadIdx.x & 31)) * 16 + i] = v[i];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<__nv_bfloat16> X(128 * 128), Y(128 * 16);
  std::vector<float> Xf(128 * 128), Yf(128 * 16);
  for (int i = 0; i < 128 * 128; ++i) { float f = (float)((i * 37 + 11) % 61 - 30) / 32.f; X[i] = __float2bfloat16(f); Xf[i] = __bfloat162float(X[i]); }
  for (int i = 0; i < 128 * 16; ++i) { float f = (float)((i * 53 + 7) % 43 - 21) / 16.f; Y[i] = __float2bfloat16(f); Yf[i] = __bfloat162float(Y[i]); }
  __nv_bfloat16 *dX, *dY; float* dD;
  cudaMalloc(&dX, X.size() * 2); cudaMalloc(&dY, Y.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  k<<<1, 128, 40960>>>(dX, dY, dD);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * 16);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 16; ++n) {
      double s = 0;
      for (int r = 0; r < 128; ++r) s += (double)Xf[r * 128 + m] * Yf[r * 16 + n];
      maxerr = fmax(maxerr, fabs(s - D[m * 16 + n])); maxref = fmax(maxref, fabs(s));
    }
  printf("MN-major A and B: max |err| = %.3e (max |ref| = %.3e)  %s  (%s)\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH",
         cudaGetErrorString(e));
  return 0;
}
