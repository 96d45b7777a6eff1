// mufu_half_bench.cu — does MUFU.TANH cost less when only lanes 0..15 of every warp are active?  (decides whether M=64 tiles,
// whose accumulator rows live on lanes 0..15 of each TMEM quadrant, would halve the tanh time of a half-size tile)
#include <cstdio>
#include <cuda_runtime.h>
template <int ACTIVE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const bool on = (threadIdx.x & 31) < ACTIVE;
  __syncthreads();
  long long t0 = clock64();
  if (on)
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
      }
    }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
int main() {
  float* d; long long* c; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&c, 148 * 8);
  for (int act : {32, 16, 8}) {
    const int iters = 1000;
    if (act == 32) k<32><<<148, 512>>>(d, c, iters); else if (act == 16) k<16><<<148, 512>>>(d, c, iters); else k<8><<<148, 512>>>(d, c, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("active lanes per warp %2d: %.2f cycles per warp-level MUFU instruction per SM (16 warps)\n", act, (double)mx / (iters * 64.0 * 16));
  }
  return 0;
}
