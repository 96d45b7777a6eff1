// sync_bench.cu — microbenchmark of grid-wide all-reduce variants (developer tool; informs grid_sync.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sync_bench scripts/sync_bench.cu && /tmp/sync_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

__device__ __forceinline__ void st_rlx(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// variant 0: cooperative groups grid.sync() + partial array re-read
// variant 1: all-gather of tagged words, every CTA's warp 0 polls all slots
// variant 2: tagged words, CTA 0 gathers and publishes one broadcast word, everyone else polls that word
// variant 3: like 1 but with __nanosleep backoff
template <int VAR>
__global__ void k(unsigned long long* slots, double* partials, int iters, int work, float* out, long long* cyc) {
  cg::grid_group gg = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grid = gridDim.x;
  __shared__ double s_tot;
  __shared__ float s_w[32];
  float acc = tid * 1e-3f + blockIdx.x;
  double tot = 0;
  unsigned epoch = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int q = 0; q < work; ++q) acc = fmaf(acc, 1.0001f, 0.5f);  // stand-in for stage work
    float w = acc * 1e-6f;
    for (int off = 16; off; off >>= 1) w += __shfl_xor_sync(~0u, w, off);
    if (lane == 0) s_w[warp] = w;
    __syncthreads();
    ++epoch;
    if (VAR == 0) {
      if (tid == 0) {
        float s = 0;
        for (int q = 0; q < (int)blockDim.x / 32; ++q) s += s_w[q];
        partials[(epoch & 1) * grid + blockIdx.x] = s;
      }
      gg.sync();
      if (warp == 0) {
        double s = 0;
        for (int c = lane; c < grid; c += 32) s += __ldcg(partials + (epoch & 1) * grid + c);
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(~0u, s, off);
        if (lane == 0) s_tot = s;
      }
      __syncthreads();
      tot = s_tot;
    } else {
      unsigned long long* buf = slots + (size_t)(epoch & 1u) * (grid + 16);
      if (warp == 0) {
        if (lane == 0) {
          float s = 0;
          for (int q = 0; q < (int)blockDim.x / 32; ++q) s += s_w[q];
          st_rlx(buf + blockIdx.x, (unsigned long long)__float_as_uint(s) | ((unsigned long long)epoch << 32));
        }
        if (VAR == 1 || VAR == 3 || blockIdx.x == 0) {
          double s = 0;
          for (int c0 = 0; c0 < grid; c0 += 128) {
            unsigned long long x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              int c = c0 + u * 32 + lane;
              x[u] = c < grid ? ld_rlx(buf + c) : ((unsigned long long)epoch << 32);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              int c = c0 + u * 32 + lane;
              while (c < grid && (unsigned)(x[u] >> 32) != epoch) {
                if (VAR == 3) __nanosleep(40);
                x[u] = ld_rlx(buf + c);
              }
              s += (double)__uint_as_float((unsigned)x[u]);
            }
          }
          for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(~0u, s, off);
          if (VAR == 2) {
            if (lane == 0) st_rlx(buf + grid + 8, (unsigned long long)__float_as_uint((float)s) | ((unsigned long long)epoch << 32));
            s = (double)(float)s;
          }
          if (lane == 0) s_tot = s;
        } else {
          if (lane == 0) {
            unsigned long long x;
            do { x = ld_rlx(buf + grid + 8); } while ((unsigned)(x >> 32) != epoch);
            s_tot = (double)__uint_as_float((unsigned)x);
          }
        }
      }
      __syncthreads();
      tot = s_tot;
    }
    acc += (float)tot * 1e-9f;
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (tid == 0) out[blockIdx.x] = acc;
}

// variant 4: counter barrier done by hand: st slot (plain) ; red.release.gpu counter ; poll ONE word ; parallel slot loads
__global__ void k4(unsigned long long* slots, unsigned* counter, int iters, int work, float* out, long long* cyc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grid = gridDim.x;
  __shared__ double s_tot;
  __shared__ float s_w[32];
  float acc = tid * 1e-3f + blockIdx.x;
  unsigned epoch = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int q = 0; q < work; ++q) acc = fmaf(acc, 1.0001f, 0.5f);
    float w = acc * 1e-6f;
    for (int off = 16; off; off >>= 1) w += __shfl_xor_sync(~0u, w, off);
    if (lane == 0) s_w[warp] = w;
    __syncthreads();
    ++epoch;
    float* buf = reinterpret_cast<float*>(slots) + (size_t)(epoch & 1u) * grid;
    if (warp == 0) {
      if (lane == 0) {
        float s = 0;
        for (int q = 0; q < (int)blockDim.x / 32; ++q) s += s_w[q];
        __stcg(buf + blockIdx.x, s);
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned seen;
        const unsigned target = epoch * grid;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while ((int)(seen - target) < 0);
      }
      __syncwarp();
      double s = 0;
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { int c = u * 32 + lane; x[u] = c < grid ? __ldcg(buf + c) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) s += (double)x[u];
      for (int c = 256 + lane; c < grid; c += 32) s += (double)__ldcg(buf + c);
      for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(~0u, s, off);
      if (lane == 0) s_tot = s;
    }
    __syncthreads();
    acc += (float)s_tot * 1e-9f;
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (tid == 0) out[blockIdx.x] = acc;
}

// variant 5: clusters of CS CTAs reduce through distributed shared memory; cluster leaders all-gather tagged words
// (variant 1 among grid/CS leaders); leaders push the total into every member's shared memory.
template <int CS>
__global__ void k5(unsigned long long* slots, int iters, int work, float* out, long long* cyc) {
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nlead = gridDim.x / CS, crank = cl.block_rank(), cid = blockIdx.x / CS;
  __shared__ double s_tot;
  __shared__ float s_cta;
  __shared__ float s_w[32];
  float acc = tid * 1e-3f + blockIdx.x;
  unsigned epoch = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int q = 0; q < work; ++q) acc = fmaf(acc, 1.0001f, 0.5f);
    float w = acc * 1e-6f;
    for (int off = 16; off; off >>= 1) w += __shfl_xor_sync(~0u, w, off);
    if (lane == 0) s_w[warp] = w;
    __syncthreads();
    ++epoch;
    if (tid == 0) {
      float s = 0;
      for (int q = 0; q < (int)blockDim.x / 32; ++q) s += s_w[q];
      s_cta = s;
    }
    cl.sync();
    if (crank == 0 && warp == 0) {
      float v = 0.f;
      if (lane < CS) v = *cl.map_shared_rank(&s_cta, lane);
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(~0u, v, off);
      unsigned long long* buf = slots + (size_t)(epoch & 1u) * nlead;
      if (lane == 0) st_rlx(buf + cid, (unsigned long long)__float_as_uint(v) | ((unsigned long long)epoch << 32));
      double s = 0;
      for (int c = lane; c < nlead; c += 32) {
        unsigned long long x = ld_rlx(buf + c);
        while ((unsigned)(x >> 32) != epoch) x = ld_rlx(buf + c);
        s += (double)__uint_as_float((unsigned)x);
      }
      for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(~0u, s, off);
      if (lane < CS) *cl.map_shared_rank(&s_tot, lane) = s;
    }
    cl.sync();
    acc += (float)s_tot * 1e-9f;
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (tid == 0) out[blockIdx.x] = acc;
}

void run4(int grid, int block, int iters, int work) {
  unsigned long long* slots; unsigned* counter; float* out; long long* cyc;
  cudaMalloc(&slots, 2 * (grid + 16) * 8); cudaMalloc(&counter, 256); cudaMalloc(&out, grid * 4); cudaMalloc(&cyc, 8);
  void* args[] = {&slots, &counter, &iters, &work, &out, &cyc};
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(counter, 0, 256);
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k4, dim3(grid), dim3(block), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("var 4 grid %d failed: %s\n", grid, cudaGetErrorString(e)); exit(1); }
  }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("var 4 grid %4d block %4d work %5d: %8.0f cycles/iter\n", grid, block, work, (double)h / iters);
  cudaFree(slots); cudaFree(counter); cudaFree(out); cudaFree(cyc);
}

template <int CS>
void run5(int grid, int block, int iters, int work) {
  if (grid % CS) return;
  unsigned long long* slots; float* out; long long* cyc;
  cudaMalloc(&slots, 2 * (grid + 16) * 8); cudaMalloc(&out, grid * 4); cudaMalloc(&cyc, 8);
  if (CS > 8) cudaFuncSetAttribute(k5<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(slots, 0, 2 * (grid + 16) * 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k5<CS>, slots, iters, work, out, cyc);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("var 5 CS %d grid %d failed: %s\n", CS, grid, cudaGetErrorString(e)); cudaGetLastError(); return; }
  }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("var 5 CS %2d grid %4d block %4d work %5d: %8.0f cycles/iter\n", CS, grid, block, work, (double)h / iters);
  cudaFree(slots); cudaFree(out); cudaFree(cyc);
}

template <int VAR>
void run(int grid, int block, int iters, int work) {
  unsigned long long* slots;
  double* partials;
  float* out;
  long long* cyc;
  cudaMalloc(&slots, 2 * (grid + 16) * 8);
  cudaMemset(slots, 0, 2 * (grid + 16) * 8);
  cudaMalloc(&partials, 2 * grid * 8);
  cudaMalloc(&out, grid * 4);
  cudaMalloc(&cyc, 8);
  void* args[] = {&slots, &partials, &iters, &work, &out, &cyc};
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(slots, 0, 2 * (grid + 16) * 8);
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k<VAR>, dim3(grid), dim3(block), args, 0, 0);
    if (e != cudaSuccess) { printf("launch failed var %d grid %d: %s\n", VAR, grid, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("run failed var %d grid %d: %s\n", VAR, grid, cudaGetErrorString(e)); exit(1); }
  }
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("var %d grid %4d block %4d work %5d: %8.0f cycles/iter\n", VAR, grid, block, work, (double)h / iters);
  cudaFree(slots); cudaFree(partials); cudaFree(out); cudaFree(cyc);
}

int main() {
  const int iters = 200;
  for (int work : {0}) {
    for (int grid : {16, 32, 64, 128, 144, 256, 288}) {
      for (int block : {128, 256, 512}) {
        run<0>(grid, block, iters, work);
        run<1>(grid, block, iters, work);
        run4(grid, block, iters, work);
        run5<4>(grid, block, iters, work);
        run5<8>(grid, block, iters, work);
        run5<16>(grid, block, iters, work);
      }
    }
  }
  return 0;
}
