// tc_rk4_wide.cu — tensor-core RK4 (3/8 rule) forward for the wide field D=64, H=256 (BF16 operands, FP32 accumulate).
//
// This is the shape where the contraction dominates (1 MFLOP per trajectory-step, arithmetic intensity ~1365 FLOP/B):
//   layer 1: 128 x 256 x 64  -> 4  tcgen05.mma (M=128, N=256, K=16) into TMEM columns [0,256)
//   layer 2: 128 x 64 x 256  -> 16 tcgen05.mma (M=128, N=64,  K=16) into TMEM columns [256,320)
// One CTA = 256 threads = one 128-trajectory tile; two threads share a trajectory row (TMEM lane): warp w reads lane
// quadrant w%4 and column half w/4, so every Runge–Kutta vector is 32 registers per thread.  tanh runs two-wide on
// packed bf16 (tanh.approx.bf16x2) and its result is the layer-2 operand without another conversion.
// Weights (132 KB fp32) arrive by TMA bulk copies through the (not yet used) A2 staging area and are re-tiled to BF16
// UMMA layout once per CTA; 146 KB shared memory and 512 TMEM columns => one CTA per SM, persistent over tiles.
#include "launch.h"
#include "tc_common.cuh"

namespace gode {

constexpr float kWT = 0.33333334f;

struct TcWideArgs {
  const float *y0, *W1, *b1, *W2, *b2;
  float* traj;
  const float* dt_dev;
  int B, T, layout;
  float dt_val[GODE_MAX_HOST_STEPS];
};

template <int D, int H>
struct TcWideShape {
  static constexpr int TILE = 128;
  static constexpr int KC1 = D / 8, KC2 = H / 8;       // 16-byte (8 x bf16) K chunks
  static constexpr int NM1 = KC1 / 2, NM2 = KC2 / 2;   // MMAs per layer
  static constexpr int DH = D / 2, HH = H / 2;          // columns per thread (two threads per row)
  static constexpr int OFF_BIAS = 0;                    // fp32 b1 | b2
  static constexpr int OFF_B1 = ((H + D) * 4 + 127) / 128 * 128;
  static constexpr int OFF_B2 = OFF_B1 + H * D * 2;
  static constexpr int OFF_A1 = OFF_B2 + D * H * 2;
  static constexpr int OFF_A2 = OFF_A1 + TILE * D * 2;
  static constexpr int OFF_BAR = OFF_A2 + TILE * H * 2;
  static constexpr int BYTES = OFF_BAR + 64;
  static_assert(TILE * H * 2 >= H * D * 4, "A2 doubles as the fp32 staging area of one weight matrix");
  static_assert(H + D <= 512, "TMEM columns");
};

__device__ __forceinline__ size_t tcw_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

template <int D, int H>
__global__ void __launch_bounds__(256, 1) tc_rk4_fwd_wide_kernel(const __grid_constant__ TcWideArgs p) {
  using S = TcWideShape<D, H>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* bias = reinterpret_cast<float*>(smem + S::OFF_BIAS);
  unsigned char* B1 = smem + S::OFF_B1;
  unsigned char* B2 = smem + S::OFF_B2;
  unsigned char* A1 = smem + S::OFF_A1;
  unsigned char* A2 = smem + S::OFF_A2;
  float* stagef = reinterpret_cast<float*>(A2);
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* mbar_m = mbar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_w + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, hf = warp >> 2;

  if (warp == 0) tc::tmem_alloc(s_tmem, 512);
  if (tid == 0) {
    tc::mbar_init(mbar_w, 1);
    tc::mbar_init(mbar_m, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  // ---- weights: TMA bulk -> fp32 staging (A2 region) -> BF16 UMMA layout, one matrix at a time ----
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, H * D * 4 + (H + D) * 4);
    tc::tma_bulk_g2s(stagef, p.W1, H * D * 4, mbar_w);
    tc::tma_bulk_g2s(bias, p.b1, H * 4, mbar_w);
    tc::tma_bulk_g2s(bias + H, p.b2, D * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 0);
  for (int idx = tid; idx < H * S::KC1; idx += 256) {  // B1[kc][n][16B] <- W1[n][kc*8 .. +8)
    const int n = idx % H, kc = idx / H;
    const float* v = stagef + n * D + kc * 8;
    *reinterpret_cast<uint4*>(B1 + (size_t)(kc * H + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  tc::fence_async_smem();  // order our generic reads/writes of the staging area before the next bulk copy lands on it
  __syncthreads();
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, D * H * 4);
    tc::tma_bulk_g2s(stagef, p.W2, D * H * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 1);
  for (int idx = tid; idx < D * S::KC2; idx += 256) {  // B2[kc][n][16B] <- W2[n][kc*8 .. +8)
    const int n = idx % D, kc = idx / D;
    const float* v = stagef + n * H + kc * 8;
    *reinterpret_cast<uint4*>(B2 + (size_t)(kc * D + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  tc::fence_async_smem();
  __syncthreads();

  constexpr uint32_t idesc1 = tc::make_idesc(tc::kFmtBF16, 128, H);
  constexpr uint32_t idesc2 = tc::make_idesc(tc::kFmtBF16, 128, D);
  const uint64_t dA1 = tc::make_smem_desc(tc::smem_u32(A1), S::TILE * 16, 128);
  const uint64_t dA2 = tc::make_smem_desc(tc::smem_u32(A2), S::TILE * 16, 128);
  const uint64_t dB1 = tc::make_smem_desc(tc::smem_u32(B1), H * 16, 128);
  const uint64_t dB2 = tc::make_smem_desc(tc::smem_u32(B2), D * 16, 128);
  const float* b1s = bias + hf * S::HH;
  const float* b2s = bias + H + hf * S::DH;
  uint32_t phase = 0;

  auto feval = [&](const float(&u)[S::DH], float(&f)[S::DH]) {
#pragma unroll
    for (int c = 0; c < S::DH / 8; ++c)
      *reinterpret_cast<uint4*>(A1 + (size_t)((hf * (S::DH / 8) + c) * S::TILE + row) * 16) =
          make_uint4(tc::pack_bf16x2(u[8 * c], u[8 * c + 1]), tc::pack_bf16x2(u[8 * c + 2], u[8 * c + 3]),
                     tc::pack_bf16x2(u[8 * c + 4], u[8 * c + 5]), tc::pack_bf16x2(u[8 * c + 6], u[8 * c + 7]));
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < S::NM1; ++j)
        tc::mma_ss<false>(tmem, dA1 + (uint64_t)((2 * j * S::TILE * 16) >> 4), dB1 + (uint64_t)((2 * j * H * 16) >> 4), idesc1, j > 0);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
    // 64 columns per iteration: four TMEM loads in flight before one wait
#pragma unroll 1
    for (int cq = 0; cq < S::HH / 64; ++cq) {
      uint32_t zr[4][16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tc::tmem_ld16_nowait(my_tmem + hf * S::HH + cq * 64 + q4 * 16, zr[q4]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t q[8];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(b1s + cq * 64 + q4 * 16 + i);
          q[i / 2] = tc::tanh_bf16x2(tc::pack_bf16x2(__uint_as_float(zr[q4][i]) + b.x, __uint_as_float(zr[q4][i + 1]) + b.y));
          q[i / 2 + 1] = tc::tanh_bf16x2(tc::pack_bf16x2(__uint_as_float(zr[q4][i + 2]) + b.z, __uint_as_float(zr[q4][i + 3]) + b.w));
        }
        const int kc = (hf * S::HH + cq * 64 + q4 * 16) / 8;
        *reinterpret_cast<uint4*>(A2 + (size_t)(kc * S::TILE + row) * 16) = make_uint4(q[0], q[1], q[2], q[3]);
        *reinterpret_cast<uint4*>(A2 + (size_t)((kc + 1) * S::TILE + row) * 16) = make_uint4(q[4], q[5], q[6], q[7]);
      }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < S::NM2; ++j)
        tc::mma_ss<false>(tmem + H, dA2 + (uint64_t)((2 * j * S::TILE * 16) >> 4), dB2 + (uint64_t)((2 * j * D * 16) >> 4), idesc2, j > 0);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
#pragma unroll
    for (int cb = 0; cb < S::DH / 16; ++cb) {
      float z[16];
      tc::tmem_ld16(my_tmem + H + hf * S::DH + cb * 16, z);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(b2s + cb * 16 + i);
        f[cb * 16 + i] = z[i] + b.x; f[cb * 16 + i + 1] = z[i + 1] + b.y;
        f[cb * 16 + i + 2] = z[i + 2] + b.z; f[cb * 16 + i + 3] = z[i + 3] + b.w;
      }
    }
  };

  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int ntiles = (p.B + S::TILE - 1) / S::TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * S::TILE + row;
    const bool valid = b < p.B;
    float y[S::DH];
#pragma unroll
    for (int i = 0; i < S::DH; ++i) y[i] = 0.f;
    if (valid) {
#pragma unroll
      for (int i = 0; i < S::DH; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(p.y0 + (size_t)b * D + hf * S::DH + i);
        y[i] = q.x; y[i + 1] = q.y; y[i + 2] = q.z; y[i + 3] = q.w;
      }
      float* o = p.traj + tcw_off(p.layout, 0, b, p.B, p.T, D) + hf * S::DH;
#pragma unroll
      for (int i = 0; i < S::DH; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
    }
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      float k[S::DH], u[S::DH], v[S::DH], acc[S::DH];
      feval(y, k);
#pragma unroll
      for (int i = 0; i < S::DH; ++i) { u[i] = y[i] + dt * k[i] * kWT; v[i] = k[i]; acc[i] = k[i]; }
      feval(u, k);
#pragma unroll
      for (int i = 0; i < S::DH; ++i) { u[i] = y[i] + dt * (k[i] - v[i] * kWT); v[i] = v[i] - k[i]; acc[i] += 3.f * k[i]; }
      feval(u, k);
#pragma unroll
      for (int i = 0; i < S::DH; ++i) { u[i] = y[i] + dt * (v[i] + k[i]); acc[i] += 3.f * k[i]; }
      feval(u, k);
#pragma unroll
      for (int i = 0; i < S::DH; ++i) y[i] = y[i] + (acc[i] + k[i]) * dt * 0.125f;
      if (valid) {
        float* o = p.traj + tcw_off(p.layout, s + 1, b, p.B, p.T, D) + hf * S::DH;
#pragma unroll
        for (int i = 0; i < S::DH; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int tc_rk4_fwd_wide(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                    int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st) {
  if (!(D == 64 && H == 256)) return GODE_ERR_SHAPE;
  using S = TcWideShape<64, 256>;
  TcWideArgs a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.B = B; a.T = T; a.layout = out_layout;
  if (dt_on_device) {
    a.dt_dev = dt;
  } else {
    if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
    for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  }
  auto kern = tc_rk4_fwd_wide_kernel<64, 256>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
  const int ntiles = (B + S::TILE - 1) / S::TILE;
  int grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, 256, S::BYTES, st>>>(a);
  return launch_status();
}

}  // namespace gode
