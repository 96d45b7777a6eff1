// tc_rk4.cu — fixed-grid RK4 (3/8 rule) forward with the MLP contractions on tcgen05 tensor cores.
//
// Tile kernel: one CTA = 128 threads = 128 trajectories; thread r owns trajectory row r for the whole time loop, so
// every Runge–Kutta vector (y, k-combinations) stays in its registers.  Per stage:
//
//   u (registers) --cvt--> A1 (smem, canonical K-major)  --tcgen05.mma x W1^T--> TMEM[:, 0:H)  --tcgen05.ld--> registers
//     +b1, tanh.approx --cvt--> A2 (smem)                --tcgen05.mma x W2^T--> TMEM[:, H:H+D) --tcgen05.ld--> +b2 = k
//
// Accumulators live in TMEM (fp32); operands are TF32 (round-to-nearest on store) or BF16.  The weights arrive once
// per CTA by TMA bulk copies (cp.async.bulk -> mbarrier complete_tx) and are re-tiled/converted in shared memory into
// the UMMA no-swizzle layout; they stay resident while the CTA walks its tiles (persistent grid-stride loop).
// Several CTAs share an SM (32 TMEM columns and ~20 KB shared memory each), so one CTA's MMA/TMEM round trip overlaps
// the others' tanh/pack epilogues.
// Tolerance of this mode: <= 2e-3 relative vs torchdiffeq (BASELINE.json north_star).
#include <stdlib.h>
#include "launch.h"
#include "tc_common.cuh"

namespace gode {

constexpr float kThird = 0.33333334f;

struct TcRk4Args {
  const float *y0, *W1, *b1, *W2, *b2;
  float* traj;
  const float* dt_dev;
  int B, T, layout;
  float dt_val[GODE_MAX_HOST_STEPS];
};

template <int D, int H, bool TF32>
struct TcShape {
  static constexpr int ELT = TF32 ? 4 : 2;      // bytes per operand element
  static constexpr int TPC = 16 / ELT;          // elements per 16-byte chunk
  static constexpr int KC1 = D / TPC;           // 16-byte K chunks of layer 1 (K = D)
  static constexpr int KC2 = H / TPC;           // layer 2 (K = H)
  static constexpr int NM1 = KC1 / 2, NM2 = KC2 / 2;  // MMAs per layer (each eats 32 bytes of K)
  static constexpr int TILE = 128;
  static constexpr int P = H * D + H + D * H + D;
  static constexpr uint32_t NCOLS = (H + D <= 32) ? 32 : (H + D <= 64) ? 64 : (H + D <= 128) ? 128 : (H + D <= 256) ? 256 : 512;
  // shared memory carve-up (bytes)
  static constexpr int OFF_RAW = 0;                               // fp32 [W1|b1|W2|b2] as bulk-copied
  static constexpr int OFF_B1 = OFF_RAW + ((P * 4 + 127) / 128) * 128;
  static constexpr int OFF_B2 = OFF_B1 + H * D * ELT;
  static constexpr int OFF_A1 = OFF_B2 + D * H * ELT;
  static constexpr int OFF_A2 = OFF_A1 + TILE * D * ELT;
  static constexpr int OFF_BAR = OFF_A2 + TILE * H * ELT;
  static constexpr int BYTES = OFF_BAR + 64;
  static_assert(D % 16 == 0 && H % 16 == 0, "tensor-core tile kernels need D, H multiples of 16");
  static_assert(H + D <= 512, "TMEM has 512 columns");
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// write one 16-byte K chunk of a row-owned operand row: elements v[0..TPC)
template <bool TF32>
__device__ __forceinline__ void store_chunk(unsigned char* dst, const float* v) {
  uint4 q;
  if constexpr (TF32) {
    q = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
  } else {
    q = make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                   tc::pack_bf16x2(v[6], v[7]));
  }
  *reinterpret_cast<uint4*>(dst) = q;
}

__device__ __forceinline__ size_t tc_traj_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

template <int D, int H, bool TF32>
__global__ void __launch_bounds__(128, 5) tc_rk4_fwd_kernel(const __grid_constant__ TcRk4Args p) {
  using S = TcShape<D, H, TF32>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* raw = reinterpret_cast<float*>(smem + S::OFF_RAW);
  const float* b1s = raw + H * D;
  const float* b2s = raw + H * D + H + D * H;
  unsigned char* B1 = smem + S::OFF_B1;
  unsigned char* B2 = smem + S::OFF_B2;
  unsigned char* A1 = smem + S::OFF_A1;
  unsigned char* A2 = smem + S::OFF_A2;
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* mbar_m = mbar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_w + 2);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform: MMA descriptors stay in uniform registers

  if (warp == 0) tc::tmem_alloc(s_tmem, S::NCOLS);
  if (tid == 0) {
    tc::mbar_init(mbar_w, 1);
    tc::mbar_init(mbar_m, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);

  // weights: TMA bulk copies, once per CTA
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, S::P * 4);
    tc::tma_bulk_g2s(raw, p.W1, H * D * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D, p.b1, H * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D + H, p.W2, D * H * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D + H + D * H, p.b2, D * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 0);
  // re-tile into the UMMA layout [k_chunk][n][16 B] (K-major B operands: B1 = W1 (N=H,K=D), B2 = W2 (N=D,K=H))
  for (int idx = tid; idx < H * S::KC1; idx += 128) {
    const int n = idx % H, kc = idx / H;
    store_chunk<TF32>(B1 + (size_t)(kc * H + n) * 16, raw + n * D + kc * S::TPC);
  }
  for (int idx = tid; idx < D * S::KC2; idx += 128) {
    const int n = idx % D, kc = idx / D;
    store_chunk<TF32>(B2 + (size_t)(kc * D + n) * 16, raw + H * D + H + n * H + kc * S::TPC);
  }
  tc::fence_async_smem();
  __syncthreads();

  constexpr uint32_t kFmt = TF32 ? tc::kFmtTF32 : tc::kFmtBF16;
  constexpr uint32_t idesc1 = tc::make_idesc(kFmt, 128, H);
  constexpr uint32_t idesc2 = tc::make_idesc(kFmt, 128, D);
  const uint64_t dA1 = tc::make_smem_desc(tc::smem_u32(A1), S::TILE * 16, 128);
  const uint64_t dA2 = tc::make_smem_desc(tc::smem_u32(A2), S::TILE * 16, 128);
  const uint64_t dB1 = tc::make_smem_desc(tc::smem_u32(B1), H * 16, 128);
  const uint64_t dB2 = tc::make_smem_desc(tc::smem_u32(B2), D * 16, 128);
  uint32_t phase = 0;

  auto feval = [&](const float(&u)[D], float(&f)[D]) {
#pragma unroll
    for (int c = 0; c < S::KC1; ++c) store_chunk<TF32>(A1 + (size_t)(c * S::TILE + tid) * 16, &u[c * S::TPC]);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < S::NM1; ++j)
        tc::mma_ss<TF32>(tmem, dA1 + (uint64_t)((2 * j * S::TILE * 16) >> 4), dB1 + (uint64_t)((2 * j * H * 16) >> 4), idesc1, j > 0);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
#pragma unroll
    for (int cb = 0; cb < H / 16; ++cb) {
      float z[16];
      tc::tmem_ld16(my_tmem + cb * 16, z);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(b1s + cb * 16 + i);
        z[i] = tc::tanh_approx(z[i] + b.x); z[i + 1] = tc::tanh_approx(z[i + 1] + b.y);
        z[i + 2] = tc::tanh_approx(z[i + 2] + b.z); z[i + 3] = tc::tanh_approx(z[i + 3] + b.w);
      }
#pragma unroll
      for (int c = 0; c < 16 / S::TPC; ++c)
        store_chunk<TF32>(A2 + (size_t)((cb * (16 / S::TPC) + c) * S::TILE + tid) * 16, &z[c * S::TPC]);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
#pragma unroll
      for (int j = 0; j < S::NM2; ++j)
        tc::mma_ss<TF32>(tmem + H, dA2 + (uint64_t)((2 * j * S::TILE * 16) >> 4), dB2 + (uint64_t)((2 * j * D * 16) >> 4), idesc2, j > 0);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
#pragma unroll
    for (int cb = 0; cb < D / 16; ++cb) {
      float z[16];
      tc::tmem_ld16(my_tmem + H + cb * 16, z);
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
    const int b = tile * S::TILE + tid;
    const bool valid = b < p.B;
    float y[D];
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = 0.f;
    if (valid) {
#pragma unroll
      for (int i = 0; i < D; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(p.y0 + (size_t)b * D + i);
        y[i] = q.x; y[i + 1] = q.y; y[i + 2] = q.z; y[i + 3] = q.w;
      }
      float* o = p.traj + tc_traj_off(p.layout, 0, b, p.B, p.T, D);
#pragma unroll
      for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
    }
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      // 3/8 rule with running combinations: after k2 only v = k1 - k2 and acc = k1 + 3 k2 are needed
      float k[D], u[D], v[D], acc[D];
      feval(y, k);  // k1
#pragma unroll
      for (int i = 0; i < D; ++i) { u[i] = y[i] + dt * k[i] * kThird; v[i] = k[i]; acc[i] = k[i]; }
      feval(u, k);  // k2
#pragma unroll
      for (int i = 0; i < D; ++i) { u[i] = y[i] + dt * (k[i] - v[i] * kThird); v[i] = v[i] - k[i]; acc[i] += 3.f * k[i]; }
      feval(u, k);  // k3
#pragma unroll
      for (int i = 0; i < D; ++i) { u[i] = y[i] + dt * (v[i] + k[i]); acc[i] += 3.f * k[i]; }
      feval(u, k);  // k4
#pragma unroll
      for (int i = 0; i < D; ++i) y[i] = y[i] + (acc[i] + k[i]) * dt * 0.125f;
      if (valid) {
        float* o = p.traj + tc_traj_off(p.layout, s + 1, b, p.B, p.T, D);
#pragma unroll
        for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, S::NCOLS);
}

// ---- BF16, D = H = 16: the lean variant ------------------------------------------------------------------------------------
// Same tile mapping (one thread per trajectory row, several CTAs per SM), with what the wide kernel taught
// (tc_rk4_wide.cu): b1 rides in the layer-1 MMA as one extra K step against a (1,1,0,..) chunk of the input tile; the hidden
// activation goes back to TMEM (tcgen05.st, packed bf16 over the consumed z1 columns) and is layer 2's A operand, so it
// never touches shared memory; the 3/8-rule update is regrouped to two live vectors and applied inside the layer-2
// epilogue, which also packs the next stage's input.  ~100 instead of ~170 instructions per trajectory-stage.
constexpr int kLeanD = 16, kLeanTile = 128, kLeanCH = kLeanTile * 16;
constexpr int LEAN_OFF_RAW = 0;                      // fp32 [W1|b1|W2|b2] (544 floats) as bulk-copied
constexpr int LEAN_OFF_B1 = 2304;                    // [W1|b1]: 4 K chunks x 16 rows x 16 B
constexpr int LEAN_OFF_B2 = LEAN_OFF_B1 + 4 * 256;   // W2: 2 K chunks x 16 rows x 16 B
constexpr int LEAN_OFF_A1 = LEAN_OFF_B2 + 2 * 256;   // input tile: chunks 0,1 = u ; 2 = (1,1,0,..) ; 3 = 0
constexpr int LEAN_OFF_BAR = LEAN_OFF_A1 + 4 * kLeanCH;
constexpr int LEAN_BYTES = LEAN_OFF_BAR + 64;
static_assert(LEAN_OFF_A1 % 128 == 0, "operand tiles are 128-byte aligned");

__global__ void __launch_bounds__(128, 8) tc_rk4_fwd_bf16_lean_kernel(const __grid_constant__ TcRk4Args p) {
  constexpr int D = kLeanD, H = 16;
  extern __shared__ __align__(128) unsigned char smem[];
  float* raw = reinterpret_cast<float*>(smem + LEAN_OFF_RAW);
  const float* b2s = raw + H * D + H + D * H;
  unsigned char* B1 = smem + LEAN_OFF_B1;
  unsigned char* B2 = smem + LEAN_OFF_B2;
  unsigned char* A1 = smem + LEAN_OFF_A1;
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(smem + LEAN_OFF_BAR);
  uint64_t* mbar_m = mbar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_w + 2);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 0) tc::tmem_alloc(s_tmem, 32);
  if (tid == 0) {
    tc::mbar_init(mbar_w, 1);
    tc::mbar_init(mbar_m, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, 544 * 4);
    tc::tma_bulk_g2s(raw, p.W1, H * D * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D, p.b1, H * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D + H, p.W2, D * H * 4, mbar_w);
    tc::tma_bulk_g2s(raw + H * D + H + D * H, p.b2, D * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 0);
  if (tid < 32) {            // B1 data chunks [kc][j][8 d]
    const int kc = tid >> 4, j = tid & 15;
    store_chunk<false>(B1 + (kc * 16 + j) * 16, raw + j * D + kc * 8);
  } else if (tid < 48) {     // K = 16, 17: b1 as two bf16 terms (exact to 2^-17) against the (1,1) columns of the input tile
    const int j = tid & 15;
    const float b = raw[H * D + j];
    const float hi = __bfloat162float(__float2bfloat16_rn(b));
    *reinterpret_cast<uint4*>(B1 + (2 * 16 + j) * 16) = make_uint4(tc::pack_bf16x2(hi, b - hi), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(B1 + (3 * 16 + j) * 16) = make_uint4(0u, 0u, 0u, 0u);
  } else if (tid < 80) {     // B2 [kc][d][8 j]
    const int kc = (tid >> 4) - 3, d = tid & 15;
    store_chunk<false>(B2 + (kc * 16 + d) * 16, raw + H * D + H + d * H + kc * 8);
  }
  *reinterpret_cast<uint4*>(A1 + 2 * kLeanCH + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(A1 + 3 * kLeanCH + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_async_smem();
  __syncthreads();

  constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, 128, 16);
  const uint32_t sA1 = tc::smem_u32(A1), sB1 = tc::smem_u32(B1), sB2 = tc::smem_u32(B2);
  uint32_t phase = 0;

  // one f evaluation; A1 holds bf16(u) on entry; upd(i, k_i) returns element i of the next stage's input (packed into A1)
  auto feval = [&](auto&& upd) {
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      const uint64_t dA = tc::make_smem_desc(sA1, kLeanCH, 128), dB = tc::make_smem_desc(sB1, 256, 128);
      tc::mma_ss<false>(tmem, dA, dB, idesc, 0);
      tc::mma_ss<false>(tmem, dA + (uint64_t)((2 * kLeanCH) >> 4), dB + (uint64_t)((2 * 256) >> 4), idesc, 1);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
    {
      uint32_t z[16], q[8];
      tc::tmem_ld16_nowait(my_tmem, z);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        q[i / 2] = tc::pack_bf16x2(tc::tanh_approx(__uint_as_float(z[i])), tc::tanh_approx(__uint_as_float(z[i + 1])));
      tc::tmem_st8(my_tmem, q);  // h over the consumed z1 columns [0,8)
      tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      tc::mma_ts_bf16(tmem + 16, tmem, tc::make_smem_desc(sB2, 256, 128), idesc, 0);
      tc::mma_commit(mbar_m);
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
    {
      uint32_t z[16], q[8];
      tc::tmem_ld16_nowait(my_tmem + 16, z);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 b = *reinterpret_cast<const float2*>(b2s + i);
        q[i / 2] = tc::pack_bf16x2(upd(i, __uint_as_float(z[i]) + b.x), upd(i + 1, __uint_as_float(z[i + 1]) + b.y));
      }
      *reinterpret_cast<uint4*>(A1 + tid * 16) = make_uint4(q[0], q[1], q[2], q[3]);
      *reinterpret_cast<uint4*>(A1 + kLeanCH + tid * 16) = make_uint4(q[4], q[5], q[6], q[7]);
    }
  };

  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int ntiles = (p.B + kLeanTile - 1) / kLeanTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * kLeanTile + tid;
    const bool valid = b < p.B;
    float y[D], w[D];
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = 0.f;
    if (valid) {
#pragma unroll
      for (int i = 0; i < D; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(p.y0 + (size_t)b * D + i);
        y[i] = q.x; y[i + 1] = q.y; y[i + 2] = q.z; y[i + 3] = q.w;
      }
      float* o = p.traj + tc_traj_off(p.layout, 0, b, p.B, p.T, D);
#pragma unroll
      for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
    }
    store_chunk<false>(A1 + tid * 16, &y[0]);
    store_chunk<false>(A1 + kLeanCH + tid * 16, &y[8]);
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      const float dt3 = dt * kThird, dt8 = dt * 0.125f, dt38 = dt * 0.375f;
      feval([&](int i, float k) { w[i] = k; return y[i] + dt3 * k; });                                  // k1 (live: y, w = k1)
      feval([&](int i, float k) {                                                                        // k2 (live: P, w)
        const float u = y[i] + dt * (k - w[i] * kThird);
        const float P = y[i] + dt8 * (w[i] + 3.f * k);
        w[i] = y[i] + dt * (w[i] - k);
        y[i] = P;
        return u;
      });
      feval([&](int i, float k) { y[i] += dt38 * k; return w[i] + dt * k; });                            // k3 (live: P')
      feval([&](int i, float k) { y[i] += dt8 * k; return y[i]; });                                      // k4
      if (valid) {
        float* o = p.traj + tc_traj_off(p.layout, s + 1, b, p.B, p.T, D);
#pragma unroll
        for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

static int launch_tc_rk4_fwd_bf16_lean(TcRk4Args& a, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc_rk4_fwd_bf16_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAN_BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
  constexpr int kCtasPerSm = 8;  // 32 TMEM columns, 12 KB shared memory, <= 64 registers each
  const int ntiles = (a.B + kLeanTile - 1) / kLeanTile;
  int grid = kCtasPerSm * sm_count();
  if (grid > ntiles) grid = ntiles;
  tc_rk4_fwd_bf16_lean_kernel<<<grid, 128, LEAN_BYTES, st>>>(a);
  return launch_status();
}

template <int D, int H, bool TF32>
static int launch_tc_rk4_fwd(TcRk4Args& a, cudaStream_t st) {
  using S = TcShape<D, H, TF32>;
  auto kern = tc_rk4_fwd_kernel<D, H, TF32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
  // Persistent grid: enough CTAs to fill every SM to its register/TMEM limit (<= 96 registers -> 5 CTAs of 128
  // threads; 32 TMEM columns each).  Not a cooperative launch, so over-subscription would only queue CTAs.
  constexpr int kCtasPerSm = 5;
  static_assert(kCtasPerSm * (int)S::NCOLS <= 512, "TMEM columns are a per-SM resource");
  const int ntiles = (a.B + S::TILE - 1) / S::TILE;
  int grid = kCtasPerSm * sm_count();
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, 128, S::BYTES, st>>>(a);
  return launch_status();
}

int tc_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
               int dt_on_device, int B, int D, int H, int T, int precision, int out_layout, float* traj,
               cudaStream_t st) {
  TcRk4Args a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.B = B; a.T = T; a.layout = out_layout;
  if (dt_on_device) {
    a.dt_dev = dt;
  } else {
    if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
    for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  }
  const bool tf32 = precision == GODE_PREC_TF32;
  if (D == 16 && H == 16) {
    if (tf32) return launch_tc_rk4_fwd<16, 16, true>(a, st);
    return launch_tc_rk4_fwd_bf16_lean(a, st);
  }
  return GODE_ERR_SHAPE;
}

}  // namespace gode
