// tc_rk4_adj_wide.cu — tensor-core continuous adjoint of the fixed-grid RK4 (3/8 rule) solve for the wide field
// D=64, H=256 (BF16 operands, FP32 accumulate in TMEM).  torchdiffeq semantics (adjoint.py): per output
// interval, from the last to the first, ONE 3/8 step of the augmented system (y, a, theta_bar) in reversed time starting
// from the STORED y_i; a += grad_y[i-1] afterwards.
//
// Every contraction of an augmented stage runs on tcgen05, all six of them off ONE set of shared-memory tiles, read
// through K-major or MN-major ("transposed") descriptors as needed (scripts/mn_major_test.cu checks that view):
//     z1   = [u|1] [W1|b1]^T        A = A1 (K-major)      B = B1 (K-major)            N = 2 x 128
//     f    = h W2^T                 A = HD (K-major)      B = B2 (K-major)            N = 64        (h = tanh z1)
//     g'   = (c a) W2               A = AA (K-major)      B = B2 (N-major view)       N = 2 x 128
//     v'   = d' W1                  A = HD (K-major)      B = B1 (N-major view)       N = 64        (d' = g' (1 - h^2))
//     dW2^T += h^T (c a)            A = HD (M-major view) B = AA (N-major view)       M = 2 x 128, N = 64, K = 128 rows
//     [dW1|db1] += d'^T [u|1]       A = HD (M-major view) B = A1 (N-major view)       M = 2 x 128, N = 80, K = 128 rows
// c = the stage's quadrature weight (dt/8, 3dt/8): carrying it on a makes g', d', v' come out scaled, so theta_bar
// needs no second copy of any operand (v = v'/c in the epilogue).  The two weight-gradient accumulators live in TMEM
// (288 columns) for the WHOLE kernel: all tiles, intervals and stages of a CTA add into them, and they are read once at
// the end -> per-CTA partial rows -> fixed-order reduction (deterministic).  db2 = sum c a is kept in registers.
//
// One CTA = one 128-trajectory tile at a time = 256 threads (two per row: TMEM lane = row, column half = warp / 4).
// TMEM: [0,128) z1 half / g' half / v' ; [128,192) f ; [224,352) dW2^T (2 x 64) ; [352,512) [dW1|db1] (2 x 80).
// Shared: B1 40 KB, B2 32 KB, A1 2 x 20 KB (double buffered: stage s+1's input is packed while stage s's is still the B
// operand of the dW1 MMAs), AA 16 KB, HD 64 KB (h, then d' in place) = 193 KB.
#include "launch.h"
#include "tc_common.cuh"

namespace gode {

namespace {

constexpr float kThird = 0.33333334f;
constexpr int D = 64, H = 256, TILE = 128;
constexpr int KC1 = D / 8 + 2;                 // A1/B1 K chunks: data + (1,1,0..) chunk carrying b1 + zero chunk
constexpr int OFF_BIAS = 0;                    // fp32 b1 | b2
constexpr int OFF_B1 = 1280;
constexpr int OFF_B2 = OFF_B1 + H * KC1 * 16;
constexpr int A1_BYTES = TILE * KC1 * 16;
constexpr int OFF_A1 = OFF_B2 + D * H * 2;
constexpr int OFF_AA = OFF_A1 + 2 * A1_BYTES;
constexpr int OFF_HD = OFF_AA + TILE * D * 2;  // also the fp32 staging area of one weight matrix in the prologue
constexpr int OFF_RED = OFF_HD + TILE * H * 2; // 4 x 64 floats for the final db2 reduction
constexpr int OFF_BAR = OFF_RED + 4 * D * 4;
constexpr int SMEM_BYTES = OFF_BAR + 64;
constexpr uint32_t T_ZG = 0, T_F = 128, T_W2T = 224, T_W1E = 352;
constexpr int NE = D + 16;                     // columns of the [dW1|db1] accumulator
constexpr int PART = H * NE + H * D + D;       // floats per CTA partial: [dW1|db1] (H x NE), dW2^T (H x D), db2 (D)
constexpr uint32_t A_MN = 1u << 15, B_MN = 1u << 16;

struct AdjArgs {
  const float *traj, *grad_traj, *W1, *b1, *W2, *b2;
  float* grad_y0;
  float* partial;
  const float* dt_dev;
  int B, T, layout;
  float dt_val[GODE_MAX_HOST_STEPS];
};

__device__ __forceinline__ size_t off3(int layout, int s, int b, int B, int T) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}
__device__ __forceinline__ uint64_t adv(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xFFFF0000u); }

__global__ void __launch_bounds__(256, 1) tc_rk4_adj_wide_kernel(const __grid_constant__ AdjArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  unsigned char* B1 = smem + OFF_B1;
  unsigned char* B2 = smem + OFF_B2;
  unsigned char* AA = smem + OFF_AA;
  unsigned char* HD = smem + OFF_HD;
  float* stagef = reinterpret_cast<float*>(HD);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* mbar_m = mbar_w + 1;
  uint64_t* mbar_b = mbar_w + 2;  // completion of the dW2^T MMAs (they read HD, which the d' epilogue overwrites)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_w + 3);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform (keeps MMA descriptors in uniform registers)
  const int q = warp & 3, hf = warp >> 2, row = q * 32 + lane;

  if (warp == 0) tc::tmem_alloc(s_tmem, 512);
  if (tid == 0) {
    tc::mbar_init(mbar_w, 1);
    tc::mbar_init(mbar_m, 1);
    tc::mbar_init(mbar_b, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const uint32_t my_t = tmem + ((uint32_t)(q * 32) << 16);

  // ---- weights -> BF16 operand tiles (once per CTA) ----
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, H * D * 4 + (H + D) * 4);
    tc::tma_bulk_g2s(stagef, p.W1, H * D * 4, mbar_w);
    tc::tma_bulk_g2s(bias, p.b1, H * 4, mbar_w);
    tc::tma_bulk_g2s(bias + H, p.b2, D * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 0);
  for (int idx = tid; idx < H * (D / 8); idx += 256) {  // B1[kc][j][8 d] <- W1[j][8kc..]
    const int n = idx % H, kc = idx / H;
    const float* v = stagef + n * D + kc * 8;
    *reinterpret_cast<uint4*>(B1 + (size_t)(kc * H + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  for (int n = tid; n < H; n += 256) {  // K = D, D+1: b1 as two bf16 terms against the (1,1) columns of A1
    const float b = bias[n];
    const float hi = __bfloat162float(__float2bfloat16_rn(b));
    *reinterpret_cast<uint4*>(B1 + (size_t)((D / 8) * H + n) * 16) = make_uint4(tc::pack_bf16x2(hi, b - hi), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(B1 + (size_t)((D / 8 + 1) * H + n) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  {  // constant K chunks of both A1 buffers: (1, 1, 0, ...) and zeros; column D of [u|1,1] is also the db1 column of dW1
    const int g = tid >> 7, r = tid & 127;
    unsigned char* a1 = smem + OFF_A1 + g * A1_BYTES;
    *reinterpret_cast<uint4*>(a1 + (size_t)((D / 8) * TILE + r) * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(a1 + (size_t)((D / 8 + 1) * TILE + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc::fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, D * H * 4);
    tc::tma_bulk_g2s(stagef, p.W2, D * H * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 1);
  for (int idx = tid; idx < D * (H / 8); idx += 256) {  // B2[kc][d][8 j] <- W2[d][8kc..]
    const int n = idx % D, kc = idx / D;
    const float* v = stagef + n * H + kc * 8;
    *reinterpret_cast<uint4*>(B2 + (size_t)(kc * D + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  tc::fence_async_smem();
  __syncthreads();

  const uint32_t sB1 = tc::smem_u32(B1), sB2 = tc::smem_u32(B2), sAA = tc::smem_u32(AA), sHD = tc::smem_u32(HD);
  const uint32_t sA1 = tc::smem_u32(smem + OFF_A1);
  const float* b2s = bias + H + hf * 32;
  constexpr uint32_t id_z = tc::make_idesc(tc::kFmtBF16, 128, 128);
  constexpr uint32_t id_f = tc::make_idesc(tc::kFmtBF16, 128, 64);
  constexpr uint32_t id_g = tc::make_idesc(tc::kFmtBF16, 128, 128) | B_MN;
  constexpr uint32_t id_v = tc::make_idesc(tc::kFmtBF16, 128, 64) | B_MN;
  constexpr uint32_t id_w2 = tc::make_idesc(tc::kFmtBF16, 128, 64) | A_MN | B_MN;
  constexpr uint32_t id_w1 = tc::make_idesc(tc::kFmtBF16, 128, NE) | A_MN | B_MN;
  uint32_t phase = 0, phase_b = 0;
  uint32_t acc_live = 0;  // 0 until the weight-gradient accumulators have been written once

  // operands written by all threads -> one elected lane issues (the lambda ends with the commit(s) it wants) -> everybody
  // waits for mbar_m.  MMAs issued AFTER the commit to mbar_m keep running under the epilogue that follows; the tensor
  // pipe completes MMAs in issue order, so any later completed commit also covers them.
  auto sync_issue = [&](auto&& issue) {
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0 && tc::elect_one()) {
      tc::fence_after_sync();
      issue();
    }
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
  };
  // z1 half h2: A = [u|1] (5 K steps of 16), B = rows [128 h2, +128) of [W1|b1]
  auto issue_z = [&](int cur, int h2) {
    const uint64_t dA = tc::make_smem_desc(sA1 + cur * A1_BYTES, TILE * 16, 128);
    const uint64_t dB = tc::make_smem_desc(sB1 + h2 * 128 * 16, H * 16, 128);
#pragma unroll
    for (int k = 0; k < KC1 / 2; ++k) tc::mma_ss<false>(tmem + T_ZG, adv(dA, k * 2 * TILE * 16), adv(dB, k * 2 * H * 16), id_z, k > 0);
  };
  // g' half h2 = (c a) W2[:, 128 h2 ..]: B = B2 read N-major (N = j: chunk stride D*16, K = d: 8-row groups 128 B apart)
  auto issue_g = [&](int h2) {
    const uint64_t dA = tc::make_smem_desc(sAA, TILE * 16, 128);
    const uint64_t dB = tc::make_smem_desc(sB2 + h2 * 16 * D * 16, 128, D * 16);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) tc::mma_ss<false>(tmem + T_ZG, adv(dA, k * 2 * TILE * 16), adv(dB, k * 256), id_g, k > 0);
  };

  float dbb[32];  // db2 partial of this thread's row and column half
#pragma unroll
  for (int i = 0; i < 32; ++i) dbb[i] = 0.f;

  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int ntiles = (p.B + TILE - 1) / TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * TILE + row;
    const bool valid = b < p.B;
    float Y[32], KY[32], A[32], WA[32];
    // pack 32 fp32 values of this thread into its 4 chunks of a [kc][row][16 B] tile
    auto pack32 = [&](unsigned char* base, const float(&v)[32], float scale) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(base + (size_t)((hf * 4 + c) * TILE + row) * 16) =
            make_uint4(tc::pack_bf16x2(scale * v[8 * c], scale * v[8 * c + 1]), tc::pack_bf16x2(scale * v[8 * c + 2], scale * v[8 * c + 3]),
                       tc::pack_bf16x2(scale * v[8 * c + 4], scale * v[8 * c + 5]), tc::pack_bf16x2(scale * v[8 * c + 6], scale * v[8 * c + 7]));
    };
    auto load32 = [&](const float* src, float(&v)[32]) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 t4 = valid ? *reinterpret_cast<const float4*>(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
      }
    };
    load32(p.grad_traj + (valid ? off3(p.layout, p.T - 1, b, p.B, p.T) : 0) + hf * 32, A);
    int cur = 0;  // A1 buffer holding the current stage's [u|1]
    for (int i = p.T - 1; i >= 1; --i) {
      const float dt = dtp[i - 1];
      const float cs[4] = {dt * 0.125f, dt * 0.375f, dt * 0.375f, dt * 0.125f};
      load32(p.traj + (valid ? off3(p.layout, i, b, p.B, p.T) : 0) + hf * 32, Y);
      pack32(smem + OFF_A1 + cur * A1_BYTES, Y, 1.f);
      pack32(AA, A, cs[0]);
#pragma unroll
      for (int j = 0; j < 32; ++j) dbb[j] += cs[0] * A[j];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float c = cs[s], rc = 1.f / c;
        unsigned char* A1n = smem + OFF_A1 + (cur ^ 1) * A1_BYTES;
        // ---- z1 -> h (two halves of 128 hidden units) ----
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          sync_issue([&] { issue_z(cur, h2); tc::mma_commit(mbar_m); });
#pragma unroll
          for (int cc = 0; cc < 4; cc += 2) {
            uint32_t za[16], zb[16];
            tc::tmem_ld16_nowait(my_t + T_ZG + hf * 64 + cc * 16, za);
            tc::tmem_ld16_nowait(my_t + T_ZG + hf * 64 + cc * 16 + 16, zb);
            tc::tmem_ld_wait();
            uint32_t qa[8], qb[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              qa[e / 2] = tc::pack_bf16x2(tc::tanh_approx(__uint_as_float(za[e])), tc::tanh_approx(__uint_as_float(za[e + 1])));
              qb[e / 2] = tc::pack_bf16x2(tc::tanh_approx(__uint_as_float(zb[e])), tc::tanh_approx(__uint_as_float(zb[e + 1])));
            }
            const int kc = (h2 * 128 + hf * 64 + cc * 16) / 8;
            *reinterpret_cast<uint4*>(HD + (size_t)(kc * TILE + row) * 16) = make_uint4(qa[0], qa[1], qa[2], qa[3]);
            *reinterpret_cast<uint4*>(HD + (size_t)((kc + 1) * TILE + row) * 16) = make_uint4(qa[4], qa[5], qa[6], qa[7]);
            *reinterpret_cast<uint4*>(HD + (size_t)((kc + 2) * TILE + row) * 16) = make_uint4(qb[0], qb[1], qb[2], qb[3]);
            *reinterpret_cast<uint4*>(HD + (size_t)((kc + 3) * TILE + row) * 16) = make_uint4(qb[4], qb[5], qb[6], qb[7]);
          }
        }
        // ---- f = h W2^T (not needed after the last stage), g' half 0; then dW2^T += h^T (c a) under the y epilogue ----
        sync_issue([&] {
          if (s < 3) {
            const uint64_t dA = tc::make_smem_desc(sHD, TILE * 16, 128);
            const uint64_t dB = tc::make_smem_desc(sB2, D * 16, 128);
#pragma unroll
            for (int k = 0; k < H / 16; ++k) tc::mma_ss<false>(tmem + T_F, adv(dA, k * 2 * TILE * 16), adv(dB, k * 2 * D * 16), id_f, k > 0);
          }
          issue_g(0);
          tc::mma_commit(mbar_m);
          const uint64_t dBa = tc::make_smem_desc(sAA, 128, TILE * 16);
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint64_t dAh = tc::make_smem_desc(sHD + mh * 16 * TILE * 16, 128, TILE * 16);
#pragma unroll
            for (int k = 0; k < TILE / 16; ++k)
              tc::mma_ss<false>(tmem + T_W2T + mh * D, adv(dAh, k * 256), adv(dBa, k * 256), id_w2, acc_live | (uint32_t)(k > 0));
          }
          tc::mma_commit(mbar_b);
        });
        // y part of the 3/8 step in reversed time: ky = -f.  Packs the next stage's u into the other A1 buffer.
        if (s < 3) {
          uint32_t za[16], zb[16];
          tc::tmem_ld16_nowait(my_t + T_F + hf * 32, za);
          tc::tmem_ld16_nowait(my_t + T_F + hf * 32 + 16, zb);
          tc::tmem_ld_wait();
          float un[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float ky = -(__uint_as_float(e < 16 ? za[e] : zb[e - 16]) + b2s[e]);
            if (s == 0) { KY[e] = ky; un[e] = Y[e] + dt * kThird * ky; }
            if (s == 1) { un[e] = Y[e] + dt * (ky - KY[e] * kThird); KY[e] = Y[e] + dt * (KY[e] - ky); }
            if (s == 2) { un[e] = KY[e] + dt * ky; }
          }
          pack32(A1n, un, 1.f);
        }
        // ---- d' = g' (1 - h^2), written over h (once dW2^T has read it); second g' half in between ----
        tc::mbar_wait(mbar_b, phase_b);
        phase_b ^= 1;
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
          if (h2 == 1) sync_issue([&] { issue_g(1); tc::mma_commit(mbar_m); });
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t g16[16];
            tc::tmem_ld16_nowait(my_t + T_ZG + hf * 64 + cc * 16, g16);
            const int kc = (h2 * 128 + hf * 64 + cc * 16) / 8;
            uint4* p0 = reinterpret_cast<uint4*>(HD + (size_t)(kc * TILE + row) * 16);
            uint4* p1 = reinterpret_cast<uint4*>(HD + (size_t)((kc + 1) * TILE + row) * 16);
            const uint4 h0 = *p0, h1 = *p1;
            tc::tmem_ld_wait();
            const uint32_t hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float ha = bf_lo(hh[e]), hb = bf_hi(hh[e]);
              o[e] = tc::pack_bf16x2(__uint_as_float(g16[2 * e]) * (1.f - ha * ha), __uint_as_float(g16[2 * e + 1]) * (1.f - hb * hb));
            }
            *p0 = make_uint4(o[0], o[1], o[2], o[3]);
            *p1 = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        // ---- v' = d' W1; then [dW1|db1] += d'^T [u|1], which nothing waits for: it runs under the a epilogue and the next
        // stage's packing, and the next commit (z1 of the next stage, or the final one) covers it ----
        sync_issue([&] {
          const uint64_t dA = tc::make_smem_desc(sHD, TILE * 16, 128);
          const uint64_t dB = tc::make_smem_desc(sB1, 128, H * 16);
#pragma unroll
          for (int k = 0; k < H / 16; ++k) tc::mma_ss<false>(tmem + T_ZG, adv(dA, k * 2 * TILE * 16), adv(dB, k * 256), id_v, k > 0);
          tc::mma_commit(mbar_m);
          const uint64_t dBu = tc::make_smem_desc(sA1 + cur * A1_BYTES, 128, TILE * 16);
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint64_t dAh = tc::make_smem_desc(sHD + mh * 16 * TILE * 16, 128, TILE * 16);
#pragma unroll
            for (int k = 0; k < TILE / 16; ++k)
              tc::mma_ss<false>(tmem + T_W1E + mh * NE, adv(dAh, k * 256), adv(dBu, k * 256), id_w1, acc_live | (uint32_t)(k > 0));
          }
        });
        acc_live = 1;
        // a part: ka = v = v'/c.  Packs the next stage's (c a) into AA (all MMAs that read AA have completed).
        {
          uint32_t za[16], zb[16];
          tc::tmem_ld16_nowait(my_t + T_ZG + hf * 32, za);
          tc::tmem_ld16_nowait(my_t + T_ZG + hf * 32 + 16, zb);
          tc::tmem_ld_wait();
          float an[32];
          float gp[32];
          if (s == 3) load32(p.grad_traj + (valid ? off3(p.layout, i - 1, b, p.B, p.T) : 0) + hf * 32, gp);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float ka = __uint_as_float(e < 16 ? za[e] : zb[e - 16]) * rc;
            if (s == 0) { WA[e] = ka; an[e] = A[e] + dt * kThird * ka; }
            if (s == 1) {
              an[e] = A[e] + dt * (ka - WA[e] * kThird);
              const float P = A[e] + dt * 0.125f * (WA[e] + 3.f * ka);
              WA[e] = A[e] + dt * (WA[e] - ka);
              A[e] = P;
            }
            if (s == 2) { an[e] = WA[e] + dt * ka; A[e] += dt * 0.375f * ka; }
            if (s == 3) { A[e] += dt * 0.125f * ka + gp[e]; }
          }
          if (s < 3) {
            pack32(AA, an, cs[s + 1]);
#pragma unroll
            for (int e = 0; e < 32; ++e) dbb[e] += cs[s + 1] * an[e];
          }
        }
        cur ^= 1;
      }
    }
    if (valid) {
      float* o = p.grad_y0 + (size_t)b * D + hf * 32;
#pragma unroll
      for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(A[e], A[e + 1], A[e + 2], A[e + 3]);
    }
  }

  // ---- this CTA's partial: accumulators out of TMEM + db2 ----
  float* part = p.partial + (size_t)blockIdx.x * PART;
  sync_issue([&] { tc::mma_commit(mbar_m); });  // drains the last [dW1|db1] MMAs
#pragma unroll 1
  for (int mh = 0; mh < 2; ++mh) {
    const int j = mh * 128 + row;
    // [dW1|db1]: NE = 80 columns; column half hf takes 40 of them (16 + 16 + 8)
#pragma unroll 1
    for (int c0 = hf * 40; c0 < hf * 40 + 40; c0 += 8) {
      uint32_t r8[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7])
                   : "r"(my_t + T_W1E + mh * NE + c0)
                   : "memory");
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 8; ++e) part[(size_t)j * NE + c0 + e] = acc_live ? __uint_as_float(r8[e]) : 0.f;
    }
#pragma unroll 1
    for (int c0 = hf * 32; c0 < hf * 32 + 32; c0 += 16) {
      float v[16];
      tc::tmem_ld16(my_t + T_W2T + mh * D + c0, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) part[(size_t)H * NE + (size_t)j * D + c0 + e] = acc_live ? v[e] : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    float v = dbb[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    dbb[e] = v;
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 32; ++e) red[q * D + hf * 32 + e] = dbb[e];
  }
  __syncthreads();
  if (tid < D) part[(size_t)H * NE + (size_t)H * D + tid] = (red[tid] + red[D + tid]) + (red[2 * D + tid] + red[3 * D + tid]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// flat gradient [W1 (H x D) | b1 (H) | W2 (D x H) | b2 (D)] = sum over CTA partials in CTA order
__global__ void tc_adj_reduce_kernel(const float* __restrict__ partial, int slices, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int n1 = H * D, n2 = n1 + H, n3 = n2 + D * H, n4 = n3 + D;
  if (e >= n4) return;
  int src;
  if (e < n1) src = (e / D) * NE + (e % D);                          // W1[j][d]  <- [dW1|db1][j][d]
  else if (e < n2) src = (e - n1) * NE + D;                          // b1[j]     <- [dW1|db1][j][D]
  else if (e < n3) src = H * NE + ((e - n2) % H) * D + (e - n2) / H;  // W2[d][j]  <- dW2^T[j][d]
  else src = H * NE + H * D + (e - n3);                              // b2[d]
  float s = 0.f;
  for (int k = 0; k < slices; ++k) s += partial[(size_t)k * PART + src];
  out[e] = s;
}

int adj_grid(int B) {
  const int ntiles = (B + TILE - 1) / TILE;
  const int g = sm_count();
  return g < ntiles ? g : ntiles;
}

}  // namespace

size_t tc_rk4_adj_wide_workspace_bytes(int B) { return sizeof(float) * (size_t)PART * (size_t)sm_count() + 256; }

int tc_rk4_adj_wide(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* dt, int dt_on_device, int B, int Dd, int Hh, int T, int layout,
                    float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (!(Dd == D && Hh == H)) return GODE_ERR_SHAPE;
  const int grid = adj_grid(B);
  if (ws_bytes < sizeof(float) * (size_t)PART * (size_t)grid) return GODE_ERR_WORKSPACE;
  AdjArgs a{};
  a.traj = traj; a.grad_traj = grad_traj; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
  a.grad_y0 = grad_y0; a.partial = reinterpret_cast<float*>(workspace);
  a.B = B; a.T = T; a.layout = layout;
  if (dt_on_device) {
    a.dt_dev = dt;
  } else {
    if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
    for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  }
  cudaError_t e = cudaFuncSetAttribute(tc_rk4_adj_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
  tc_rk4_adj_wide_kernel<<<grid, 256, SMEM_BYTES, st>>>(a);
  const int n = H * D + H + D * H + D;
  tc_adj_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(a.partial, grid, grad_params);
  return launch_status();
}

}  // namespace gode
