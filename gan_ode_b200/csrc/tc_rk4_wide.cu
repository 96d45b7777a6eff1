// tc_rk4_wide.cu — tensor-core RK4 (3/8 rule) forward for the wide field D=64, H=256 (BF16 operands, FP32 accumulate).
//
// This is the shape where the contraction dominates (1 MFLOP per trajectory-step, arithmetic intensity ~1365 FLOP/B):
//   layer 1: 128 x 256 x (64+16)  -> 5  tcgen05.mma (M=128, N=256, K=16), the 5th K step carries b1
//   layer 2: 128 x 64 x 256       -> 16 tcgen05.mma (M=128, N=64,  K=16) with A = h read from TMEM
// One 128-trajectory tile = 256 threads, two per trajectory row (TMEM lane): warp w reads lane quadrant w%4 and column
// half w/4, so every Runge–Kutta vector is 32 registers per thread.  Weights (132 KB fp32) arrive by TMA bulk copies
// through a staging area and are re-tiled to BF16 UMMA layout once per CTA; persistent over tiles.
// (An earlier single-tile kernel with h staged through shared memory — 357 TFLOP/s, tensor pipe 17 % — was replaced by
// the two-group kernel below; with one group idle it is also the faster one for <= 148 tiles: 182 vs 192 us at B = 4096.)
#include <stdio.h>
#include <stdlib.h>
#include "launch.h"
#include "tc_common.cuh"

namespace gode {

constexpr float kWT = 0.33333334f;

struct TcWideArgs {
  const float *y0, *W1, *b1, *W2, *b2;
  float* traj;
  const float* dt_dev;
  int B, T, layout;
  long long* dbg;  // developer phase timing (GODE_TCW_DBG=1), normally null
  float dt_val[GODE_MAX_HOST_STEPS];
};

__device__ __forceinline__ size_t tcw_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

// ---- two tiles in flight per SM ------------------------------------------------------------------------------------------------
// A single resident tile runs MMA -> epilogue -> MMA serially (ncu of that first kernel: tensor pipe 17 %, XU 34 %).
// Binding resource of this shape: tanh.  One stage of one 128-row tile is 128 x 256 tanh = 2048 cycles of the SM's
// 16-lane MUFU pipe (measured: scripts/tmem_bench.cu, 15.98 tanh/clk/SM, bf16x2 is two MUFU ops) against 1024 cycles of
// tcgen05 MMA, so the design goal is "MUFU never idle":
//   * one CTA = 512 threads = two independent 256-thread groups, each with its own tile, TMEM half (256 columns), A1
//     buffer, named barrier and MMA mbarrier;
//   * the groups ALTERNATE their layer-1 (tanh) epilogues through a pair of 512-thread named barriers used as a token,
//     so one group's MMAs, layer-2 epilogue and Runge–Kutta algebra run under the other group's tanh;
//   * h = tanh(z1 + b1) never touches shared memory: it is written back to TMEM (tcgen05.st, packed bf16 pairs) over z1
//     columns its thread has already consumed, and layer 2 takes its A operand from TMEM;
//   * layer 2 is split into two N=32 halves and its first eight K steps are issued in the MIDDLE of the tanh epilogue;
//   * TMEM loads are software-pipelined one 16-column chunk ahead;
//   * the 3/8-rule algebra is regrouped so at most TWO D-vectors are live across an f evaluation
//     (P = y + dt/8 (k1 + 3 k2) and w = y + dt (k1 - k2) replace y, acc, v) and is applied per 16-column chunk of the
//     layer-2 epilogue, which also packs the next stage's input straight into A1: 128 registers per thread, no spills.
// TMEM columns of a group, per column half hb = 128*hf (thread (row, hf) owns lane `row`, columns hb..hb+127):
//   z1 chunk c (16 fp32 columns)  hb + 16c                     c = 0..7
//   h  chunk c (8 packed columns) hb + 8c (c < 4), hb + 64 + 8(c-4) (c >= 4)   [written after z1 chunk c was read]
//   layer-2 accumulator, N half hf (32 fp32 columns)  hb + 32  [free once z1 chunks 2,3 are consumed]
template <int D, int H>
struct TcWide2Shape {
  static constexpr int TILE = 128;
  static constexpr int KC1 = D / 8 + 2;  // + one K step that carries b1: A column of ones x (b1_hi, b1_lo) rows of B1
  static constexpr int NM1 = KC1 / 2;
  static constexpr int DH = D / 2, HH = H / 2;
  static constexpr int OFF_BIAS = 0;
  static constexpr int OFF_B1 = ((H + D) * 4 + 127) / 128 * 128;
  static constexpr int OFF_B2 = OFF_B1 + H * KC1 * 16;
  static constexpr int OFF_A1 = OFF_B2 + D * H * 2;               // two groups
  static constexpr int A1_BYTES = TILE * KC1 * 16;
  static constexpr int OFF_STAGE = OFF_A1 + 2 * A1_BYTES;           // fp32 staging of one weight matrix
  static constexpr int OFF_BAR = OFF_STAGE + H * D * 4;
  static constexpr int BYTES = OFF_BAR + 64;
  static_assert(H == 256 && D == 64, "TMEM column plan is written for H=256, D=64");
};

__device__ __forceinline__ uint32_t tcw2_hcol(int c) { return 8u * c; }
#ifndef GODE_TANH_FMA_EVERY
#define GODE_TANH_FMA_EVERY 1000
#endif
// Measured (round 2, scripts/tanh_poly_probe.py, two runs; B = 18 944 / 151 552, us): pairs of every eight on the FMA pipe
//   0: 182.7 / 815.5    1: 178.2 / 784.8    2: 176.6 / 798.2    3: 174.4 / 836.0    4: 180.7 / 870.8
// error against the FP32 solve 3.4e-4 -> 3.5e-4 (N = 1).  The epilogue is not bound by MUFU throughput alone: with two epilogue
// warps per scheduler the twelve extra issue slots of a polynomial pair are only partly hidden; one pair in eight is the best
// trade in the throughput regime (759 TFLOP/s = 0.56 of the sustained BF16 GEMM rate), three at one tile per SM.
constexpr int kTanhPolyDefault = 1;   // GODE_TANH_POLY_N default: pairs of every eight on the FMA pipe (0: none)
constexpr int kTanhFmaEvery = GODE_TANH_FMA_EVERY;  // one pair in every kTanhFmaEvery goes to the FMA pipe (1000: none).
// Measured (B = 151 552): none 844 us, every 4th pair 858 us, every 2nd pair 930 us -- the extra ~11 issue slots per pair
// cost more than the MUFU cycles they free with only two epilogue warps per scheduler, so the default is OFF.

// POLY_N: of the eight tanh pairs of a 16-column chunk, this many are evaluated by tc::tanh_pair_poly on the FMA pipe (no MUFU op)
template <int D, int H, int POLY_N>
__global__ void __launch_bounds__(512, 1) tc_rk4_fwd_wide2_kernel(const __grid_constant__ TcWideArgs p) {
  using S = TcWide2Shape<D, H>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* bias = reinterpret_cast<float*>(smem + S::OFF_BIAS);
  unsigned char* B1 = smem + S::OFF_B1;
  unsigned char* B2 = smem + S::OFF_B2;
  float* stagef = reinterpret_cast<float*>(smem + S::OFF_STAGE);
  uint64_t* mbar_w = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* mbar_g = mbar_w + 1;  // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_w + 3);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction (the compiler can rely on it)
  const int grp = warp >> 3, gtid = tid & 255, gwarp = warp & 7;
  const bool issuer_warp = gwarp == 0;
  const int row = (gwarp & 3) * 32 + lane, hf = gwarp >> 2;
  unsigned char* A1 = smem + S::OFF_A1 + grp * S::A1_BYTES;

  if (warp == 0) tc::tmem_alloc(s_tmem, 512);
  if (tid == 0) {
    tc::mbar_init(mbar_w, 1);
    tc::mbar_init(mbar_g, 1);
    tc::mbar_init(mbar_g + 1, 1);
    tc::mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0) + (uint32_t)(grp * 256);                              // this group's column base
  const uint32_t my_tmem = tmem + ((uint32_t)((gwarp & 3) * 32) << 16) + hf * S::HH;  // + lane quadrant + column half

  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, H * D * 4 + (H + D) * 4);
    tc::tma_bulk_g2s(stagef, p.W1, H * D * 4, mbar_w);
    tc::tma_bulk_g2s(bias, p.b1, H * 4, mbar_w);
    tc::tma_bulk_g2s(bias + H, p.b2, D * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 0);
  for (int idx = tid; idx < H * (D / 8); idx += 512) {
    const int n = idx % H, kc = idx / H;
    const float* v = stagef + n * D + kc * 8;
    *reinterpret_cast<uint4*>(B1 + (size_t)(kc * H + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  for (int n = tid; n < H; n += 512) {  // K = D, D+1: b1 split into two bf16 terms (exact to 2^-17), against A columns of ones
    const float b = bias[n];
    const float hi = __bfloat162float(__float2bfloat16_rn(b));
    *reinterpret_cast<uint4*>(B1 + (size_t)((D / 8) * H + n) * 16) = make_uint4(tc::pack_bf16x2(hi, b - hi), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(B1 + (size_t)((D / 8 + 1) * H + n) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 256) {  // the constant K step of both groups' A1 tiles: (1, 1, 0, ...)
    const int g = tid >> 7, r = tid & 127;
    unsigned char* a1 = smem + S::OFF_A1 + g * S::A1_BYTES;
    *reinterpret_cast<uint4*>(a1 + (size_t)((D / 8) * S::TILE + r) * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(a1 + (size_t)((D / 8 + 1) * S::TILE + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  tc::fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tc::mbar_expect_tx(mbar_w, D * H * 4);
    tc::tma_bulk_g2s(stagef, p.W2, D * H * 4, mbar_w);
  }
  tc::mbar_wait(mbar_w, 1);
  for (int idx = tid; idx < D * (H / 8); idx += 512) {
    const int n = idx % D, kc = idx / D;
    const float* v = stagef + n * H + kc * 8;
    *reinterpret_cast<uint4*>(B2 + (size_t)(kc * D + n) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
  tc::fence_async_smem();
  __syncthreads();

  constexpr uint32_t idesc1 = tc::make_idesc(tc::kFmtBF16, 128, H);
  constexpr uint32_t idesc2 = tc::make_idesc(tc::kFmtBF16, 128, D);
  const uint32_t sA1 = tc::smem_u32(A1), sB1 = tc::smem_u32(B1), sB2 = tc::smem_u32(B2);  // descriptors are rebuilt by the
  const float* b2s = bias + H + hf * S::DH;
  uint64_t* mbar_m = mbar_g + grp;
  const uint32_t bar_id = 1 + grp, tok_mine = 3 + grp, tok_other = 3 + (grp ^ 1);
  uint32_t phase = 0;

  // one 16-column chunk of the layer-1 epilogue (b1 is already inside z1): tanh, pack, store over consumed z1 columns
  auto tanh_chunk = [&](const uint32_t(&z)[16], int c) {
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if ((i / 2) % kTanhFmaEvery == kTanhFmaEvery - 1) {  // this pair on the FMA pipe + one MUFU.RCP (tc_common.cuh)
        const float2 th = tc::tanh_pair_fma(__uint_as_float(z[i]), __uint_as_float(z[i + 1]));
        q[i / 2] = tc::pack_bf16x2(th.x, th.y);
      } else if ((i / 2) % 2 == 1 ? (i / 4) < POLY_N : (i / 4) + 4 < POLY_N) {  // this pair on the FMA pipe alone (polynomial):
        // the odd pairs first, so that polynomial and MUFU pairs alternate in the instruction stream
        const float2 th = tc::tanh_pair_poly(__uint_as_float(z[i]), __uint_as_float(z[i + 1]));
        q[i / 2] = tc::pack_bf16x2(th.x, th.y);
      } else {
        q[i / 2] = tc::pack_bf16x2(tc::tanh_approx(__uint_as_float(z[i])), tc::tanh_approx(__uint_as_float(z[i + 1])));
      }
    }
    tc::tmem_st8(my_tmem + tcw2_hcol(c), q);
  };
  // layer-2 MMAs: 16 K steps, A = packed h in TMEM (half j/8, 8 columns per step), D = this group's columns [64,128)
  auto issue_layer2 = [&]() {
    const uint64_t dB2 = tc::make_smem_desc(sB2, D * 16, 128);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      tc::mma_ts_bf16(tmem + 64, tmem + (j / 8) * S::HH + tcw2_hcol(j % 8), dB2 + (uint64_t)((2 * j * D * 16) >> 4), idesc2, j > 0);
  };

  // One f evaluation of the group's tile; `upd(i, k_i)` consumes element i of the result and returns element i of the
  // NEXT stage's input, which is packed into A1 on the spot.  A1 must hold the current input on entry.
#ifdef GODE_TCW_TIMING  // developer build: per-phase clock64 accounting (see tc_rk4_fwd_wide), costs registers
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nfe = 0, ev = 0;
  long long tlast = 0;
#define TCW_T(i) do { if (p.dbg) { const long long _n = clock64(); tph[i] += _n - tlast; tlast = _n; if (blockIdx.x == 0 && gtid == 0 && nfe >= 16 && nfe < 24) p.dbg[64 + (grp * 8 + (nfe - 16)) * 8 + ev++] = _n; } } while (0)
#define TCW_BEGIN() do { if (p.dbg) { tlast = clock64(); ev = 0; } } while (0)
#define TCW_END() do { ++nfe; } while (0)
#else
#define TCW_T(i) do { } while (0)
#define TCW_BEGIN() do { } while (0)
#define TCW_END() do { } while (0)
#endif
  auto feval = [&](bool last_token, auto&& upd) {
    TCW_BEGIN();
    tc::fence_async_smem();
    tc::fence_before_sync();
    tc::named_bar_sync(bar_id, 256);
    if (issuer_warp && tc::elect_one()) {
      tc::fence_after_sync();
      const uint64_t dA1 = tc::make_smem_desc(sA1, S::TILE * 16, 128);
      const uint64_t dB1 = tc::make_smem_desc(sB1, H * 16, 128);
#pragma unroll
      for (int j = 0; j < S::NM1; ++j)
        tc::mma_ss<false>(tmem, dA1 + (uint64_t)((2 * j * S::TILE * 16) >> 4), dB1 + (uint64_t)((2 * j * H * 16) >> 4), idesc1, j > 0);
      tc::mma_commit(mbar_m);
    }
    TCW_T(0);
    tc::mbar_wait(mbar_m, phase);       // normally long complete: MMA1 ran under the other group's tanh epilogue
    phase ^= 1;
    tc::fence_after_sync();
    uint32_t za[16], zb[16];
    tc::tmem_ld16_nowait(my_tmem, za);
    tc::tmem_ld_wait();
    TCW_T(1);
    tc::named_bar_sync(tok_mine, 512);  // the other group has finished its tanh epilogue: the MUFU pipe is ours
    TCW_T(2);
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      tc::tmem_ld16_nowait(my_tmem + 16 * (c + 1), zb);
      tanh_chunk(za, c);
      tc::tmem_ld_wait();
      if (c + 2 < 8) tc::tmem_ld16_nowait(my_tmem + 16 * (c + 2), za);
      tanh_chunk(zb, c + 1);
      tc::tmem_ld_wait();
    }
    TCW_T(3);
    if (!last_token) tc::named_bar_arrive(tok_other, 512);  // hand the MUFU pipe to the other group (our last tanh ops are
    tc::tmem_st_wait();                                     // issued; the stores drain while the other group starts)
    tc::fence_before_sync();
    tc::named_bar_sync(bar_id, 256);
    if (issuer_warp && tc::elect_one()) {
      tc::fence_after_sync();
      issue_layer2();
      tc::mma_commit(mbar_m);
    }
    TCW_T(4);
    tc::mbar_wait(mbar_m, phase);
    phase ^= 1;
    tc::fence_after_sync();
    TCW_T(5);
    tc::tmem_ld16_nowait(my_tmem - hf * S::HH + 64 + hf * S::DH, za);
    tc::tmem_ld16_nowait(my_tmem - hf * S::HH + 64 + hf * S::DH + 16, zb);
    tc::tmem_ld_wait();
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      uint32_t q[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 b = *reinterpret_cast<const float2*>(b2s + cb * 16 + i);
        const float u0 = upd(cb * 16 + i, __uint_as_float(cb ? zb[i] : za[i]) + b.x);
        const float u1 = upd(cb * 16 + i + 1, __uint_as_float(cb ? zb[i + 1] : za[i + 1]) + b.y);
        q[i / 2] = tc::pack_bf16x2(u0, u1);
      }
      *reinterpret_cast<uint4*>(A1 + (size_t)((hf * 4 + cb * 2) * S::TILE + row) * 16) = make_uint4(q[0], q[1], q[2], q[3]);
      *reinterpret_cast<uint4*>(A1 + (size_t)((hf * 4 + cb * 2 + 1) * S::TILE + row) * 16) = make_uint4(q[4], q[5], q[6], q[7]);
    }
    TCW_T(6);
    TCW_END();
  };

  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int ntiles = (p.B + S::TILE - 1) / S::TILE;
  const int n_iter = (ntiles + 2 * gridDim.x - 1) / (2 * gridDim.x);  // both groups run the same number of token rounds
  if (grp == 1) tc::named_bar_arrive(3, 512);                       // group 0 takes the first turn
  for (int it = 0; it < n_iter; ++it) {
    const int tile = blockIdx.x + grp * gridDim.x + it * 2 * gridDim.x;
    const bool last_iter = it + 1 == n_iter;
    if (tile >= ntiles) {  // no tile for this group in this round: keep the token moving
      for (int s = 0; s < 4 * (p.T - 1); ++s) {
        tc::named_bar_sync(tok_mine, 512);
        if (!(last_iter && s + 1 == 4 * (p.T - 1) && grp == 1)) tc::named_bar_arrive(tok_other, 512);
      }
      continue;
    }
    const int b = tile * S::TILE + row;
    const bool valid = b < p.B;
    float y[S::DH], w[S::DH];
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
#pragma unroll
    for (int c = 0; c < S::DH / 8; ++c)
      *reinterpret_cast<uint4*>(A1 + (size_t)((hf * (S::DH / 8) + c) * S::TILE + row) * 16) =
          make_uint4(tc::pack_bf16x2(y[8 * c], y[8 * c + 1]), tc::pack_bf16x2(y[8 * c + 2], y[8 * c + 3]),
                     tc::pack_bf16x2(y[8 * c + 4], y[8 * c + 5]), tc::pack_bf16x2(y[8 * c + 6], y[8 * c + 7]));
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      const float dt3 = dt * kWT, dt8 = dt * 0.125f, dt38 = dt * 0.375f;
      feval(false, [&](int i, float k) { w[i] = k; return y[i] + dt3 * k; });                     // k1 (live: y, w = k1)
      feval(false, [&](int i, float k) {                                                            // k2 (live: P, w)
        const float u = y[i] + dt * (k - w[i] * kWT);
        const float P = y[i] + dt8 * (w[i] + 3.f * k);
        w[i] = y[i] + dt * (w[i] - k);
        y[i] = P;
        return u;
      });
      feval(false, [&](int i, float k) { y[i] += dt38 * k; return w[i] + dt * k; });               // k3 (live: P')
      feval(last_iter && s + 2 == p.T && grp == 1, [&](int i, float k) { y[i] += dt8 * k; return y[i]; });  // k4
      if (valid) {
        float* o = p.traj + tcw_off(p.layout, s + 1, b, p.B, p.T, D) + hf * S::DH;
#pragma unroll
        for (int i = 0; i < S::DH; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      }
    }
  }
#ifdef GODE_TCW_TIMING
  if (p.dbg && blockIdx.x == 0 && (gtid == 0 || gtid == 255))
    for (int i = 0; i < 8; ++i) p.dbg[(grp * 2 + (gtid ? 1 : 0)) * 8 + i] = tph[i];
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(*s_tmem, 512);
}

int tc_rk4_fwd_wide(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                    int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st) {
  if (!(D == 64 && H == 256)) return GODE_ERR_SHAPE;
  TcWideArgs a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.B = B; a.T = T; a.layout = out_layout;
  if (dt_on_device) {
    a.dt_dev = dt;
  } else {
    if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
    for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  }
  using S2 = TcWide2Shape<64, 256>;
  const int ntiles = (B + S2::TILE - 1) / S2::TILE;
  int grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  // share of the tanh evaluations moved from the MUFU pipe to the FMA pipe (polynomial): developer switch GODE_TANH_POLY_N =
  // 0..4 pairs of every eight; default: see profiles/README.md round 2
  static int poly_n = -1;
  if (poly_n < 0) {
    const char* pe = getenv("GODE_TANH_POLY_N");
    poly_n = pe ? atoi(pe) : kTanhPolyDefault;
  }
  auto kern = poly_n == 1 ? tc_rk4_fwd_wide2_kernel<64, 256, 1>
              : poly_n == 2 ? tc_rk4_fwd_wide2_kernel<64, 256, 2>
              : poly_n == 3 ? tc_rk4_fwd_wide2_kernel<64, 256, 3>
              : poly_n == 4 ? tc_rk4_fwd_wide2_kernel<64, 256, 4>
                            : tc_rk4_fwd_wide2_kernel<64, 256, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
#ifdef GODE_TCW_TIMING
  const char* dbg = getenv("GODE_TCW_DBG");
  if (dbg && dbg[0] == '1') {
    cudaMalloc(&a.dbg, 256 * sizeof(long long));
    cudaMemset(a.dbg, 0, 256 * sizeof(long long));
    kern<<<grid, 512, S2::BYTES, st>>>(a);
    cudaStreamSynchronize(st);
    long long h[256];
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.dbg);
    const double nf = 4.0 * (T - 1) * ((ntiles + 2 * grid - 1) / (2 * grid));
    static const char* nm[8] = {"pack+sync+issue1", "mma1 wait+ld0", "token wait", "tanh epilogue", "mid/end sync+issue2", "mma2 wait", "layer-2 epilogue+RK", "-"};
    for (int w = 0; w < 4; ++w) {
      fprintf(stderr, "[tcw dbg] grp %d thread %3d cycles/feval:", w / 2, (w & 1) ? 255 : 0);
      for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.0f", nm[i], h[w * 8 + i] / nf);
      fprintf(stderr, "\n");
    }
    if (dbg[1] == 't') {  // GODE_TCW_DBG=1t: raw event times of fevals 16..23 (relative to the first), both groups
      long long t0 = h[64];
      for (int g = 0; g < 2; ++g)
        for (int f = 0; f < 8; ++f) {
          fprintf(stderr, "[tcw tl] g%d f%d:", g, f);
          for (int e = 0; e < 8; ++e) fprintf(stderr, " %6lld", h[64 + (g * 8 + f) * 8 + e] ? h[64 + (g * 8 + f) * 8 + e] - t0 : -1);
          fprintf(stderr, "\n");
        }
    }
    return launch_status();
  }
#endif
  kern<<<grid, 512, S2::BYTES, st>>>(a);
  return launch_status();
}

}  // namespace gode
