// tc_rk4_adj_small.cu — tensor-core continuous adjoint of the fixed-grid RK4 (3/8 rule) solve for the REFERENCE shape
// D = H = 16 (BF16 operands, FP32 accumulate in TMEM).  Same algorithm and operand tricks as tc_rk4_adj_wide.cu
// (torchdiffeq adjoint.py semantics: per output interval one 3/8 step of (y, a, theta_bar) in reversed time from the stored
// y_i; c = stage quadrature weight carried on a, so g', d', v' come out scaled), re-cut for tiny contractions:
//   * one thread per trajectory row, 128 rows per CTA, 3 CTAs per SM (TMEM 128 columns, 60 KB shared memory each);
//   * every MMA is M = 128, N = 16 (one or two K steps), except the weight gradients, which are ONE stacked contraction per
//     stage:  [h | d']^T (M = 32 useful of 128, MN-major view of the HD tile)  x  [c a | u | 1] (N = 48, MN-major view of the
//     stage-input tile)  ->  rows 0..15 x cols 0..15 = dW2^T, rows 16..31 x cols 16..31 = dW1, col 32 = db1 (the other two
//     blocks are by-products).  Its accumulator stays in TMEM for the whole kernel and nothing waits for it;
//   * db2 = sum c a in registers; per-CTA partial rows -> fixed-order reduction (deterministic).
// The FP32 adjoint of this shape is FMA-pipe bound (3 x 512 FMAs per stage per trajectory); here the FMAs are on the tensor
// pipe and a stage is ~16 tanh + ~250 epilogue instructions per trajectory.
#include "launch.h"
#include "tc_common.cuh"

namespace gode {

namespace {

constexpr float kThird = 0.33333334f;
constexpr int D = 16, H = 16, TILE = 128;
constexpr int CH = TILE * 16;                  // bytes of one 8-column chunk of a 128-row tile
// stage-input tile R (double buffered): chunks 0,1 = c*a ; 2,3 = u ; 4 = (1,1,0,...) ; 5 = 0   -> N = 48 for the dW MMAs
constexpr int R_CHUNKS = 6, R_BYTES = R_CHUNKS * CH;
// HD tile: chunks 0,1 = h ; 2,3 = d' ; 4..15 = 0 (M = 128 view for the dW MMAs)
constexpr int HD_CHUNKS = 16, HD_BYTES = HD_CHUNKS * CH;
constexpr int OFF_B2F = 0;                     // fp32 b2 (16) | staging of raw weights (fp32, 544 floats)
constexpr int OFF_RAW = 64;
constexpr int OFF_B1 = 2304;                   // [W1|b1] K-major: 4 chunks x 16 rows x 16 B
constexpr int OFF_B2 = OFF_B1 + 4 * 256;       // W2 K-major: 2 chunks x 16 rows x 16 B
constexpr int OFF_R = OFF_B2 + 2 * 256;        // 3840
constexpr int OFF_HD = OFF_R + 2 * R_BYTES;
constexpr int OFF_RED = OFF_HD + HD_BYTES;     // 4 x 16 floats
constexpr int OFF_BAR = OFF_RED + 256;
constexpr int SMEM_BYTES = OFF_BAR + 64;
static_assert(OFF_R % 128 == 0 && OFF_HD % 128 == 0, "operand tiles are 128-byte aligned");
constexpr uint32_t T_Z = 0, T_F = 16, T_G = 32, T_V = 48, T_W = 64, T_COLS = 128;
constexpr int P = H * D + H + D * H + D;       // 544
constexpr uint32_t A_MN = 1u << 15, B_MN = 1u << 16;

struct AdjArgs {
  const float *traj, *grad_traj, *W1, *b1, *W2, *b2;
  float* grad_y0;
  float* partial;  // [grid][P]
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

__global__ void __launch_bounds__(128, 3) tc_rk4_adj_small_kernel(const __grid_constant__ AdjArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* b2f = reinterpret_cast<float*>(smem + OFF_B2F);
  float* raw = reinterpret_cast<float*>(smem + OFF_RAW);
  unsigned char* B1 = smem + OFF_B1;
  unsigned char* B2 = smem + OFF_B2;
  unsigned char* HD = smem + OFF_HD;
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  uint64_t* mbar_m = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_m + 1);
  const int tid = threadIdx.x, lane = tid & 31, row = tid;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 0) tc::tmem_alloc(s_tmem, T_COLS);
  if (tid == 0) {
    tc::mbar_init(mbar_m, 1);
    tc::mbar_fence_init();
  }
  // ---- weights: plain loads (2.2 KB) -> BF16 operand tiles; constant chunks of R and HD ----
  for (int e = tid; e < H * D; e += 128) { raw[e] = p.W1[e]; raw[H * D + H + e] = p.W2[e]; }
  if (tid < H) raw[H * D + tid] = p.b1[tid];
  if (tid < D) b2f[tid] = p.b2[tid];
  __syncthreads();
  if (tid < 2 * 16) {  // B1 data chunks: [kc][j][8 d] <- W1[j][8 kc ..]
    const int kc = tid >> 4, j = tid & 15;
    const float* v = raw + j * D + kc * 8;
    *reinterpret_cast<uint4*>(B1 + (kc * 16 + j) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  } else if (tid < 3 * 16) {  // K = 16, 17: b1 as two bf16 terms against the (1,1) columns of the input tile
    const int j = tid & 15;
    const float b = raw[H * D + j];
    const float hi = __bfloat162float(__float2bfloat16_rn(b));
    *reinterpret_cast<uint4*>(B1 + (2 * 16 + j) * 16) = make_uint4(tc::pack_bf16x2(hi, b - hi), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(B1 + (3 * 16 + j) * 16) = make_uint4(0u, 0u, 0u, 0u);
  } else if (tid < 5 * 16) {  // B2: [kc][d][8 j] <- W2[d][8 kc ..]
    const int kc = (tid >> 4) - 3, d = tid & 15;
    const float* v = raw + H * D + H + d * H + kc * 8;
    *reinterpret_cast<uint4*>(B2 + (kc * 16 + d) * 16) =
        make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    unsigned char* r = smem + OFF_R + g * R_BYTES;
    *reinterpret_cast<uint4*>(r + 4 * CH + row * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(r + 5 * CH + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int c = 4; c < HD_CHUNKS; ++c) *reinterpret_cast<uint4*>(HD + c * CH + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const uint32_t my_t = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t sB1 = tc::smem_u32(B1), sB2 = tc::smem_u32(B2), sHD = tc::smem_u32(HD), sR = tc::smem_u32(smem + OFF_R);
  constexpr uint32_t id_kk = tc::make_idesc(tc::kFmtBF16, 128, 16);            // both operands K-major
  constexpr uint32_t id_kn = tc::make_idesc(tc::kFmtBF16, 128, 16) | B_MN;     // B read N-major
  constexpr uint32_t id_w = tc::make_idesc(tc::kFmtBF16, 128, 48) | A_MN | B_MN;
  uint32_t phase = 0, acc_live = 0;

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
  auto pack16 = [&](unsigned char* dst_chunk0, const float(&v)[16], float scale) {  // two chunks of this thread's row
    *reinterpret_cast<uint4*>(dst_chunk0 + row * 16) =
        make_uint4(tc::pack_bf16x2(scale * v[0], scale * v[1]), tc::pack_bf16x2(scale * v[2], scale * v[3]),
                   tc::pack_bf16x2(scale * v[4], scale * v[5]), tc::pack_bf16x2(scale * v[6], scale * v[7]));
    *reinterpret_cast<uint4*>(dst_chunk0 + CH + row * 16) =
        make_uint4(tc::pack_bf16x2(scale * v[8], scale * v[9]), tc::pack_bf16x2(scale * v[10], scale * v[11]),
                   tc::pack_bf16x2(scale * v[12], scale * v[13]), tc::pack_bf16x2(scale * v[14], scale * v[15]));
  };

  float dbb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dbb[i] = 0.f;

  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int ntiles = (p.B + TILE - 1) / TILE;
  int cur = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * TILE + row;
    const bool valid = b < p.B;
    float Y[16], KY[16], A[16], WA[16];
    auto load16 = [&](const float* src, float(&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 t4 = valid ? *reinterpret_cast<const float4*>(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
      }
    };
    load16(p.grad_traj + (valid ? off3(p.layout, p.T - 1, b, p.B, p.T) : 0), A);
    float Yn[16], Gn[16];  // stored state of the next interval / upstream gradient of this one, fetched an interval ahead
    load16(p.traj + (valid ? off3(p.layout, p.T - 1, b, p.B, p.T) : 0), Yn);
    load16(p.grad_traj + (valid ? off3(p.layout, p.T - 2, b, p.B, p.T) : 0), Gn);
    for (int i = p.T - 1; i >= 1; --i) {
      const float dt = dtp[i - 1];
      const float cs[4] = {dt * 0.125f, dt * 0.375f, dt * 0.375f, dt * 0.125f};
      float gp[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) { Y[e] = Yn[e]; gp[e] = Gn[e]; }
      if (i > 1) {
        load16(p.traj + (valid ? off3(p.layout, i - 1, b, p.B, p.T) : 0), Yn);
        load16(p.grad_traj + (valid ? off3(p.layout, i - 2, b, p.B, p.T) : 0), Gn);
      }
      {
        unsigned char* r = smem + OFF_R + cur * R_BYTES;
        pack16(r, A, cs[0]);
        pack16(r + 2 * CH, Y, 1.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) dbb[j] += cs[0] * A[j];
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float c = cs[s], rc = 1.f / c;
        const uint32_t sRc = sR + cur * R_BYTES;
        unsigned char* rn = smem + OFF_R + (cur ^ 1) * R_BYTES;
        // ---- z1 = [u|1][W1|b1]^T ----
        sync_issue([&] {
          const uint64_t dA = tc::make_smem_desc(sRc + 2 * CH, CH, 128);
          const uint64_t dB = tc::make_smem_desc(sB1, 256, 128);
          tc::mma_ss<false>(tmem + T_Z, dA, dB, id_kk, 0);
          tc::mma_ss<false>(tmem + T_Z, adv(dA, 2 * CH), adv(dB, 2 * 256), id_kk, 1);
          tc::mma_commit(mbar_m);
        });
        uint32_t hq[8];
        {
          uint32_t z[16];
          tc::tmem_ld16_nowait(my_t + T_Z, z);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; e += 2)
            hq[e / 2] = tc::pack_bf16x2(tc::tanh_approx(__uint_as_float(z[e])), tc::tanh_approx(__uint_as_float(z[e + 1])));
          *reinterpret_cast<uint4*>(HD + row * 16) = make_uint4(hq[0], hq[1], hq[2], hq[3]);
          *reinterpret_cast<uint4*>(HD + CH + row * 16) = make_uint4(hq[4], hq[5], hq[6], hq[7]);
        }
        // ---- f = h W2^T (not needed after the last stage) and g' = (c a) W2 ----
        sync_issue([&] {
          if (s < 3) tc::mma_ss<false>(tmem + T_F, tc::make_smem_desc(sHD, CH, 128), tc::make_smem_desc(sB2, 256, 128), id_kk, 0);
          tc::mma_ss<false>(tmem + T_G, tc::make_smem_desc(sRc, CH, 128), tc::make_smem_desc(sB2, 128, 256), id_kn, 0);
          tc::mma_commit(mbar_m);
        });
        {
          uint32_t zf[16], zg[16];
          if (s < 3) tc::tmem_ld16_nowait(my_t + T_F, zf);
          tc::tmem_ld16_nowait(my_t + T_G, zg);
          tc::tmem_ld_wait();
          if (s < 3) {  // y part of the 3/8 step in reversed time: ky = -f; next stage's u -> the other input tile
            float un[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float ky = -(__uint_as_float(zf[e]) + b2f[e]);
              if (s == 0) { KY[e] = ky; un[e] = Y[e] + dt * kThird * ky; }
              if (s == 1) { un[e] = Y[e] + dt * (ky - KY[e] * kThird); KY[e] = Y[e] + dt * (KY[e] - ky); }
              if (s == 2) { un[e] = KY[e] + dt * ky; }
            }
            pack16(rn + 2 * CH, un, 1.f);
          }
          uint32_t o[8];  // d' = g' (1 - h^2) -> HD chunks 2,3
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float ha = bf_lo(hq[e]), hb = bf_hi(hq[e]);
            o[e] = tc::pack_bf16x2(__uint_as_float(zg[2 * e]) * (1.f - ha * ha), __uint_as_float(zg[2 * e + 1]) * (1.f - hb * hb));
          }
          *reinterpret_cast<uint4*>(HD + 2 * CH + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(HD + 3 * CH + row * 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        // ---- v' = d' W1; then the stacked weight-gradient contraction, which nothing waits for ----
        sync_issue([&] {
          tc::mma_ss<false>(tmem + T_V, tc::make_smem_desc(sHD + 2 * CH, CH, 128), tc::make_smem_desc(sB1, 128, 256), id_kn, 0);
          tc::mma_commit(mbar_m);
          const uint64_t dAh = tc::make_smem_desc(sHD, 128, CH);   // [h | d' | 0...]^T : M-major view, 16 chunks
          const uint64_t dBr = tc::make_smem_desc(sRc, 128, CH);   // [c a | u | 1 | 0]  : N-major view, 6 chunks
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k) tc::mma_ss<false>(tmem + T_W, adv(dAh, k * 256), adv(dBr, k * 256), id_w, acc_live | (uint32_t)(k > 0));
        });
        acc_live = 1;
        {  // a part: ka = v'/c; next stage's (c a) -> the other input tile
          uint32_t zv[16];
          tc::tmem_ld16_nowait(my_t + T_V, zv);
          tc::tmem_ld_wait();
          float an[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float ka = __uint_as_float(zv[e]) * rc;
            if (s == 0) { WA[e] = ka; an[e] = A[e] + dt * kThird * ka; }
            if (s == 1) {
              an[e] = A[e] + dt * (ka - WA[e] * kThird);
              const float Pn = A[e] + dt * 0.125f * (WA[e] + 3.f * ka);
              WA[e] = A[e] + dt * (WA[e] - ka);
              A[e] = Pn;
            }
            if (s == 2) { an[e] = WA[e] + dt * ka; A[e] += dt * 0.375f * ka; }
            if (s == 3) { A[e] += dt * 0.125f * ka + gp[e]; }
          }
          if (s < 3) {
            pack16(rn, an, cs[s + 1]);
#pragma unroll
            for (int e = 0; e < 16; ++e) dbb[e] += cs[s + 1] * an[e];
          }
        }
        cur ^= 1;
      }
    }
    if (valid) {
      float* o = p.grad_y0 + (size_t)b * D;
#pragma unroll
      for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(A[e], A[e + 1], A[e + 2], A[e + 3]);
    }
  }

  // ---- this CTA's partial row [W1|b1|W2|b2] ----
  sync_issue([&] { tc::mma_commit(mbar_m); });  // drains the last weight-gradient MMAs
  float* part = p.partial + (size_t)blockIdx.x * P;
  if (warp == 0) {
    uint32_t w0[16], w1[16], w2[16];
    tc::tmem_ld16_nowait(my_t + T_W, w0);
    tc::tmem_ld16_nowait(my_t + T_W + 16, w1);
    tc::tmem_ld16_nowait(my_t + T_W + 32, w2);
    tc::tmem_ld_wait();
    if (lane < 16) {  // row j of [h]^T [c a]: dW2^T[j][d] -> W2[d][j]
#pragma unroll
      for (int d = 0; d < 16; ++d) part[H * D + H + d * H + lane] = acc_live ? __uint_as_float(w0[d]) : 0.f;
    } else {          // row j of [d']^T [u|1]: dW1[j][d], db1[j]
      const int j = lane - 16;
#pragma unroll
      for (int d = 0; d < 16; ++d) part[j * D + d] = acc_live ? __uint_as_float(w1[d]) : 0.f;
      part[H * D + j] = acc_live ? __uint_as_float(w2[0]) : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float v = dbb[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    dbb[e] = v;
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) red[warp * 16 + e] = dbb[e];
  }
  __syncthreads();
  if (tid < D) part[H * D + H + D * H + tid] = (red[tid] + red[16 + tid]) + (red[32 + tid] + red[48 + tid]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, T_COLS);
}

__global__ void tc_adj_small_reduce_kernel(const float* __restrict__ partial, int rows, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P) return;
  float s = 0.f;
  for (int k = 0; k < rows; ++k) s += partial[(size_t)k * P + e];
  out[e] = s;
}

int adj_grid(int B) {
  const int ntiles = (B + TILE - 1) / TILE, cap = sm_count() * 3;
  return ntiles < cap ? ntiles : cap;
}

}  // namespace

size_t tc_rk4_adj_small_workspace_bytes(int B) { (void)B; return sizeof(float) * (size_t)P * (size_t)(sm_count() * 4) + 256; }

int tc_rk4_adj_small(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                     const float* b2, const float* dt, int dt_on_device, int B, int Dd, int Hh, int T, int layout,
                     float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (!(Dd == D && Hh == H)) return GODE_ERR_SHAPE;
  const int grid = adj_grid(B);
  if (ws_bytes < sizeof(float) * (size_t)P * (size_t)grid) return GODE_ERR_WORKSPACE;
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
  cudaError_t e = cudaFuncSetAttribute(tc_rk4_adj_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return -(1000 + (int)e);
  tc_rk4_adj_small_kernel<<<grid, 128, SMEM_BYTES, st>>>(a);
  tc_adj_small_reduce_kernel<<<(P + 255) / 256, 256, 0, st>>>(a.partial, grid, grad_params);
  return launch_status();
}

}  // namespace gode
