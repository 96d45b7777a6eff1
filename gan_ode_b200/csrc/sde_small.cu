// sde_small.cu — neural SDE sampler: fixed-step Euler–Maruyama with counter-based Philox Brownian increments, FP32.
// Reference: models/mocogan_sde.py:6-27 (SDEFunc: two ODEFunc-shaped MLPs, diagonal Ito noise) and :57-59
//   sdeint_adjoint(sde, x, linspace(0,1,T), method='euler', adjoint_method='euler', dt=2.5e-2)
// torchsde semantics restated in oracle/torchsde_restatement.py: the step grid comes from fp32 time
// accumulation (41 steps for the reference call) and the frames are LINEAR interpolations between the two straddling
// steps.  The host builds that grid exactly as torchsde does and passes it by value; the kernel does every step of every
// trajectory in one launch:  y <- y + f(y) h + g(y) (.) dW,   dW = sqrt(h) N(0,1).
//
// Brownian increments: either a caller-supplied table dW (n_steps,B,D) (parity with a given path), or generated in the
// kernel by Philox4x32-10 keyed on `seed` with counter (global trajectory index, step index, d_block, stream=0)
// -> 4 uint32 -> 2x Box–Muller -> 4 normals for state components 4*d_block .. +3 (oracle/philox.py::normals).  The
// backward kernel regenerates them from the same counters, so no noise is ever stored and results do not depend on how
// the batch is sharded over GPUs.
#include "small_field.cuh"
#include "launch.h"

namespace gode {

constexpr int kSdeMaxSteps = 320;
constexpr int kSdeMaxT = 64;

struct SdeArgs {
  const float *y0, *fW1, *fb1, *fW2, *fb2, *gW1, *gb1, *gW2, *gb2;
  const float* dW;          // (n_steps,B,D) or nullptr -> Philox
  const float* grad_out;    // bwd: (T,B,D)/(B,T,D)
  const float* states_in;   // bwd: (n_steps,B,D) state at the START of every step
  float* out;               // fwd: frames
  float* states;            // fwd: (n_steps,B,D) or nullptr
  float* grad_y0;
  float* grad_params;       // flat [f: W1|b1|W2|b2 | g: W1|b1|W2|b2]
  ReduceWs ws;
  unsigned long long seed;
  long long traj_offset;    // global index of local trajectory 0 (data parallel shards)
  int B, T, layout, n_steps;
  float h[kSdeMaxSteps];    // step sizes (fp32, as torchsde accumulates them)
  short out_step[kSdeMaxT]; // frame j is emitted after this step (frame 0 = y0)
  float w0[kSdeMaxT], w1[kSdeMaxT];  // frame_j = w0 * y_k + w1 * y_{k+1}
};

__device__ __forceinline__ size_t sde_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned int x) { return (float)x * 2.3283064e-10f + 1.1641532e-10f; }

// the lane's DL standard normals for (trajectory, step): components l*DL .. l*DL+DL-1
template <int DL>
__device__ __forceinline__ void philox_normals(unsigned long long seed, unsigned long long traj, int step, int l, float (&z)[DL]) {
  static_assert(DL == 1 || DL == 2 || DL == 4, "lane slice must tile a 4-wide Philox block");
  const int d0 = l * DL;
  const uint4 r = philox4x32_10(make_uint4((unsigned int)traj, (unsigned int)step, (unsigned int)(d0 >> 2), 0u),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  // Box–Muller on the pair(s) this lane needs only (a lane owning 1 or 2 components needs one of the block's two pairs),
  // with the hardware transcendental units: log2 (MUFU.LG2), rsqrt/sqrt, sin/cos of 2*pi*u via sinpi/cospi-style exact
  // range (u in (0,1) -> the argument of MUFU.SIN/COS stays in (0, 2*pi)).  Differences from the CPU contract
  // (oracle/philox.py, numpy float32) are ~1e-6 relative, far inside the 2e-5 the stream test allows.
  auto pair = [](unsigned int a, unsigned int b, float& n0, float& n1) {
    const float rad = sqrtf(-1.3862944f * __log2f(u01(a)));  // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
    float sn, cs;
    __sincosf(6.2831855f * u01(b), &sn, &cs);
    n0 = rad * sn;
    n1 = rad * cs;
  };
  if constexpr (DL == 4) {
    pair(r.x, r.y, z[0], z[1]);
    pair(r.z, r.w, z[2], z[3]);
  } else {
    const bool second = (d0 & 2) != 0;
    float n0, n1;
    pair(second ? r.z : r.x, second ? r.w : r.y, n0, n1);
    if constexpr (DL == 2) { z[0] = n0; z[1] = n1; }
    else z[0] = (d0 & 1) ? n1 : n0;
  }
}

template <int D, int H, int L>
__device__ __forceinline__ void brownian(const SdeArgs& p, int b, bool valid, int step, int l, float sqrt_h,
                                         float (&dw)[Shape<D, H, L>::DL]) {
  using S = Shape<D, H, L>;
  if (p.dW) {
#pragma unroll
    for (int i = 0; i < S::DL; ++i) dw[i] = 0.f;
    if (valid) load_frag<S::DL>(p.dW + ((size_t)step * p.B + b) * D + l * S::DL, dw);
  } else {
    philox_normals<S::DL>(p.seed, (unsigned long long)(p.traj_offset + b), step, l, dw);
#pragma unroll
    for (int i = 0; i < S::DL; ++i) dw[i] *= sqrt_h;
  }
}

// ---- forward -----------------------------------------------------------------------------------------------------------
template <int D, int H, int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sde_em_fwd_kernel(const __grid_constant__ SdeArgs p) {
  using S = Shape<D, H, L>;
  __shared__ __align__(16) float s_lines[WARPS * FwdLines<D, H, L>::kFloatsPerWarp];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / L, l = lane % L;
  FwdLines<D, H, L> ln;
  ln.bind(s_lines + warp * FwdLines<D, H, L>::kFloatsPerWarp, g);
  RowWeights<D, H, L> wf, wg;
  wf.load(p.fW1, p.fb1, p.fW2, p.fb2, l);
  wg.load(p.gW1, p.gb1, p.gW2, p.gb2, l);
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    float y[S::DL];
#pragma unroll
    for (int i = 0; i < S::DL; ++i) y[i] = 0.f;
    if (valid) {
      load_frag<S::DL>(p.y0 + (size_t)b * D + l * S::DL, y);
      store_frag<S::DL>(p.out + sde_off(p.layout, 0, b, p.B, p.T, D) + l * S::DL, y);
    }
    int jout = 1;
    for (int k = 0; k < p.n_steps; ++k) {
      const float h = p.h[k];
      if (p.states && valid) store_frag<S::DL>(p.states + ((size_t)k * p.B + b) * D + l * S::DL, y);
      float f[S::DL], gg[S::DL], dw[S::DL], y1[S::DL], hk[S::HL];
      brownian<D, H, L>(p, b, valid, k, l, sqrtf(h), dw);
      mlp_forward<D, H, L>(wf, ln, l, y, f, hk);
      mlp_forward<D, H, L>(wg, ln, l, y, gg, hk);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) y1[i] = y[i] + f[i] * h + gg[i] * dw[i];
      while (jout < p.T && p.out_step[jout] == k) {
        const float a0 = p.w0[jout], a1 = p.w1[jout];
        float o[S::DL];
#pragma unroll
        for (int i = 0; i < S::DL; ++i) o[i] = a0 * y[i] + a1 * y1[i];
        if (valid) store_frag<S::DL>(p.out + sde_off(p.layout, jout, b, p.B, p.T, D) + l * S::DL, o);
        ++jout;
      }
#pragma unroll
      for (int i = 0; i < S::DL; ++i) y[i] = y1[i];
    }
  }
}

// ---- backward: exact reverse-mode through the Euler–Maruyama steps, given the same increments --------------------------------
// Only the hidden activations of the two MLPs are recomputed (the field values themselves are not needed).
template <int D, int H, int L, class Lines>
__device__ __forceinline__ void mlp_hidden(const RowWeights<D, H, L>& w, const Lines& ln, int l,
                                           const float (&u)[Shape<D, H, L>::DL], float (&hk)[Shape<D, H, L>::HL],
                                           bool restore_u) {
  using S = Shape<D, H, L>;
  if (restore_u) {
    __syncwarp();
    store_frag<S::DL>(ln.y + l * S::DL, u);
    __syncwarp();
  }
#pragma unroll
  for (int jl = 0; jl < S::HL; ++jl) hk[jl] = tanhf(dot_line<D>(w.w1[jl], ln.y, w.b1[jl]));
  __syncwarp();  // everyone has read ln.h of the previous use before it is overwritten
  store_frag<S::HL>(ln.h + l * S::HL, hk);
}

template <int D, int H, int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sde_em_bwd_kernel(const __grid_constant__ SdeArgs p) {
  using S = Shape<D, H, L>;
  using BL = BwdLines<D, H, L>;
  using CW = ColWeights<D, H, L>;
  extern __shared__ __align__(16) float smem[];
  float* s_lines = smem;
  float* s_cwf = s_lines + WARPS * BL::kFloatsPerWarp;
  float* s_cwg = s_cwf + CW::kFloats;
  float* s_red = s_cwg + CW::kFloats;  // WARPS * P
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  SyncState ss;
  ss.begin(p.ws.gs);
  BL ln;
  ln.bind(s_lines + warp * BL::kFloatsPerWarp, g);
  CW cwf, cwg;
  cwf.bind(s_cwf);
  cwg.bind(s_cwg);
  cwf.stage(p.fW1, p.fW2, tid, WARPS * 32);
  cwg.stage(p.gW1, p.gW2, tid, WARPS * 32);
  RowWeights<D, H, L> wf, wg;
  wf.load(p.fW1, p.fb1, p.fW2, p.fb2, l);
  wg.load(p.gW1, p.gb1, p.gW2, p.gb2, l);
  GradAcc<D, H, L> af, ag;
  af.zero();
  ag.zero();
  __syncthreads();
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    const float sc = valid ? 1.f : 0.f;
    float yb[S::DL];  // cotangent of y_{k+1}
#pragma unroll
    for (int i = 0; i < S::DL; ++i) yb[i] = 0.f;
    int jout = p.T - 1;
    for (int k = p.n_steps - 1; k >= 0; --k) {
      const float h = p.h[k];
      float yk[S::DL], ykb[S::DL], dw[S::DL], hf[S::HL], hg[S::HL];
#pragma unroll
      for (int i = 0; i < S::DL; ++i) { yk[i] = 0.f; ykb[i] = 0.f; }
      if (valid) load_frag<S::DL>(p.states_in + ((size_t)k * p.B + b) * D + l * S::DL, yk);
      // frames emitted after this step: frame = w0 y_k + w1 y_{k+1}
      while (jout >= 1 && p.out_step[jout] == k) {
        float go[S::DL];
#pragma unroll
        for (int i = 0; i < S::DL; ++i) go[i] = 0.f;
        if (valid) load_frag<S::DL>(p.grad_out + sde_off(p.layout, jout, b, p.B, p.T, D) + l * S::DL, go);
        const float a0 = p.w0[jout], a1 = p.w1[jout];
#pragma unroll
        for (int i = 0; i < S::DL; ++i) { ykb[i] = fmaf(a0, go[i], ykb[i]); yb[i] = fmaf(a1, go[i], yb[i]); }
        --jout;
      }
      brownian<D, H, L>(p, b, valid, k, l, sqrtf(h), dw);
      // y_{k+1} = y_k + f(y_k) h + g(y_k) (.) dW
      float cf[S::DL], cg[S::DL], vb[S::DL];
#pragma unroll
      for (int i = 0; i < S::DL; ++i) { cf[i] = h * yb[i]; cg[i] = dw[i] * yb[i]; ykb[i] += yb[i]; }
      mlp_hidden<D, H, L>(wf, ln, l, yk, hf, true);
      mlp_vjp<D, H, L>(cwf, ln, l, hf, cf, sc, vb, af);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) ykb[i] += vb[i];
      mlp_hidden<D, H, L>(wg, ln, l, yk, hg, false);  // ln.y still holds y_k
      mlp_vjp<D, H, L>(cwg, ln, l, hg, cg, sc, vb, ag);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) yb[i] = ykb[i] + vb[i];
    }
    if (valid) {
      float g0[S::DL];
      load_frag<S::DL>(p.grad_out + sde_off(p.layout, 0, b, p.B, p.T, D) + l * S::DL, g0);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) g0[i] += yb[i];
      store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, g0);
    }
  }
  // two parameter sets: reduce one after the other (same persistent counter, running target; separate partial rows)
  ReduceWs w1 = p.ws, w2 = p.ws;
  w2.partials = p.ws.partials + (size_t)gridDim.x * S::P;
  reduce_param_grads<D, H, L, WARPS>(af, s_red, w1, ss, p.grad_params, lane, warp, tid);
  __syncthreads();
  reduce_param_grads<D, H, L, WARPS>(ag, s_red, w2, ss, p.grad_params + S::P, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
}

// ---- host ----------------------------------------------------------------------------------------------------------------
size_t sde_small_workspace_bytes(int D, int H) {
  const int P = H * D + H + D * H + D;
  return (size_t)GODE_SYNC_REGION_BYTES + sizeof(float) * 2 * (size_t)P * (size_t)bwd_grid_cap();
}

static int fill_grid(SdeArgs& a, const float* h_host, int n_steps, const int* out_step_host, const float* w0_host,
                     const float* w1_host, int T) {
  if (n_steps > kSdeMaxSteps || T > kSdeMaxT) return GODE_ERR_T_TOO_LONG;
  a.n_steps = n_steps;
  for (int i = 0; i < n_steps; ++i) a.h[i] = h_host[i];
  for (int j = 0; j < T; ++j) { a.out_step[j] = (short)out_step_host[j]; a.w0[j] = w0_host[j]; a.w1[j] = w1_host[j]; }
  return GODE_OK;
}

int sde_small_fwd(const float* y0, const float* const* fw, const float* const* gw, const float* h_host, int n_steps,
                  const int* out_step_host, const float* w0_host, const float* w1_host, int B, int D, int H, int T,
                  const float* dW, unsigned long long seed, long long traj_offset, int layout, float* out, float* states,
                  cudaStream_t st) {
  SdeArgs a{};
  a.y0 = y0; a.fW1 = fw[0]; a.fb1 = fw[1]; a.fW2 = fw[2]; a.fb2 = fw[3];
  a.gW1 = gw[0]; a.gb1 = gw[1]; a.gW2 = gw[2]; a.gb2 = gw[3];
  a.dW = dW; a.seed = seed; a.traj_offset = traj_offset; a.B = B; a.T = T; a.layout = layout; a.out = out; a.states = states;
  if (int rc = fill_grid(a, h_host, n_steps, out_step_host, w0_host, w1_host, T)) return rc;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  constexpr int WARPS = 4, L = 8;
  using S = Shape<16, 16, L>;
  const int per_cta = WARPS * S::G;
  int grid = (B + per_cta - 1) / per_cta;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  sde_em_fwd_kernel<16, 16, L, WARPS><<<grid, WARPS * 32, 0, st>>>(a);
  return launch_status();
}

int sde_small_bwd(const float* states, const float* grad_out, const float* const* fw, const float* const* gw,
                  const float* h_host, int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host,
                  int B, int D, int H, int T, const float* dW, unsigned long long seed, long long traj_offset, int layout,
                  float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  SdeArgs a{};
  a.states_in = states; a.grad_out = grad_out;
  a.fW1 = fw[0]; a.fb1 = fw[1]; a.fW2 = fw[2]; a.fb2 = fw[3];
  a.gW1 = gw[0]; a.gb1 = gw[1]; a.gW2 = gw[2]; a.gb2 = gw[3];
  a.dW = dW; a.seed = seed; a.traj_offset = traj_offset; a.B = B; a.T = T; a.layout = layout;
  a.grad_y0 = grad_y0; a.grad_params = grad_params;
  if (int rc = fill_grid(a, h_host, n_steps, out_step_host, w0_host, w1_host, T)) return rc;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  constexpr int WARPS = 4, L = 16;
  using S = Shape<16, 16, L>;
  auto kern = sde_em_bwd_kernel<16, 16, L, WARPS>;
  const size_t smem = sizeof(float) * (WARPS * BwdLines<16, 16, L>::kFloatsPerWarp + 2 * ColWeights<16, 16, L>::kFloats + WARPS * S::P);
  static int limit_cache = 0;
  int cap = coop_limit(kern, WARPS * 32, smem, limit_cache);
  if (cap <= 0) return GODE_ERR_COOP;
  if (cap > bwd_grid_cap()) cap = bwd_grid_cap();
  const int per_cta = WARPS * S::G;
  int grid = (B + per_cta - 1) / per_cta;
  if (grid > cap) grid = cap;
  if (ws_bytes < sde_small_workspace_bytes(D, H)) return GODE_ERR_WORKSPACE;
  grid_sync_bind(a.ws.gs, workspace);
  a.ws.partials = reinterpret_cast<float*>(ws_scratch(workspace));
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(WARPS * 32), args, smem, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

}  // namespace gode
