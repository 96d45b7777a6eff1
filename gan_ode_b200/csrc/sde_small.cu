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
#include "philox.cuh"
#include "launch.h"

namespace gode {

constexpr int kSdeMaxSteps = GODE_SDE_MAX_STEPS;
constexpr int kSdeMaxT = GODE_SDE_MAX_FRAMES;

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

// ---- Brownian path on a cell grid (torchsde's stochastic adjoint) ----------------------------------------------
// sdeint_adjoint's backward re-solves every output interval in reverse with its own step grid, so it needs increments over
// intervals that are not forward steps.  The path is therefore sampled on the UNION of the forward and reverse step times
// ("cells", built by the host exactly as oracle/torchsde_restatement.py::adjoint_time_grid): cell r carries an independent
// N(0, len_r) increment — Philox counter (global trajectory, cell, d_block, stream = 1) or row r of a caller's (R,B,D)
// table — and the increment of any step is the left-to-right fp32 sum of the cells it covers.  Forward and backward regenerate
// the same cells from the same counters: nothing is stored, and the result does not depend on how the batch is sharded.
constexpr int kSdeMaxCells = GODE_SDE_MAX_CELLS;
constexpr int kSdeMaxRev = GODE_SDE_MAX_REV_STEPS;

struct SdeCellArgs {
  int R;                             // number of cells
  short fwd_lo[kSdeMaxSteps + 1];    // forward step k covers cells [fwd_lo[k], fwd_lo[k+1])
  float rs[kSdeMaxCells];            // sqrt(cell length), fp32
};

template <int DL>
__device__ __forceinline__ void philox_cell_normals(unsigned long long seed, unsigned long long traj, int cell, int l, float (&z)[DL]) {
  static_assert(DL == 1 || DL == 2 || DL == 4, "lane slice must tile a 4-wide Philox block");
  const int d0 = l * DL;
  const uint4 r = philox4x32_10(make_uint4((unsigned int)traj, (unsigned int)cell, (unsigned int)(d0 >> 2), 1u),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  auto pair = [](unsigned int a, unsigned int b, float& n0, float& n1) {
    const float rad = sqrtf(-1.3862944f * __log2f(u01(a)));
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

// W(cell hi) - W(cell lo): sum of the cells [lo, hi), left to right
template <int D, int DL>
__device__ __forceinline__ void brownian_cells(const float* __restrict__ dW, unsigned long long seed, long long traj_offset,
                                               const float* __restrict__ rs, int B, int b, bool valid, int lo, int hi, int l,
                                               float (&dw)[DL]) {
#pragma unroll
  for (int i = 0; i < DL; ++i) dw[i] = 0.f;
  for (int r = lo; r < hi; ++r) {
    float z[DL];
    if (dW) {
#pragma unroll
      for (int i = 0; i < DL; ++i) z[i] = 0.f;
      if (valid) load_frag<DL>(dW + ((size_t)r * B + b) * D + l * DL, z);
    } else {
      philox_cell_normals<DL>(seed, (unsigned long long)(traj_offset + b), r, l, z);
      const float sr = rs[r];
#pragma unroll
      for (int i = 0; i < DL; ++i) z[i] *= sr;
    }
#pragma unroll
    for (int i = 0; i < DL; ++i) dw[i] = (r == lo) ? z[i] : dw[i] + z[i];
  }
}

// ---- forward -----------------------------------------------------------------------------------------------------------
// CELLS: the increments come from the cell grid (second kernel parameter) instead of one draw per step.
template <int D, int H, int L, int WARPS, bool CELLS>
__global__ void __launch_bounds__(WARPS * 32) sde_em_fwd_kernel(const __grid_constant__ SdeArgs p,
                                                                const __grid_constant__ SdeCellArgs c) {
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
      if constexpr (CELLS) brownian_cells<D, S::DL>(p.dW, p.seed, p.traj_offset, c.rs, p.B, b, valid, c.fwd_lo[k], c.fwd_lo[k + 1], l, dw);
      else brownian<D, H, L>(p, b, valid, k, l, sqrtf(h), dw);
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
  reduce_param_grads<D, H, L, WARPS, false>(af, s_red, w1, ss, p.grad_params, lane, warp, tid);
  __syncthreads();
  reduce_param_grads<D, H, L, WARPS>(ag, s_red, w2, ss, p.grad_params + S::P, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
}

// ---- backward of sdeint_adjoint: torchsde's stochastic adjoint (adjoint.py / adjoint_sde.py, Euler, (ito, diagonal)) -----------
// Restated in oracle/torchsde_restatement.py::sdeint_adjoint.  Per output interval i = T-1 .. 1 the augmented state
// (y, a, theta_bar) is integrated over [-t_i, -t_{i-1}] with the same dt (host-built step list: 3 steps per interval for
// the reference call, 45 in all) against the REVERSED Brownian path, then y <- frames[i-1], a <- a + grad_frames[i-1].  One
// Euler step of length h over the forward-time interval [t - h, t] with increment v = W(t) - W(t - h):
//     f, g       drift / diffusion MLPs at y;   s = 1 - tanh'^2 of the diffusion's hidden layer
//     gdg        = J_g^T g            (vjp(g, y, g): torchsde's "double Stratonovich correction", f_corr = f - gdg)
//     a_dg       = J_g^T a
//     y     <- y - (f - gdg) h - g v
//     a     <- a + [J_f^T a - d(a . gdg)/dy] h + J_g^T (a_dg h + a v)
//     theta <- theta + [d(a . f)/dtheta_f] h  (drift)      + [-d(a . gdg)/dtheta_g] h + d(g . (a_dg h + a v))/dtheta_g  (diffusion)
// with the second-order term d(a . gdg) in closed form for g = W2 tanh(W1 y + b1) + b2 (p = W1 a, q = W2^T g, r = p s):
//     gbar = W2 r,  hbar = W2^T gbar - 2 h p q,  zbar = s hbar,  d/dy = W1^T zbar,
//     d/dW1 = zbar y^T + (s q) a^T,  d/db1 = zbar,  d/dW2 = g r^T + gbar h^T,  d/db2 = gbar
// (checked against autograd of the restatement to 1e-16 in fp64).  Only the forward's OUTPUT frames are read: no per-step
// states are stored for this backward.
struct SdeAdjArgs {
  const float* frames;
  const float* grad_out;
  const float *fW1, *fb1, *fW2, *fb2, *gW1, *gb1, *gW2, *gb2;
  const float* dW;          // (R,B,D) cell increments or nullptr -> Philox
  float* grad_y0;
  float* grad_params;       // flat [f: W1|b1|W2|b2 | g: W1|b1|W2|b2]
  ReduceWs ws;
  unsigned long long seed;
  long long traj_offset;
  int B, T, layout, n_rev;
  short ibeg[kSdeMaxT], iend[kSdeMaxT];   // reverse steps of output interval i: [ibeg[i], iend[i])
  float h[kSdeMaxRev];                    // reverse step sizes (fp32, as torchsde accumulates -t)
  short lo[kSdeMaxRev], hi[kSdeMaxRev];   // cells covered by reverse step n, in forward time
};

template <int D, int H, int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sde_adjoint_bwd_kernel(const __grid_constant__ SdeAdjArgs p,
                                                                     const __grid_constant__ SdeCellArgs c) {
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
    float y[S::DL], a[S::DL];
#pragma unroll
    for (int i = 0; i < S::DL; ++i) { y[i] = 0.f; a[i] = 0.f; }
    if (valid) {
      load_frag<S::DL>(p.frames + sde_off(p.layout, p.T - 1, b, p.B, p.T, D) + l * S::DL, y);
      load_frag<S::DL>(p.grad_out + sde_off(p.layout, p.T - 1, b, p.B, p.T, D) + l * S::DL, a);
    }
    for (int i = p.T - 1; i >= 1; --i) {
      for (int n = p.ibeg[i]; n < p.iend[i]; ++n) {
        const float hs = p.h[n];
        float v[S::DL];
        brownian_cells<D, S::DL>(p.dW, p.seed, p.traj_offset, c.rs, p.B, b, valid, p.lo[n], p.hi[n], l, v);
        // A. drift MLP: f and J_f^T a (parameter part scaled by the step)
        float f[S::DL], hf[S::HL], vyf[S::DL];
        __syncwarp();
        mlp_forward<D, H, L>(wf, ln, l, y, f, hf);                       // ln.y = y, ln.h = hf
        mlp_vjp<D, H, L>(cwf, ln, l, hf, a, sc * hs, vyf, af);           // ln.a = a, ln.dl = delta_f
        // B. diffusion MLP: hidden layer, s, g
        float hg[S::HL], sg[S::HL], gg[S::DL];
        mlp_hidden<D, H, L>(wg, ln, l, y, hg, false);                    // ln.h = hg (ln.y still y)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < S::HL; ++j) sg[j] = 1.f - hg[j] * hg[j];
#pragma unroll
        for (int d = 0; d < S::DL; ++d) gg[d] = dot_line<H>(wg.w2[d], ln.h, wg.b2[d]);
        // C. with a in ln.a: t_a = W2^T a, p = W1 a; then g into ln.a: q = W2^T g
        float ta[S::HL], pp[S::HL], qq[S::HL];
#pragma unroll
        for (int j = 0; j < S::HL; ++j) {
          ta[j] = dot_smem<D>(cwg.w2t + (j * L + l) * S::YS, ln.a);
          pp[j] = dot_line<D>(wg.w1[j], ln.a, 0.f);
        }
        __syncwarp();
        store_frag<S::DL>(ln.a + l * S::DL, gg);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < S::HL; ++j) qq[j] = dot_smem<D>(cwg.w2t + (j * L + l) * S::YS, ln.a);
        // D. a_dg = W1^T (s t_a), gdg = W1^T (s q)
        float tmp[S::HL], adg[S::DL], gdg[S::DL];
#pragma unroll
        for (int j = 0; j < S::HL; ++j) tmp[j] = sg[j] * ta[j];
        store_frag<S::HL>(ln.dl + l * S::HL, tmp);
        __syncwarp();
#pragma unroll
        for (int d = 0; d < S::DL; ++d) adg[d] = dot_smem<H>(cwg.w1t + (d * L + l) * S::HS, ln.dl);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < S::HL; ++j) tmp[j] = sg[j] * qq[j];
        store_frag<S::HL>(ln.dl + l * S::HL, tmp);
        __syncwarp();
#pragma unroll
        for (int d = 0; d < S::DL; ++d) gdg[d] = dot_smem<H>(cwg.w1t + (d * L + l) * S::HS, ln.dl);
        __syncwarp();
        // E. r = p s into ln.dl: gbar = W2 r; dW2 -= h (g r^T)
        float rr[S::HL], gbar[S::DL];
#pragma unroll
        for (int j = 0; j < S::HL; ++j) rr[j] = pp[j] * sg[j];
        store_frag<S::HL>(ln.dl + l * S::HL, rr);
        __syncwarp();
        const float ms = -sc * hs;   // the second-order term enters f_corr = f - gdg with a minus sign, times the step
#pragma unroll
        for (int d = 0; d < S::DL; ++d) {
          gbar[d] = dot_line<H>(wg.w2[d], ln.dl, 0.f);
          axpy_line<H>(ag.w2[d], ms * gg[d], ln.dl);
        }
        // F. gbar into ln.a: hbar = W2^T gbar - 2 h p q; dW2 -= h (gbar h^T), db2 -= h gbar
        __syncwarp();
        store_frag<S::DL>(ln.a + l * S::DL, gbar);
        __syncwarp();
        float zb[S::HL];
#pragma unroll
        for (int j = 0; j < S::HL; ++j) {
          const float hb = dot_smem<D>(cwg.w2t + (j * L + l) * S::YS, ln.a) - 2.f * hg[j] * pp[j] * qq[j];
          zb[j] = sg[j] * hb;
        }
#pragma unroll
        for (int d = 0; d < S::DL; ++d) {
          axpy_line<H>(ag.w2[d], ms * gbar[d], ln.h);
          ag.b2[d] += ms * gbar[d];
        }
        // G. zbar into ln.dl: ybar_phi = W1^T zbar; dW1 -= h (zbar y^T + (s q) a^T), db1 -= h zbar
        __syncwarp();
        store_frag<S::HL>(ln.dl + l * S::HL, zb);
        __syncwarp();
        float yphi[S::DL];
#pragma unroll
        for (int d = 0; d < S::DL; ++d) yphi[d] = dot_smem<H>(cwg.w1t + (d * L + l) * S::HS, ln.dl);
        store_frag<S::DL>(ln.a + l * S::DL, a);   // (every lane has finished reading gbar from ln.a: the syncwarps above)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < S::HL; ++j) {
          axpy_line<D>(ag.w1[j], ms * zb[j], ln.y);
          axpy_line<D>(ag.w1[j], ms * sg[j] * qq[j], ln.a);
          ag.b1[j] += ms * zb[j];
        }
        // H. J_g^T c and the matching parameter terms for c = a_dg h + a v  (vjp is linear in its cotangent)
        float cc[S::DL], vb[S::DL];
#pragma unroll
        for (int d = 0; d < S::DL; ++d) cc[d] = fmaf(adg[d], hs, a[d] * v[d]);
        __syncwarp();
        mlp_vjp<D, H, L>(cwg, ln, l, hg, cc, sc, vb, ag);               // ln.y = y, ln.h = hg
        // I. Euler step of (y, a)
#pragma unroll
        for (int d = 0; d < S::DL; ++d) {
          const float ynew = y[d] - (f[d] - gdg[d]) * hs - gg[d] * v[d];
          a[d] = a[d] + (vyf[d] - yphi[d]) * hs + vb[d];
          y[d] = ynew;
        }
      }
      // y <- frames[i-1], a <- a + grad_frames[i-1]
      float go[S::DL];
#pragma unroll
      for (int d = 0; d < S::DL; ++d) { go[d] = 0.f; y[d] = 0.f; }
      if (valid) {
        load_frag<S::DL>(p.frames + sde_off(p.layout, i - 1, b, p.B, p.T, D) + l * S::DL, y);
        load_frag<S::DL>(p.grad_out + sde_off(p.layout, i - 1, b, p.B, p.T, D) + l * S::DL, go);
      }
#pragma unroll
      for (int d = 0; d < S::DL; ++d) a[d] += go[d];
    }
    if (valid) store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, a);
  }
  ReduceWs w1 = p.ws, w2 = p.ws;
  w2.partials = p.ws.partials + (size_t)gridDim.x * S::P;
  reduce_param_grads<D, H, L, WARPS, false>(af, s_red, w1, ss, p.grad_params, lane, warp, tid);
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

static int fill_cells(SdeCellArgs& c, const int* fwd_lo_host, int n_steps, const float* cell_sqrt_host, int R) {
  if (R < 1 || R > kSdeMaxCells || n_steps > kSdeMaxSteps) return GODE_ERR_T_TOO_LONG;
  c.R = R;
  for (int r = 0; r < R; ++r) c.rs[r] = cell_sqrt_host[r];
  if (fwd_lo_host)
    for (int k = 0; k <= n_steps; ++k) {
      if (fwd_lo_host[k] < 0 || fwd_lo_host[k] > R) return GODE_ERR_ARG;
      c.fwd_lo[k] = (short)fwd_lo_host[k];
    }
  return GODE_OK;
}

int sde_small_fwd(const float* y0, const float* const* fw, const float* const* gw, const float* h_host, int n_steps,
                  const int* out_step_host, const float* w0_host, const float* w1_host, int B, int D, int H, int T,
                  const float* dW, unsigned long long seed, long long traj_offset, int layout, float* out, float* states,
                  cudaStream_t st, const int* fwd_lo_host, const float* cell_sqrt_host, int R) {
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
  static SdeCellArgs no_cells{};
  if (fwd_lo_host) {   // increments summed over the cell grid (sdeint_adjoint with the stochastic adjoint)
    SdeCellArgs c{};
    if (int rc = fill_cells(c, fwd_lo_host, n_steps, cell_sqrt_host, R)) return rc;
    sde_em_fwd_kernel<16, 16, L, WARPS, true><<<grid, WARPS * 32, 0, st>>>(a, c);
  } else {
    sde_em_fwd_kernel<16, 16, L, WARPS, false><<<grid, WARPS * 32, 0, st>>>(a, no_cells);
  }
  return launch_status();
}

int sde_small_adjoint_bwd(const float* frames, const float* grad_out, const float* const* fw, const float* const* gw,
                          int n_rev, const float* h_rev_host, const int* rev_lo_host, const int* rev_hi_host,
                          const int* ibeg_host, const int* iend_host, const float* cell_sqrt_host, int R, int B, int D, int H,
                          int T, const float* dW, unsigned long long seed, long long traj_offset, int layout, float* grad_y0,
                          float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (n_rev < 1 || n_rev > kSdeMaxRev || T > kSdeMaxT) return GODE_ERR_T_TOO_LONG;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  SdeAdjArgs a{};
  SdeCellArgs c{};
  if (int rc = fill_cells(c, nullptr, 0, cell_sqrt_host, R)) return rc;
  a.frames = frames; a.grad_out = grad_out;
  a.fW1 = fw[0]; a.fb1 = fw[1]; a.fW2 = fw[2]; a.fb2 = fw[3];
  a.gW1 = gw[0]; a.gb1 = gw[1]; a.gW2 = gw[2]; a.gb2 = gw[3];
  a.dW = dW; a.seed = seed; a.traj_offset = traj_offset; a.B = B; a.T = T; a.layout = layout; a.n_rev = n_rev;
  a.grad_y0 = grad_y0; a.grad_params = grad_params;
  for (int n = 0; n < n_rev; ++n) {
    if (rev_lo_host[n] < 0 || rev_hi_host[n] > R || rev_lo_host[n] >= rev_hi_host[n]) return GODE_ERR_ARG;
    a.h[n] = h_rev_host[n]; a.lo[n] = (short)rev_lo_host[n]; a.hi[n] = (short)rev_hi_host[n];
  }
  for (int i = 0; i < T; ++i) {
    if (i >= 1 && (ibeg_host[i] < 0 || iend_host[i] > n_rev || ibeg_host[i] > iend_host[i])) return GODE_ERR_ARG;
    a.ibeg[i] = (short)(i >= 1 ? ibeg_host[i] : 0); a.iend[i] = (short)(i >= 1 ? iend_host[i] : 0);
  }
  constexpr int WARPS = 4, L = 16;
  using S = Shape<16, 16, L>;
  auto kern = sde_adjoint_bwd_kernel<16, 16, L, WARPS>;
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
  void* args[] = {(void*)&a, (void*)&c};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(WARPS * 32), args, smem, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
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
