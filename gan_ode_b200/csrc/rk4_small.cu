// rk4_small.cu — fixed-grid solvers for the small-field shapes, FP32: RK4 (3/8 rule, METHOD 0), Euler (1), midpoint (2).
//   rk4_fwd_kernel          torchdiffeq FixedGridODESolver.integrate + rk4_alt_step_func        
//   rk4_adjoint_bwd_kernel  torchdiffeq OdeintAdjointMethod.backward with method='rk4'          
//   rk4_backprop_bwd_kernel autograd through the same forward (discretise-then-optimise)        
// One launch per call: every trajectory's whole time loop runs inside the kernel.
#include "small_field.cuh"
#include "philox.cuh"
#include "launch.h"

namespace gode {

constexpr int kRk4 = GODE_METHOD_RK4, kEuler = GODE_METHOD_EULER, kMidpoint = GODE_METHOD_MIDPOINT;
constexpr float kOneThird = 0.33333334f;  // float(1/3): torchdiffeq multiplies fp32 tensors by the Python double 1/3

struct Rk4Args {
  const float *y0, *W1, *b1, *W2, *b2;
  const float* traj_in;     // bwd: stored forward trajectory
  const float* grad_traj;   // bwd: upstream gradient, same layout as traj
  float* traj;              // fwd output
  float* grad_y0;           // bwd output (B,D)
  float* grad_params;       // bwd output flat [W1|b1|W2|b2]
  ReduceWs ws;
  const float* dt_dev;      // device dt table, or nullptr -> dt_val
  int B, T, layout;
  // (B,T,D) layout only: row stride of traj / traj_in and of grad_traj in floats (0 = D).  Lets the codes live inside a wider
  // buffer — the motion columns of the generator's z = cat([z_content, z_motion], 1) (models/mocogan.py:264-267).
  int ld_traj, ld_grad;
  // fused sampler prologue (models/mocogan_ode.py:136-140): y0 = linear(x), x ~ N(0, I) from Philox stream 2,
  // linear = LeakyReLU(Wb LeakyReLU(Wa x + ba) + bb) (pre_Wa null: linear is nn.Identity, y0 = x)
  const float *pre_Wa, *pre_ba, *pre_Wb, *pre_bb;
  float pre_slope;
  int pre_hidden;           // 64 in the reference
  unsigned long long seed;
  long long traj_offset;    // global index of local trajectory 0
  const long long* traj_ids;  // optional (B): global trajectory index of every row (subset solves), else traj_offset + b
  float* noise_out;         // (B,D): the drawn x, kept for the pre-MLP's backward
  // options['step_size'] under odeint_adjoint: interval i (i = T-1 .. 1) is integrated with the sub-steps
  // sub_dt[sub_beg[i] .. sub_end[i]) (device arrays; signed like dt), carrying y along inside the interval
  const float* sub_dt;
  const int *sub_beg, *sub_end;
  float dt_val[GODE_MAX_HOST_STEPS];
};

__device__ __forceinline__ size_t traj_off(int layout, int s, int b, int B, int T, int D, int ld = 0) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * (size_t)(ld ? ld : D);
}

// ------------------------------------------------------------------------------------------------------------
constexpr int kPreHidden = 64;   // width of the reference's pre-MLP (models/mocogan_ode.py:123-131)

// FUSED: the sampler's prologue runs in the kernel — x from Philox, y0 = linear(x) — instead of reading y0.
template <int D, int H, int L, int WARPS, int METHOD, bool FUSED = false>
__global__ void __launch_bounds__(WARPS * 32) rk4_fwd_kernel(const __grid_constant__ Rk4Args p) {
  griddep_launch_dependents();   // a backward launched with PDL may stage its weights under this kernel's tail
  using S = Shape<D, H, L>;
  __shared__ __align__(16) float s_lines[WARPS * FwdLines<D, H, L>::kFloatsPerWarp];
  // fused prologue: pre-MLP weights (row-padded) and one hidden line per trajectory of the CTA
  constexpr int PS = D + kPad, QS = kPreHidden + kPad;
  __shared__ __align__(16) float s_pre[FUSED ? kPreHidden * PS + D * QS + kPreHidden + D + WARPS * S::G * QS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / L, l = lane % L;
  FwdLines<D, H, L> ln;
  ln.bind(s_lines + warp * FwdLines<D, H, L>::kFloatsPerWarp, g);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  float* s_wa = s_pre;                          // [64][PS]
  float* s_wb = s_wa + kPreHidden * PS;         // [D][QS]
  float* s_ba = s_wb + D * QS;                  // [64]
  float* s_bb = s_ba + kPreHidden;              // [D]
  float* s_hid = s_bb + D + (warp * S::G + g) * QS;   // this trajectory's hidden line
  if constexpr (FUSED) {
    if (p.pre_Wa) {
      for (int e = threadIdx.x; e < kPreHidden * D; e += WARPS * 32) {
        s_wa[(e / D) * PS + e % D] = p.pre_Wa[e];                         // Wa[j][i]
        s_wb[(e / kPreHidden) * QS + e % kPreHidden] = p.pre_Wb[e];       // Wb[d][j]
      }
      for (int e = threadIdx.x; e < kPreHidden; e += WARPS * 32) s_ba[e] = p.pre_ba[e];
      for (int e = threadIdx.x; e < D; e += WARPS * 32) s_bb[e] = p.pre_bb[e];
    }
    __syncthreads();
  }
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    float y[S::DL];
    if constexpr (FUSED) {
      // models/mocogan_ode.py:136-140: x = randn(B, D); x = self.linear(x)
      const unsigned long long gid = (unsigned long long)(p.traj_ids ? (valid ? p.traj_ids[b] : 0) : p.traj_offset + b);
      float x[S::DL];
      philox_normals<S::DL>(p.seed, gid, 0, l, x, 2u);
      if (valid && p.noise_out) store_frag<S::DL>(p.noise_out + (size_t)b * D + l * S::DL, x);
      if (p.pre_Wa) {
        __syncwarp();
        store_frag<S::DL>(ln.y + l * S::DL, x);
        __syncwarp();
        constexpr int JL = kPreHidden / L;   // hidden units of the pre-MLP per lane
        float hid[JL];
#pragma unroll
        for (int jl = 0; jl < JL; ++jl) {
          const int j = l * JL + jl;
          const float z = dot_smem<D>(s_wa + j * PS, ln.y) + s_ba[j];
          hid[jl] = z > 0.f ? z : p.pre_slope * z;
        }
        store_frag<JL>(s_hid + l * JL, hid);
        __syncwarp();
#pragma unroll
        for (int dl = 0; dl < S::DL; ++dl) {
          const int d = l * S::DL + dl;
          const float z = dot_smem<kPreHidden>(s_wb + d * QS, s_hid) + s_bb[d];
          y[dl] = z > 0.f ? z : p.pre_slope * z;
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int i = 0; i < S::DL; ++i) y[i] = x[i];
      }
    } else {
      if (valid) load_frag<S::DL>(p.y0 + (size_t)b * D + l * S::DL, y);
      else {
#pragma unroll
        for (int i = 0; i < S::DL; ++i) y[i] = 0.f;
      }
    }
    if (valid) store_frag<S::DL>(p.traj + traj_off(p.layout, 0, b, p.B, p.T, D, p.ld_traj) + l * S::DL, y);
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      float k1[S::DL], k2[S::DL], k3[S::DL], k4[S::DL], u[S::DL], hk[S::HL];
      mlp_forward<D, H, L>(w, ln, l, y, k1, hk);
      if constexpr (METHOD == kEuler) {          // fixed_grid.py::Euler: dy = dt * f(t0, y0)
#pragma unroll
        for (int i = 0; i < S::DL; ++i) y[i] = y[i] + dt * k1[i];
        if (valid) store_frag<S::DL>(p.traj + traj_off(p.layout, s + 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, y);
        continue;
      }
      if constexpr (METHOD == kMidpoint) {       // fixed_grid.py::Midpoint: y_mid = y0 + f(y0) * (dt/2); dy = dt * f(y_mid)
        const float half_dt = 0.5f * dt;
#pragma unroll
        for (int i = 0; i < S::DL; ++i) u[i] = y[i] + k1[i] * half_dt;
        mlp_forward<D, H, L>(w, ln, l, u, k2, hk);
#pragma unroll
        for (int i = 0; i < S::DL; ++i) y[i] = y[i] + dt * k2[i];
        if (valid) store_frag<S::DL>(p.traj + traj_off(p.layout, s + 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, y);
        continue;
      }
#pragma unroll
      for (int i = 0; i < S::DL; ++i) u[i] = y[i] + dt * k1[i] * kOneThird;
      mlp_forward<D, H, L>(w, ln, l, u, k2, hk);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) u[i] = y[i] + dt * (k2[i] - k1[i] * kOneThird);
      mlp_forward<D, H, L>(w, ln, l, u, k3, hk);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) u[i] = y[i] + dt * (k1[i] - k2[i] + k3[i]);
      mlp_forward<D, H, L>(w, ln, l, u, k4, hk);
#pragma unroll
      for (int i = 0; i < S::DL; ++i) y[i] = y[i] + (k1[i] + 3.f * (k2[i] + k3[i]) + k4[i]) * dt * 0.125f;
      if (valid) store_frag<S::DL>(p.traj + traj_off(p.layout, s + 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Continuous adjoint, one 3/8-rule step of the augmented system (y, a, theta_bar) per output interval, in
// reversed time: dy/ds = -f, da/ds = +a^T df/dy, dtheta_bar/ds = +a^T df/dtheta.  y is reset to the stored forward
// value at the start of every interval and grad_traj[i-1] is added to a at its end, exactly as adjoint.py does.
// theta_bar is linear in the stage contributions, so each lane keeps its rows of theta_bar in registers across ALL
// intervals and ALL its trajectories and the grid reduces once at the end.
template <int D, int H, int L, int WARPS, int METHOD, bool SUB = false>
__global__ void __launch_bounds__(WARPS * 32) rk4_adjoint_bwd_kernel(const __grid_constant__ Rk4Args p) {
  using S = Shape<D, H, L>;
  using BL = BwdLines<D, H, L>;
  extern __shared__ __align__(16) float smem[];
  float* s_lines = smem;                                          // WARPS * BL::kFloatsPerWarp
  float* s_cw = s_lines + WARPS * BL::kFloatsPerWarp;             // ColWeights::kFloats
  float* s_red = s_cw + ColWeights<D, H, L>::kFloats;             // WARPS * P
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  BL ln;
  ln.bind(s_lines + warp * BL::kFloatsPerWarp, g);
  ColWeights<D, H, L> cw;
  cw.bind(s_cw);
  cw.stage(p.W1, p.W2, tid, WARPS * 32);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  GradAcc<D, H, L> acc;
  acc.zero();
  griddep_wait();        // PDL launch: the weight prologue above overlapped the forward's tail; its outputs are read below
  SyncState ss;
  ss.begin(p.ws.gs);
  __syncthreads();
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    float a[S::DL];
#pragma unroll
    for (int i = 0; i < S::DL; ++i) a[i] = 0.f;
    // stored state and upstream gradient of an interval are fetched one interval ahead (they are cold global reads, and
    // at small batch the kernel is a chain of dependent latencies)
    float yn[S::DL], gn[S::DL];
#pragma unroll
    for (int i = 0; i < S::DL; ++i) { yn[i] = 0.f; gn[i] = 0.f; }
    if (valid) {
      load_frag<S::DL>(p.grad_traj + traj_off(p.layout, p.T - 1, b, p.B, p.T, D, p.ld_grad) + l * S::DL, a);
      load_frag<S::DL>(p.traj_in + traj_off(p.layout, p.T - 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, yn);
      load_frag<S::DL>(p.grad_traj + traj_off(p.layout, p.T - 2, b, p.B, p.T, D, p.ld_grad) + l * S::DL, gn);
    }
    const float sc = valid ? 1.f : 0.f;  // padded lanes must not pollute theta_bar
    // one step of size dt of the augmented system (y, a, theta_bar) in reversed time; y is advanced only when the interval
    // is sub-stepped (SUB: options['step_size'] under odeint_adjoint) — with one step per interval it is reloaded anyway
    auto aug_step = [&](const float dt, float (&y)[S::DL]) {
      const float c18 = dt * 0.125f, c38 = 3.f * c18;
      float f[S::DL], v[S::DL], hk[S::HL], uy[S::DL], ua[S::DL];
      float k1y[S::DL], k1a[S::DL], k2y[S::DL], k2a[S::DL], k3y[S::DL], k3a[S::DL];
      if constexpr (METHOD == kEuler) {          // one Euler step of the augmented system from (y_i, a_i)
        mlp_forward<D, H, L>(w, ln, l, y, f, hk);
        mlp_vjp<D, H, L>(cw, ln, l, hk, a, sc * dt, v, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          a[c] = a[c] + dt * v[c];
          if constexpr (SUB) y[c] = y[c] - dt * f[c];
        }
        return;
      }
      if constexpr (METHOD == kMidpoint) {       // X1 = X0 + dt F(X0 + dt/2 F(X0)): only the second stage feeds theta_bar
        const float half_dt = 0.5f * dt;
        mlp_forward<D, H, L>(w, ln, l, y, f, hk);
        mlp_vjp<D, H, L>(cw, ln, l, hk, a, 0.f, v, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) { uy[c] = y[c] - f[c] * half_dt; ua[c] = a[c] + v[c] * half_dt; }
        mlp_forward<D, H, L>(w, ln, l, uy, f, hk);
        mlp_vjp<D, H, L>(cw, ln, l, hk, ua, sc * dt, v, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          a[c] = a[c] + dt * v[c];
          if constexpr (SUB) y[c] = y[c] - dt * f[c];
        }
        return;
      }
      // stage 1
      mlp_forward<D, H, L>(w, ln, l, y, f, hk);
      mlp_vjp<D, H, L>(cw, ln, l, hk, a, sc * c18, v, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { k1y[c] = -f[c]; k1a[c] = v[c]; uy[c] = y[c] + dt * k1y[c] * kOneThird; ua[c] = a[c] + dt * k1a[c] * kOneThird; }
      // stage 2
      mlp_forward<D, H, L>(w, ln, l, uy, f, hk);
      mlp_vjp<D, H, L>(cw, ln, l, hk, ua, sc * c38, v, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { k2y[c] = -f[c]; k2a[c] = v[c]; uy[c] = y[c] + dt * (k2y[c] - k1y[c] * kOneThird); ua[c] = a[c] + dt * (k2a[c] - k1a[c] * kOneThird); }
      // stage 3
      mlp_forward<D, H, L>(w, ln, l, uy, f, hk);
      mlp_vjp<D, H, L>(cw, ln, l, hk, ua, sc * c38, v, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { k3y[c] = -f[c]; k3a[c] = v[c]; uy[c] = y[c] + dt * (k1y[c] - k2y[c] + k3y[c]); ua[c] = a[c] + dt * (k1a[c] - k2a[c] + k3a[c]); }
      // stage 4
      mlp_forward<D, H, L>(w, ln, l, uy, f, hk);
      mlp_vjp<D, H, L>(cw, ln, l, hk, ua, sc * c18, v, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        a[c] = a[c] + (k1a[c] + 3.f * (k2a[c] + k3a[c]) + v[c]) * dt * 0.125f;
        if constexpr (SUB) y[c] = y[c] + (k1y[c] + 3.f * (k2y[c] + k3y[c]) - f[c]) * dt * 0.125f;
      }
    };
    for (int i = p.T - 1; i >= 1; --i) {
      float y[S::DL], gprev[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { y[c] = yn[c]; gprev[c] = gn[c]; }
      if (valid && i > 1) {
        load_frag<S::DL>(p.traj_in + traj_off(p.layout, i - 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, yn);
        load_frag<S::DL>(p.grad_traj + traj_off(p.layout, i - 2, b, p.B, p.T, D, p.ld_grad) + l * S::DL, gn);
      }
      if constexpr (SUB) {
        for (int q = p.sub_beg[i]; q < p.sub_end[i]; ++q) aug_step(p.sub_dt[q], y);
      } else {
        aug_step(dtp[i - 1], y);
      }
#pragma unroll
      for (int c = 0; c < S::DL; ++c) a[c] += gprev[c];
    }
    if (valid) store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, a);
  }
  reduce_param_grads<D, H, L, WARPS>(acc, s_red, p.ws, ss, p.grad_params, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
}

// ------------------------------------------------------------------------------------------------------------
// Exact reverse-mode through the forward's arithmetic.  Per step the four stage inputs u_s and tanh vectors are
// recomputed from the stored y_s (bit-identical to the forward), then the stages are walked 4..1.
template <int D, int H, int L, int WARPS, int METHOD>
__global__ void __launch_bounds__(WARPS * 32) rk4_backprop_bwd_kernel(const __grid_constant__ Rk4Args p) {
  using S = Shape<D, H, L>;
  using BL = BwdLines<D, H, L>;
  extern __shared__ __align__(16) float smem[];
  float* s_lines = smem;
  float* s_cw = s_lines + WARPS * BL::kFloatsPerWarp;
  float* s_red = s_cw + ColWeights<D, H, L>::kFloats;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  BL ln;
  ln.bind(s_lines + warp * BL::kFloatsPerWarp, g);
  ColWeights<D, H, L> cw;
  cw.bind(s_cw);
  cw.stage(p.W1, p.W2, tid, WARPS * 32);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  GradAcc<D, H, L> acc;
  acc.zero();
  griddep_wait();        // PDL launch: the weight prologue above overlapped the forward's tail; its outputs are read below
  SyncState ss;
  ss.begin(p.ws.gs);
  __syncthreads();
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    const float sc = valid ? 1.f : 0.f;
    float yb[S::DL];  // cotangent of y_{s+1}
#pragma unroll
    for (int c = 0; c < S::DL; ++c) yb[c] = 0.f;
    float yn[S::DL], gn[S::DL];  // next step's stored state / upstream gradient, fetched one step ahead
#pragma unroll
    for (int c = 0; c < S::DL; ++c) { yn[c] = 0.f; gn[c] = 0.f; }
    if (valid) {
      load_frag<S::DL>(p.grad_traj + traj_off(p.layout, p.T - 1, b, p.B, p.T, D, p.ld_grad) + l * S::DL, yb);
      load_frag<S::DL>(p.traj_in + traj_off(p.layout, p.T - 2, b, p.B, p.T, D, p.ld_traj) + l * S::DL, yn);
      load_frag<S::DL>(p.grad_traj + traj_off(p.layout, p.T - 2, b, p.B, p.T, D, p.ld_grad) + l * S::DL, gn);
    }
    for (int s = p.T - 2; s >= 0; --s) {
      const float dt = dtp[s];
      float y[S::DL], gs[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { y[c] = yn[c]; gs[c] = gn[c]; }
      if (valid && s > 0) {
        load_frag<S::DL>(p.traj_in + traj_off(p.layout, s - 1, b, p.B, p.T, D, p.ld_traj) + l * S::DL, yn);
        load_frag<S::DL>(p.grad_traj + traj_off(p.layout, s - 1, b, p.B, p.T, D, p.ld_grad) + l * S::DL, gn);
      }
      // recompute (same expressions as the forward kernel)
      float k1[S::DL], k2[S::DL], k3[S::DL], k4[S::DL], u2[S::DL], u3[S::DL], u4[S::DL];
      float h1[S::HL], h2[S::HL], h3[S::HL], h4[S::HL];
      mlp_forward<D, H, L>(w, ln, l, y, k1, h1);
      if constexpr (METHOD == kEuler) {          // y1 = y0 + dt f(y0)
        float cot[S::DL], ub[S::DL];
#pragma unroll
        for (int c = 0; c < S::DL; ++c) cot[c] = dt * yb[c];
        mlp_vjp<D, H, L>(cw, ln, l, h1, cot, sc, ub, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) yb[c] += ub[c] + gs[c];
        continue;
      }
      if constexpr (METHOD == kMidpoint) {       // u = y0 + (dt/2) f(y0); y1 = y0 + dt f(u)
        const float half_dt = 0.5f * dt;
        float cot[S::DL], ub[S::DL];
#pragma unroll
        for (int c = 0; c < S::DL; ++c) u2[c] = y[c] + k1[c] * half_dt;
        mlp_forward<D, H, L>(w, ln, l, u2, k2, h2);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) cot[c] = dt * yb[c];
        mlp_vjp<D, H, L>(cw, ln, l, h2, cot, sc, ub, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) { yb[c] += ub[c]; cot[c] = half_dt * ub[c]; }
        regather<D, H, L>(ln, l, y, h1);
        mlp_vjp<D, H, L>(cw, ln, l, h1, cot, sc, ub, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) yb[c] += ub[c] + gs[c];
        continue;
      }
#pragma unroll
      for (int c = 0; c < S::DL; ++c) u2[c] = y[c] + dt * k1[c] * kOneThird;
      mlp_forward<D, H, L>(w, ln, l, u2, k2, h2);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) u3[c] = y[c] + dt * (k2[c] - k1[c] * kOneThird);
      mlp_forward<D, H, L>(w, ln, l, u3, k3, h3);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) u4[c] = y[c] + dt * (k1[c] - k2[c] + k3[c]);
      mlp_forward<D, H, L>(w, ln, l, u4, k4, h4);  // leaves u4 / h4 in the lines
      // reverse
      const float c18 = dt * 0.125f, c38 = 3.f * c18, dt3 = dt * kOneThird;
      float kb1[S::DL], kb2[S::DL], kb3[S::DL], kb4[S::DL], ub[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { kb1[c] = c18 * yb[c]; kb2[c] = c38 * yb[c]; kb3[c] = c38 * yb[c]; kb4[c] = c18 * yb[c]; }
      mlp_vjp<D, H, L>(cw, ln, l, h4, kb4, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { yb[c] += ub[c]; kb1[c] += dt * ub[c]; kb2[c] -= dt * ub[c]; kb3[c] += dt * ub[c]; }
      regather<D, H, L>(ln, l, u3, h3);
      mlp_vjp<D, H, L>(cw, ln, l, h3, kb3, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { yb[c] += ub[c]; kb2[c] += dt * ub[c]; kb1[c] -= dt3 * ub[c]; }
      regather<D, H, L>(ln, l, u2, h2);
      mlp_vjp<D, H, L>(cw, ln, l, h2, kb2, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { yb[c] += ub[c]; kb1[c] += dt3 * ub[c]; }
      regather<D, H, L>(ln, l, y, h1);
      mlp_vjp<D, H, L>(cw, ln, l, h1, kb1, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) yb[c] += ub[c] + gs[c];
      (void)k4;
    }
    if (valid) store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, yb);
  }
  reduce_param_grads<D, H, L, WARPS>(acc, s_red, p.ws, ss, p.grad_params, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
}

// ------------------------------------------------------------------------------------------------------------
// host side
template <int D, int H, int L, int WARPS>
static size_t bwd_smem_bytes() {
  return sizeof(float) * (WARPS * BwdLines<D, H, L>::kFloatsPerWarp + ColWeights<D, H, L>::kFloats + WARPS * Shape<D, H, L>::P);
}

static int fill_dt(Rk4Args& a, const float* dt, int dt_on_device, int T) {
  if (dt_on_device) { a.dt_dev = dt; return GODE_OK; }
  if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
  a.dt_dev = nullptr;
  for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  return GODE_OK;
}

template <int D, int H, int L, int METHOD>
static int launch_rk4_fwd(Rk4Args& a, cudaStream_t st) {
  constexpr int WARPS = 4;
  using S = Shape<D, H, L>;
  const int per_cta = WARPS * S::G;
  int grid = (a.B + per_cta - 1) / per_cta;
  const int cap = sm_count() * 16;  // 16 CTAs of 128 threads fill an SM; beyond that, loop
  if (grid > cap) grid = cap;
  rk4_fwd_kernel<D, H, L, WARPS, METHOD><<<grid, WARPS * 32, 0, st>>>(a);
  return launch_status();
}

template <int D, int H, int L, bool ADJOINT, int METHOD, bool SUB = false>
static int launch_rk4_bwd(Rk4Args& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  constexpr int WARPS = 4;
  using S = Shape<D, H, L>;
  auto kern = ADJOINT ? rk4_adjoint_bwd_kernel<D, H, L, WARPS, METHOD, SUB> : rk4_backprop_bwd_kernel<D, H, L, WARPS, METHOD>;
  const size_t smem = bwd_smem_bytes<D, H, L, WARPS>();
  cudaError_t e;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(1000 + (int)e);
  }
  static int limit_cache = 0;
  int cap = coop_limit(kern, WARPS * 32, smem, limit_cache);
  if (cap <= 0) return GODE_ERR_COOP;
  if (cap > bwd_grid_cap()) cap = bwd_grid_cap();
  const int per_cta = WARPS * S::G;
  int grid = (a.B + per_cta - 1) / per_cta;
  if (grid > cap) grid = cap;
  if (ws_bytes < bwd_workspace_bytes(S::P)) return GODE_ERR_WORKSPACE;
  grid_sync_bind(a.ws.gs, workspace);
  a.ws.partials = reinterpret_cast<float*>(ws_scratch(workspace));
  void* args[] = {(void*)&a};
  e = coop_launch(kern, grid, WARPS * 32, args, smem, st, (thread_launch_flags() & GODE_LAUNCH_PDL_BWD) != 0);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

int rk4_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                  int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st, int method) {
  Rk4Args a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.B = B; a.T = T; a.layout = out_layout;
  if (int rc = fill_dt(a, dt, dt_on_device, T)) return rc;
  if (D == 16 && H == 16) {
    if (method == kEuler) return launch_rk4_fwd<16, 16, 8, kEuler>(a, st);
    if (method == kMidpoint) return launch_rk4_fwd<16, 16, 8, kMidpoint>(a, st);
    return launch_rk4_fwd<16, 16, 8, kRk4>(a, st);
  }
  return GODE_ERR_SHAPE;
}

// SURVEY §8 f2 — the sampler fused around the solve (opt-in: it changes which random numbers are consumed).
int rk4_small_fused_sampler_fwd(const float* pre_Wa, const float* pre_ba, const float* pre_Wb, const float* pre_bb,
                                float pre_slope, int pre_hidden, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                                unsigned long long seed, long long traj_offset, const long long* traj_ids, int out_layout,
                                float* out, int ld_out, float* noise_out, cudaStream_t st) {
  if (pre_Wa && pre_hidden != kPreHidden) return GODE_ERR_SHAPE;
  Rk4Args a{};
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = out; a.B = B; a.T = T; a.layout = out_layout; a.ld_traj = ld_out;
  a.pre_Wa = pre_Wa; a.pre_ba = pre_ba; a.pre_Wb = pre_Wb; a.pre_bb = pre_bb; a.pre_slope = pre_slope; a.pre_hidden = pre_hidden;
  a.seed = seed; a.traj_offset = traj_offset; a.traj_ids = traj_ids; a.noise_out = noise_out;
  if (int rc = fill_dt(a, dt, dt_on_device, T)) return rc;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  constexpr int WARPS = 4;
  using S = Shape<16, 16, 8>;
  const int per_cta = WARPS * S::G;
  int grid = (B + per_cta - 1) / per_cta;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  rk4_fwd_kernel<16, 16, 8, WARPS, kRk4, true><<<grid, WARPS * 32, 0, st>>>(a);
  return launch_status();
}

int rk4_small_bwd(bool adjoint, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                  const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                  int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st,
                  int method, const GodeWorld* xchg, int ld_traj, int ld_grad, const SubSteps* sub) {
  Rk4Args a{};
  a.ld_traj = ld_traj; a.ld_grad = ld_grad;
  if (xchg) {  // all-reduce over ranks fused into the reduction tail (small_field.cuh::reduce_param_grads)
    a.ws.w_rank = xchg->rank; a.ws.w_world = xchg->world; a.ws.w_ctr = xchg->launch_ctr;
    a.ws.w_slots = reinterpret_cast<unsigned long long* const*>(xchg->slots_dev);
  }
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj_in = traj; a.grad_traj = grad_traj; a.grad_y0 = grad_y0;
  a.grad_params = grad_params; a.B = B; a.T = T; a.layout = layout;
  if (sub) {   // sub-stepped adjoint: the step sizes come from the device tables
    if (!adjoint || !sub->dt || !sub->beg || !sub->end) return GODE_ERR_ARG;
    if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
    a.sub_dt = sub->dt; a.sub_beg = sub->beg; a.sub_end = sub->end;
    if (method == kEuler) return launch_rk4_bwd<16, 16, 8, true, kEuler, true>(a, workspace, ws_bytes, st);
    if (method == kMidpoint) return launch_rk4_bwd<16, 16, 8, true, kMidpoint, true>(a, workspace, ws_bytes, st);
    return launch_rk4_bwd<16, 16, 8, true, kRk4, true>(a, workspace, ws_bytes, st);
  }
  if (int rc = fill_dt(a, dt, dt_on_device, T)) return rc;
  if (D == 16 && H == 16) {
    if (method == kEuler)
      return adjoint ? launch_rk4_bwd<16, 16, 8, true, kEuler>(a, workspace, ws_bytes, st)
                     : launch_rk4_bwd<16, 16, 8, false, kEuler>(a, workspace, ws_bytes, st);
    if (method == kMidpoint)
      return adjoint ? launch_rk4_bwd<16, 16, 8, true, kMidpoint>(a, workspace, ws_bytes, st)
                     : launch_rk4_bwd<16, 16, 8, false, kMidpoint>(a, workspace, ws_bytes, st);
    return adjoint ? launch_rk4_bwd<16, 16, 8, true, kRk4>(a, workspace, ws_bytes, st)
                   : launch_rk4_bwd<16, 16, 8, false, kRk4>(a, workspace, ws_bytes, st);
  }
  return GODE_ERR_SHAPE;
}

}  // namespace gode
