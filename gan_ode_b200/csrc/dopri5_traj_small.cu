// dopri5_traj_small.cu — adaptive dopri5 with PER-TRAJECTORY step control (opt-in, GODE_NORM_TRAJ).
//
// torchdiffeq shares one (t, dt) over the whole batch (dopri5_small.cu reproduces that).  Here every trajectory carries
// its own (t, dt), error norm = RMS over its own D components, own accept/reject sequence — i.e. exactly what
// torchdiffeq computes when it is called with B = 1 per trajectory (that is the parity oracle).  No grid-wide
// reduction: the norm is an xor-shuffle over the L lanes of the trajectory, so the kernel needs no cooperative launch,
// scales to any batch, and a stiff trajectory no longer shortens everyone's steps.
// The 32/L trajectories of a warp advance in lock-step per ATTEMPT (each with its own dt); a warp is done when all its
// trajectories are; finished ones idle with their state frozen.
#include "dopri5_common.cuh"

namespace gode {

constexpr int kTrajMaxT = 256;

struct Dp5TrajArgs {
  const float *y0, *W1, *b1, *W2, *b2;
  const float* grad_traj;
  float* traj;
  float* grad_y0;
  float* grad_params;
  ReduceWs ws;
  GodeStepLog* log;       // status = OR over trajectories; counts = max over trajectories
  int32_t* mailbox;       // mapped host int (gode_set_status_mailbox) or null
  int32_t* n_acc;         // (B)
  int32_t* n_att;         // (B)
  double* att_dt; float* att_er; uint8_t* att_acc;   // (log_capacity, B) or null
  float* ckpt;            // (ckpt_capacity, B, D)
  double* acc_t0; double* acc_dt;                     // (ckpt_capacity, B)
  GodeAdaptiveOpts o;
  int B, T, layout;
  double t[kTrajMaxT];
};

__device__ __forceinline__ size_t tj_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

// sum over the L lanes of a trajectory (aligned L-lane segment of the warp)
template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = L / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int D, int H, int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) dopri5_traj_fwd_kernel(const __grid_constant__ Dp5TrajArgs p) {
  using S = Shape<D, H, L>;
  __shared__ __align__(16) float s_lines[WARPS * FwdLines<D, H, L>::kFloatsPerWarp];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  FwdLines<D, H, L> ln;
  ln.bind(s_lines + warp * FwdLines<D, H, L>::kFloatsPerWarp, g);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  const float rtol32 = (float)p.o.rtol, atol32 = (float)p.o.atol;
  const float invD = 1.f / (float)D;
  const int stride = gridDim.x * WARPS * S::G;
  int st_or = 0, max_att = 0, max_acc = 0, max_nfe = 0;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    float y0[S::DL], k[7][S::DL], hk[S::HL];
#pragma unroll
    for (int c = 0; c < S::DL; ++c) y0[c] = 0.f;
    if (valid) {
      load_frag<S::DL>(p.y0 + (size_t)b * D + l * S::DL, y0);
      store_frag<S::DL>(p.traj + tj_off(p.layout, 0, b, p.B, p.T, D) + l * S::DL, y0);
    }
    field<D, H, L>(w, ln, l, p.o.fsign, y0, k[0], hk);
    int nfe = 1, status = 0;
    double t0 = p.t[0], dt;
    {  // misc.py::_select_initial_step with the trajectory's own norms
      float scale[S::DL], s0 = 0.f, s1 = 0.f, bad = 0.f;
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        scale[c] = atol32 + fabsf(y0[c]) * rtol32;
        const float r0 = y0[c] / scale[c], r1 = k[0][c] / scale[c];
        s0 = fmaf(r0, r0, s0); s1 = fmaf(r1, r1, s1);
        if (!isfinite(y0[c])) bad = 1.f;
      }
      s0 = group_sum<L>(s0); s1 = group_sum<L>(s1); bad = group_sum<L>(bad);
      if (bad > 0.f && valid) status |= GODE_ST_NONFINITE;
      if (p.o.first_step > 0.0) {
        dt = p.o.first_step;
      } else {
        const float d0 = sqrtf(s0 * invD), d1 = sqrtf(s1 * invD);
        const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
        float u[S::DL], f1[S::DL];
#pragma unroll
        for (int c = 0; c < S::DL; ++c) u[c] = y0[c] + h0 * k[0][c];
        field<D, H, L>(w, ln, l, p.o.fsign, u, f1, hk);
        nfe++;
        float s2 = 0.f;
#pragma unroll
        for (int c = 0; c < S::DL; ++c) { const float r = (f1[c] - k[0][c]) / scale[c]; s2 = fmaf(r, r, s2); }
        s2 = group_sum<L>(s2);
        const float d2 = sqrtf(s2 * invD) / h0;
        const float h1 = (d1 <= 1e-15f && d2 <= 1e-15f) ? fmaxf(1e-6f, h0 * 1e-3f) : powf(0.01f / fmaxf(d1, d2), 0.2f);
        dt = (double)fminf(100.f * h0, h1);
      }
    }
    int iout = 1, n_att = 0, n_acc = 0, n_steps = 0;
    bool done = !valid || status != 0;
    while (__any_sync(0xffffffffu, !done)) {
      const bool act = !done;
      if (act && n_steps >= p.o.max_num_steps) { status |= GODE_ST_MAX_STEPS; done = true; }
      if (act && !done && !(t0 + dt > t0)) { status |= GODE_ST_DT_UNDERFLOW; done = true; }
      const bool go = act && !done;
      const double t1 = t0 + dt;
      const float dt32 = go ? (float)dt : 0.f;
      float u[S::DL];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float s = k[0][c] * (kBeta[i][0] * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) s = fmaf(k[j][c], kBeta[i][j] * dt32, s);
          u[c] = y0[c] + s;
        }
        field<D, H, L>(w, ln, l, p.o.fsign, u, k[i + 1], hk);
      }
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        float e = k[0][c] * (dt32 * kCErr[0]);
#pragma unroll
        for (int j = 2; j < 7; ++j) e = fmaf(k[j][c], dt32 * kCErr[j], e);
        const float r = e / (atol32 + rtol32 * fmaxf(fabsf(y0[c]), fabsf(u[c])));
        ss = fmaf(r, r, ss);
      }
      const float er = sqrtf(group_sum<L>(ss) * invD);
      if (go) {
        nfe += 6;
        bool accept = er <= 1.f;
        if (dt > p.o.max_step) accept = false;
        if (dt <= p.o.min_step) accept = true;
        if (l == 0 && n_att < p.o.log_capacity && p.att_dt) {
          p.att_dt[(size_t)n_att * p.B + b] = dt; p.att_er[(size_t)n_att * p.B + b] = er;
          p.att_acc[(size_t)n_att * p.B + b] = accept ? 1 : 0;
        }
        if (accept) {
          if (p.o.ckpt_capacity > 0) {
            if (n_acc < p.o.ckpt_capacity) {
              store_frag<S::DL>(p.ckpt + ((size_t)n_acc * p.B + b) * D + l * S::DL, y0);
              if (l == 0) { p.acc_t0[(size_t)n_acc * p.B + b] = t0; p.acc_dt[(size_t)n_acc * p.B + b] = dt; }
            } else {
              status |= GODE_ST_CKPT_OVERFLOW;
            }
          }
          if (iout < p.T && p.t[iout] <= t1) {
            float ca[S::DL], cb[S::DL], cc[S::DL], cd[S::DL];
#pragma unroll
            for (int c = 0; c < S::DL; ++c) {
              float m = k[0][c] * (dt32 * kCMid[0]);
#pragma unroll
              for (int j = 2; j < 7; ++j) m = fmaf(k[j][c], dt32 * kCMid[j], m);
              const float ymid = y0[c] + m, f0 = k[0][c], f1 = k[6][c], y1 = u[c];
              ca[c] = 2.f * dt32 * (f1 - f0) - 8.f * (y1 + y0[c]) + 16.f * ymid;
              cb[c] = dt32 * (5.f * f0 - 3.f * f1) + 18.f * y0[c] + 14.f * y1 - 32.f * ymid;
              cc[c] = dt32 * (f1 - 4.f * f0) - 11.f * y0[c] - 5.f * y1 + 16.f * ymid;
              cd[c] = dt32 * f0;
            }
            const double inv_span = 1.0 / (t1 - t0);
            while (iout < p.T && p.t[iout] <= t1) {
              const float x = (float)((p.t[iout] - t0) * inv_span);
              float o[S::DL];
#pragma unroll
              for (int c = 0; c < S::DL; ++c) {
                float tot = y0[c] + x * cd[c];
                float xp = x * x;
                tot = tot + xp * cc[c];
                xp = xp * x;
                tot = tot + xp * cb[c];
                xp = xp * x;
                tot = tot + xp * ca[c];
                o[c] = tot;
              }
              store_frag<S::DL>(p.traj + tj_off(p.layout, iout, b, p.B, p.T, D) + l * S::DL, o);
              ++iout;
              n_steps = -1;
            }
          }
#pragma unroll
          for (int c = 0; c < S::DL; ++c) { y0[c] = u[c]; k[0][c] = k[6][c]; }
          t0 = t1;
          ++n_acc;
        }
        dt = optimal_step(dt, er, p.o);
        dt = fmin(fmax(dt, p.o.min_step), p.o.max_step);
        ++n_att;
        ++n_steps;
        if (iout >= p.T || status != 0) done = true;
      }
    }
    if (valid && l == 0) { p.n_acc[b] = n_acc; p.n_att[b] = n_att; }
    if (valid) { st_or |= status; max_att = max(max_att, n_att); max_acc = max(max_acc, n_acc); max_nfe = max(max_nfe, nfe); }
  }
  if (st_or) {
    atomicOr(&p.log->status, st_or);
    if (p.mailbox && (threadIdx.x & 31) == 0) *reinterpret_cast<volatile int32_t*>(p.mailbox) = st_or;
  }
  atomicMax(&p.log->n_attempts, max_att);
  atomicMax(&p.log->n_accepted, max_acc);
  atomicMax(&p.log->nfe, max_nfe);
}

template <int D, int H, int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) dopri5_traj_bwd_kernel(const __grid_constant__ Dp5TrajArgs p) {
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
  const float poison = p.log->status != 0 ? __int_as_float(0x7fc00000) : 0.f;
  acc.fill(poison);
  SyncState ss;
  ss.begin(p.ws.gs);
  __syncthreads();
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    const int n_mine = valid ? min(p.n_acc[b], p.o.ckpt_capacity) : 0;
    int n_warp = n_mine;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n_warp = max(n_warp, __shfl_xor_sync(0xffffffffu, n_warp, off));
    float ybar[S::DL], fbar[S::DL];
#pragma unroll
    for (int c = 0; c < S::DL; ++c) { ybar[c] = poison; fbar[c] = 0.f; }
    int iout = p.T - 1;
    for (int it = 0; it < n_warp; ++it) {
      const int s = n_mine - 1 - it;          // this trajectory's step for this round (its own last step first)
      const bool act = s >= 0;
      const float sc = act ? 1.f : 0.f;
      const double t0 = act ? p.acc_t0[(size_t)s * p.B + b] : 0.0, dtd = act ? p.acc_dt[(size_t)s * p.B + b] : 0.0;
      const double t1 = t0 + dtd;
      const float dt32 = (float)dtd;
      float y0[S::DL], k[7][S::DL], h[7][S::HL], u[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) y0[c] = 0.f;
      if (act) load_frag<S::DL>(p.ckpt + ((size_t)s * p.B + b) * D + l * S::DL, y0);
      field<D, H, L>(w, ln, l, p.o.fsign, y0, k[0], h[0]);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float sum = k[0][c] * (kBeta[i][0] * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], kBeta[i][j] * dt32, sum);
          u[c] = y0[c] + sum;
        }
        field<D, H, L>(w, ln, l, p.o.fsign, u, k[i + 1], h[i + 1]);
      }
      float y0b[S::DL], ymb[S::DL], f0b[S::DL], kb[7][S::DL], yb1[S::DL], fb1[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { y0b[c] = 0.f; ymb[c] = 0.f; f0b[c] = 0.f; yb1[c] = ybar[c]; fb1[c] = fbar[c]; }
      if (act) {
        const double inv_span = 1.0 / (t1 - t0);
        while (iout >= 1 && p.t[iout] > t0) {
          float gout[S::DL];
          load_frag<S::DL>(p.grad_traj + tj_off(p.layout, iout, b, p.B, p.T, D) + l * S::DL, gout);
          const float x = (float)((p.t[iout] - t0) * inv_span);
          const float p2 = x * x, p3 = p2 * x, p4 = p3 * x;
          const float cy0 = 1.f - 11.f * p2 + 18.f * p3 - 8.f * p4, cy1 = -5.f * p2 + 14.f * p3 - 8.f * p4;
          const float cym = 16.f * p2 - 32.f * p3 + 16.f * p4;
          const float cf0 = dt32 * (x - 4.f * p2 + 5.f * p3 - 2.f * p4), cf1 = dt32 * (p2 - 3.f * p3 + 2.f * p4);
#pragma unroll
          for (int c = 0; c < S::DL; ++c) {
            y0b[c] = fmaf(cy0, gout[c], y0b[c]); yb1[c] = fmaf(cy1, gout[c], yb1[c]); ymb[c] = fmaf(cym, gout[c], ymb[c]);
            f0b[c] = fmaf(cf0, gout[c], f0b[c]); fb1[c] = fmaf(cf1, gout[c], fb1[c]);
          }
          --iout;
        }
      }
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        y0b[c] += ymb[c];
#pragma unroll
        for (int j = 0; j < 7; ++j) kb[j][c] = (dt32 * kCMid[j]) * ymb[c];
        kb[6][c] += fb1[c];
        kb[0][c] += f0b[c];
      }
      float ub[S::DL], cot[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) cot[c] = p.o.fsign * kb[6][c];
      mlp_vjp<D, H, L>(cw, ln, l, h[6], cot, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        ub[c] += yb1[c];
        y0b[c] += ub[c];
#pragma unroll
        for (int j = 0; j < 6; ++j) kb[j][c] = fmaf(kBeta[5][j] * dt32, ub[c], kb[j][c]);
      }
#pragma unroll
      for (int i = 4; i >= 0; --i) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float sum = k[0][c] * (kBeta[i][0] * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], kBeta[i][j] * dt32, sum);
          u[c] = y0[c] + sum;
        }
        regather<D, H, L>(ln, l, u, h[i + 1]);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) cot[c] = p.o.fsign * kb[i + 1][c];
        mlp_vjp<D, H, L>(cw, ln, l, h[i + 1], cot, sc, ub, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          y0b[c] += ub[c];
#pragma unroll
          for (int j = 0; j <= i; ++j) kb[j][c] = fmaf(kBeta[i][j] * dt32, ub[c], kb[j][c]);
        }
      }
      // k1: FSAL hands its cotangent to the previous step, except for the trajectory's first step (s == 0) where it is
      // the field at y(t0).  Both cases execute the same VJP (warp-synchronous), the inactive case with a zero cotangent.
      const bool first = act && s == 0;
      regather<D, H, L>(ln, l, y0, h[0]);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) cot[c] = first ? p.o.fsign * kb[0][c] : 0.f;
      mlp_vjp<D, H, L>(cw, ln, l, h[0], cot, first ? 1.f : 0.f, ub, acc);
      if (act) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          ybar[c] = first ? y0b[c] + ub[c] : y0b[c];
          fbar[c] = first ? 0.f : kb[0][c];
        }
      }
    }
    if (valid) {
      float g0[S::DL];
      load_frag<S::DL>(p.grad_traj + tj_off(p.layout, 0, b, p.B, p.T, D) + l * S::DL, g0);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) g0[c] += ybar[c];
      store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, g0);
    }
  }
  reduce_param_grads<D, H, L, WARPS>(acc, s_red, p.ws, ss, p.grad_params, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
}

// ---- host -----------------------------------------------------------------------------------------------------------------
int dopri5_traj_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                          float* traj, GodeStepLog* log, int32_t* n_acc, int32_t* n_att, double* att_dt, float* att_er,
                          uint8_t* att_acc, float* ckpt, double* acc_t0, double* acc_dt, cudaStream_t st) {
  if (T > kTrajMaxT) return GODE_ERR_T_TOO_LONG;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  Dp5TrajArgs a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.log = log; a.n_acc = n_acc; a.n_att = n_att;
  a.mailbox = status_mailbox();
  a.att_dt = att_dt; a.att_er = att_er; a.att_acc = att_acc; a.ckpt = ckpt; a.acc_t0 = acc_t0; a.acc_dt = acc_dt;
  a.o = *opts; a.B = B; a.T = T; a.layout = out_layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  cudaError_t e = cudaMemsetAsync(log, 0, sizeof(GodeStepLog), st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  constexpr int WARPS = 4, L = 8;
  const int per_cta = WARPS * Shape<16, 16, L>::G;
  int grid = (B + per_cta - 1) / per_cta;
  const int cap = sm_count() * 12;
  if (grid > cap) grid = cap;
  dopri5_traj_fwd_kernel<16, 16, L, WARPS><<<grid, WARPS * 32, 0, st>>>(a);
  return launch_status();
}

int dopri5_traj_small_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, int layout, const GodeStepLog* log,
                          const int32_t* n_acc, const float* ckpt, const double* acc_t0, const double* acc_dt,
                          int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                          size_t ws_bytes, cudaStream_t st) {
  if (T > kTrajMaxT) return GODE_ERR_T_TOO_LONG;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  constexpr int WARPS = 4, L = 8;
  using S = Shape<16, 16, L>;
  Dp5TrajArgs a{};
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.grad_traj = grad_traj; a.log = const_cast<GodeStepLog*>(log);
  a.n_acc = const_cast<int32_t*>(n_acc); a.ckpt = const_cast<float*>(ckpt);
  a.acc_t0 = const_cast<double*>(acc_t0); a.acc_dt = const_cast<double*>(acc_dt);
  a.o.ckpt_capacity = ckpt_capacity; a.o.fsign = fsign; a.grad_y0 = grad_y0; a.grad_params = grad_params;
  a.B = B; a.T = T; a.layout = layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  auto kern = dopri5_traj_bwd_kernel<16, 16, L, WARPS>;
  const size_t smem = sizeof(float) * (WARPS * BwdLines<16, 16, L>::kFloatsPerWarp + ColWeights<16, 16, L>::kFloats + WARPS * S::P);
  static int limit_cache = 0;
  int cap = coop_limit(kern, WARPS * 32, smem, limit_cache);
  if (cap <= 0) return GODE_ERR_COOP;
  if (cap > bwd_grid_cap()) cap = bwd_grid_cap();
  const int per_cta = WARPS * S::G;
  int grid = (B + per_cta - 1) / per_cta;
  if (grid > cap) grid = cap;
  if (ws_bytes < bwd_workspace_bytes(S::P)) return GODE_ERR_WORKSPACE;
  grid_sync_bind(a.ws.gs, workspace);
  a.ws.partials = reinterpret_cast<float*>(ws_scratch(workspace));
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(WARPS * 32), args, smem, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

}  // namespace gode
