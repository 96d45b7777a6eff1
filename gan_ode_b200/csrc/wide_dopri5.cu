// wide_dopri5.cu — torchdiffeq's dopri5 for WIDE fields (D, H multiples of 32; e.g. D=64 / H=256, BASELINE.json configs[3]'s
// "larger motion latent"), FP32, with the batch-global step controller of the reference solver
// (torchdiffeq/_impl/rk_common.py::_adaptive_step, misc.py::_select_initial_step, interp.py::_interp_fit) and the gradient of
// the recorded accepted steps (backprop-through-solver, what `odeint` + autograd gives).
//
// The lane-split kernels of dopri5_small.cu keep every trajectory in registers for the whole solve; at D=64 / H=256 the
// weights alone are 136 KB of shared memory per CTA, so one CTA per SM with one WARP per trajectory holds 1184 trajectories at
// a time.  Instead of limiting the batch to that:
//   * ONE persistent cooperative kernel, grid = one CTA per SM; warp w owns trajectories w, w + NW, w + 2 NW, ... for the
//     whole solve (nothing crosses warps except the error norm);
//   * between attempts a trajectory lives in global memory as (y, f(y)) — 2 D floats, ping-pong buffers [current | candidate];
//     an attempt reads (y0, f0), runs the six stages in registers, writes the candidate (y1, f1); accepting flips the buffers.
//     B = 4096 at D = 64 is 4 MB of state: it stays in the 126 MB L2;
//   * what an accepted step would store is written SPECULATIVELY during the attempt, while its inputs are still in registers:
//     the checkpoint y0 into slot n_accepted, and every output that lands in (t0, t0 + dt] through the dense-output polynomial.
//     A rejected attempt leaves garbage only in places the next accepted step overwrites (outputs >= iout, checkpoint slot
//     n_accepted), so there is no second pass over the batch after the accept decision;
//   * one grid all-reduce per attempt (grid_sync.cuh, persistent tags), fp64 controller identical in every thread.
// Backward: ordinary launch, one warp per trajectory walks the recorded steps in reverse (recompute six stages, seven VJPs)
// and emits its RK-weighted cotangent / activation rows to scratch, laid out [step][trajectory][stage] so that the valid rows
// are a prefix whose length the contraction kernels read from the device log (n_accepted is not known on the host without a
// sync); the same deterministic split-K contraction as wide_rk4.cu turns the rows into dW1, db1, dW2, db2.
#include "dopri5_common.cuh"
#include "grid_sync.cuh"
#include "wide_field.cuh"

namespace gode {

constexpr int kWideMaxT = GODE_ADAPTIVE_MAX_T;   // output times passed by value

struct WideDp5Args {
  const float *y0, *W1, *b1, *W2, *b2;
  const float* grad_traj;
  float* traj;
  float* grad_y0;
  GodeStepLog* log;
  int32_t* mailbox;
  double* att_t0; double* att_dt; float* att_er; uint8_t* att_acc;
  float* ckpt; double* acc_t0; double* acc_dt;
  GridSyncWs gs;
  float* state;               // [2][B][2*D] ping-pong (y, f)
  float *sa, *su, *sh, *sd;   // backward rows: (N, D), (N, D), (N, H), (N, H), N = ckpt_capacity * B * 7
  GodeAdaptiveOpts o;
  int B, T, layout;
  double t[kWideMaxT];
};

// ------------------------------------------------------------------------------------------------------------------------
template <int D, int H>
__global__ void __launch_bounds__(kWideWarps * 32) wide_dopri5_fwd_kernel(const __grid_constant__ WideDp5Args p) {
  using W = Wide<D, H>;
  using TB = Tableau<GODE_TAB_DOPRI5>;
  constexpr int NS = TB::NS, DL = W::DL;
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_f[kWideWarps * kGsMaxVals];
  __shared__ double s_d[kGsMaxVals];
  const int tid = threadIdx.x, l = tid & 31, warp = tid >> 5;
  SyncState ss;
  ss.begin(p.gs);
  W::stage(smem, p.W1, p.b1, p.W2, p.b2, tid, kWideWarps * 32);
  W w;
  w.bind(smem, warp);
  __syncthreads();
  const int gw = blockIdx.x * kWideWarps + warp, nw = gridDim.x * kWideWarps;
  const bool logger = (blockIdx.x == 0 && tid == 0);
  const double inv_n = 1.0 / ((double)p.B * (double)D);
  const float rtol32 = (float)p.o.rtol, atol32 = (float)p.o.atol, fsign = p.o.fsign;
  float* st[2] = {p.state, p.state + (size_t)p.B * 2 * D};
  int cur = 0, status = 0, nfe = 1;
  double t0 = p.t[0], dt;

  // ---- f0, the first output, misc.py::_select_initial_step (order = 4) ---------------------------------------------------
  {
    double v[3] = {0.0, 0.0, 0.0};
    for (int b = gw; b < p.B; b += nw) {
      float y[DL], f0[DL], hk[W::HL];
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        y[c] = p.y0[(size_t)b * D + l + 32 * c];
        p.traj[w_off(p.layout, 0, b, p.B, p.T, D) + l + 32 * c] = y[c];
      }
      w.forward(l, y, f0, hk);
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        f0[c] *= fsign;
        st[0][(size_t)b * 2 * D + l + 32 * c] = y[c];
        st[0][(size_t)b * 2 * D + D + l + 32 * c] = f0[c];
        const float scale = atol32 + fabsf(y[c]) * rtol32;
        const float r0 = y[c] / scale, r1 = f0[c] / scale;
        v[0] += (double)r0 * (double)r0;
        v[1] += (double)r1 * (double)r1;
        if (!isfinite(y[c])) v[2] += 1.0;
      }
    }
    grid_allreduce_sum<3, kWideWarps>(v, s_f, s_d, p.gs, ss, l, warp);
    if (v[2] > 0.0) status |= GODE_ST_NONFINITE;
    if (p.o.first_step > 0.0) {
      dt = p.o.first_step;
    } else {
      const float d0 = sqrtf((float)(v[0] * inv_n)), d1 = sqrtf((float)(v[1] * inv_n));
      const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
      double v2[1] = {0.0};
      for (int b = gw; b < p.B; b += nw) {
        float y[DL], f0[DL], u[DL], f1[DL], hk[W::HL];
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          y[c] = st[0][(size_t)b * 2 * D + l + 32 * c];
          f0[c] = st[0][(size_t)b * 2 * D + D + l + 32 * c];
          u[c] = y[c] + h0 * f0[c];
        }
        w.forward(l, u, f1, hk);
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          const float scale = atol32 + fabsf(y[c]) * rtol32;
          const float r = (f1[c] * fsign - f0[c]) / scale;
          v2[0] += (double)r * (double)r;
        }
      }
      nfe++;
      grid_allreduce_sum<1, kWideWarps>(v2, s_f, s_d, p.gs, ss, l, warp);
      const float d2 = sqrtf((float)(v2[0] * inv_n)) / h0;
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
      else h1 = powf(0.01f / fmaxf(d1, d2), 1.f / (float)TB::ORDER);
      dt = (double)fminf(100.f * h0, h1);
    }
  }
  if (logger) p.log->dt0 = dt;

  // ---- solvers.py::AdaptiveStepsizeODESolver.integrate / rk_common.py::_adaptive_step ---------------------------------------
  int iout = 1, n_att = 0, n_acc = 0, n_steps = 0;
  while (iout < p.T && status == 0) {
    if (n_steps >= p.o.max_num_steps) { status |= GODE_ST_MAX_STEPS; break; }
    if (!(t0 + dt > t0)) { status |= GODE_ST_DT_UNDERFLOW; break; }
    const double t1 = t0 + dt;
    const float dt32 = (float)dt;
    const bool any_out = p.t[iout] <= t1;
    const bool keep = p.o.ckpt_capacity > 0 && n_acc < p.o.ckpt_capacity;
    const double span = t1 - t0;
    double inv_span = (double)__frcp_rn((float)span);
    inv_span = inv_span * (2.0 - span * inv_span);
    inv_span = inv_span * (2.0 - span * inv_span);
    // outputs inside this (candidate) step and their normalised positions, once per attempt: lane q examines output iout + q
    // (the fp64 arithmetic of all of them side by side; the output times increase, so the hits are a prefix).  More than 32
    // hits in one step: the rest is evaluated per trajectory below.
    const int ioq = iout + l < p.T ? iout + l : p.T - 1;
    const bool inq = iout + l < p.T && p.t[ioq] <= t1;
    const float xq = inq ? (float)((p.t[ioq] - t0) * inv_span) : 0.f;
    const int n_out = __popc(__ballot_sync(0xffffffffu, inq));
    const float* sc = st[cur];
    float* sn = st[cur ^ 1];
    double v[1] = {0.0};
    for (int b = gw; b < p.B; b += nw) {
      float y0[DL], k[NS + 1][DL], u[DL], hk[W::HL];
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        y0[c] = sc[(size_t)b * 2 * D + l + 32 * c];
        k[0][c] = sc[(size_t)b * 2 * D + D + l + 32 * c];
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          float s = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) s = fmaf(k[j][c], TB::beta(i, j) * dt32, s);
          u[c] = y0[c] + s;
        }
        w.forward(l, u, k[i + 1], hk);
#pragma unroll
        for (int c = 0; c < DL; ++c) k[i + 1][c] *= fsign;
      }
      // FSAL: u is y1, k[NS] is f1
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        float e = k[0][c] * (dt32 * TB::cerr(0));
#pragma unroll
        for (int j = 1; j <= NS; ++j)
          if (TB::cerr(j) != 0.f) e = fmaf(k[j][c], dt32 * TB::cerr(j), e);
        const float tol = atol32 + rtol32 * fmaxf(fabsf(y0[c]), fabsf(u[c]));
        const float r = e / tol;
        v[0] += (double)r * (double)r;
        sn[(size_t)b * 2 * D + l + 32 * c] = u[c];
        sn[(size_t)b * 2 * D + D + l + 32 * c] = k[NS][c];
        if (keep) p.ckpt[((size_t)n_acc * p.B + b) * D + l + 32 * c] = y0[c];
      }
      if (any_out) {   // interp.py::_interp_fit + _interp_evaluate for every output inside this (candidate) step
        float ca[DL], cb[DL], cc[DL], cd[DL];
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          float m = k[0][c] * (dt32 * TB::cmid(0));
#pragma unroll
          for (int j = 1; j <= NS; ++j)
            if (TB::cmid(j) != 0.f) m = fmaf(k[j][c], dt32 * TB::cmid(j), m);
          const float ymid = y0[c] + m, f0 = k[0][c], f1 = k[NS][c], y1 = u[c];
          ca[c] = 2.f * dt32 * (f1 - f0) - 8.f * (y1 + y0[c]) + 16.f * ymid;
          cb[c] = dt32 * (5.f * f0 - 3.f * f1) + 18.f * y0[c] + 14.f * y1 - 32.f * ymid;
          cc[c] = dt32 * (f1 - 4.f * f0) - 11.f * y0[c] - 5.f * y1 + 16.f * ymid;
          cd[c] = dt32 * f0;
        }
        for (int io = iout; io < p.T && (io - iout < n_out || (n_out == 32 && p.t[io] <= t1)); ++io) {
          const float x = io - iout < 32 ? __shfl_sync(0xffffffffu, xq, io - iout) : (float)((p.t[io] - t0) * inv_span);
#pragma unroll
          for (int c = 0; c < DL; ++c) {
            float tot = y0[c] + x * cd[c];
            float xp = x * x;
            tot = tot + xp * cc[c];
            xp = xp * x;
            tot = tot + xp * cb[c];
            xp = xp * x;
            tot = tot + xp * ca[c];
            p.traj[w_off(p.layout, io, b, p.B, p.T, D) + l + 32 * c] = tot;
          }
        }
      }
    }
    nfe += NS;
    grid_allreduce_sum<1, kWideWarps>(v, s_f, s_d, p.gs, ss, l, warp);
    const float er = sqrtf((float)(v[0] * inv_n));
    bool accept = er <= 1.f;
    if (dt > p.o.max_step) accept = false;
    if (dt <= p.o.min_step) accept = true;
    if (logger && n_att < p.o.log_capacity) {
      p.att_t0[n_att] = t0; p.att_dt[n_att] = dt; p.att_er[n_att] = er; p.att_acc[n_att] = accept ? 1 : 0;
    }
    if (accept) {
      if (p.o.ckpt_capacity > 0) {
        if (keep) {
          if (logger) { p.acc_t0[n_acc] = t0; p.acc_dt[n_acc] = dt; }
        } else {
          status |= GODE_ST_CKPT_OVERFLOW;
        }
      }
      while (iout < p.T && p.t[iout] <= t1) {
        ++iout;
        n_steps = -1;  // max_num_steps is counted per output interval
      }
      cur ^= 1;
      t0 = t1;
      ++n_acc;
    }
    dt = optimal_step<TB::ORDER>(dt, er, p.o);
    dt = fmin(fmax(dt, p.o.min_step), p.o.max_step);
    ++n_att;
    ++n_steps;
  }
  if (logger) {
    p.log->status = status;
    if (status != 0 && p.mailbox) *reinterpret_cast<volatile int32_t*>(p.mailbox) = status;
    p.log->n_attempts = n_att;
    p.log->n_accepted = n_acc;
    p.log->nfe = nfe;
    p.log->t_final = t0;
    ss.finish(p.gs);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// reverse-mode through the recorded steps (same algebra as dopri5_backprop_bwd_kernel); parameter part as scratch rows
template <int D, int H>
__global__ void __launch_bounds__(kWideWarps * 32) wide_dopri5_backprop_kernel(const __grid_constant__ WideDp5Args p) {
  using W = Wide<D, H>;
  using TB = Tableau<GODE_TAB_DOPRI5>;
  constexpr int NS = TB::NS, DL = W::DL, HL = W::HL;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, l = tid & 31, warp = tid >> 5;
  W::stage(smem, p.W1, p.b1, p.W2, p.b2, tid, kWideWarps * 32);
  W w;
  w.bind(smem, warp);
  __syncthreads();
  const int n_acc = min(p.log->n_accepted, p.o.ckpt_capacity);
  // a failed forward has no valid replay: NaN gradients instead of silently truncated ones
  const float poison = p.log->status != 0 ? __int_as_float(0x7fc00000) : 0.f;
  const float fsign = p.o.fsign;
  for (int b = blockIdx.x * kWideWarps + warp; b < p.B; b += gridDim.x * kWideWarps) {
    float ybar[DL], fbar[DL];   // cotangents of the next step's (y0, f0) == this step's (y1, f1)
#pragma unroll
    for (int c = 0; c < DL; ++c) { ybar[c] = poison; fbar[c] = 0.f; }
    int iout = p.T - 1;
    for (int s = n_acc - 1; s >= 0; --s) {
      const double t0 = p.acc_t0[s], dtd = p.acc_dt[s], t1 = t0 + dtd;
      const float dt32 = (float)dtd;
      float y0[DL], k[NS + 1][DL], h[NS + 1][HL], u[DL];
#pragma unroll
      for (int c = 0; c < DL; ++c) y0[c] = p.ckpt[((size_t)s * p.B + b) * D + l + 32 * c];
      // recompute the step exactly as the forward did
      w.forward(l, y0, k[0], h[0]);
#pragma unroll
      for (int c = 0; c < DL; ++c) k[0][c] *= fsign;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          float sum = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], TB::beta(i, j) * dt32, sum);
          u[c] = y0[c] + sum;
        }
        w.forward(l, u, k[i + 1], h[i + 1]);
#pragma unroll
        for (int c = 0; c < DL; ++c) k[i + 1][c] *= fsign;
      }
      // cotangents of the interpolation inputs (y0, y1, ymid, f0, f1) from every output inside (t0, t1]
      float y0b[DL], ymb[DL], f0b[DL], kb[NS + 1][DL];
#pragma unroll
      for (int c = 0; c < DL; ++c) { y0b[c] = 0.f; ymb[c] = 0.f; f0b[c] = 0.f; }
      const double inv_span = 1.0 / (t1 - t0);
      while (iout >= 1 && p.t[iout] > t0) {
        const float x = (float)((p.t[iout] - t0) * inv_span);
        const float p2 = x * x, p3 = p2 * x, p4 = p3 * x;
        const float cy0 = 1.f - 11.f * p2 + 18.f * p3 - 8.f * p4;
        const float cy1 = -5.f * p2 + 14.f * p3 - 8.f * p4;
        const float cym = 16.f * p2 - 32.f * p3 + 16.f * p4;
        const float cf0 = dt32 * (x - 4.f * p2 + 5.f * p3 - 2.f * p4);
        const float cf1 = dt32 * (p2 - 3.f * p3 + 2.f * p4);
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          const float gout = p.grad_traj[w_off(p.layout, iout, b, p.B, p.T, D) + l + 32 * c];
          y0b[c] = fmaf(cy0, gout, y0b[c]);
          ybar[c] = fmaf(cy1, gout, ybar[c]);
          ymb[c] = fmaf(cym, gout, ymb[c]);
          f0b[c] = fmaf(cf0, gout, f0b[c]);
          fbar[c] = fmaf(cf1, gout, fbar[c]);
        }
        --iout;
      }
      // ymid = y0 + sum_j k_j dt cmid_j ; f1 = k[NS] ; f0 = k[0]
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        y0b[c] += ymb[c];
#pragma unroll
        for (int j = 0; j <= NS; ++j) kb[j][c] = (dt32 * TB::cmid(j)) * ymb[c];
        kb[NS][c] += fbar[c];
        kb[0][c] += f0b[c];
      }
      const size_t row0 = ((size_t)s * p.B + b) * (NS + 1);
      float ub[DL], cot[DL], dl_[HL];
      // stage `stage` evaluated f at `uin` with tanh vector hk; cot = cotangent reaching f's output
      auto back = [&](int stage, const float (&uin)[DL], const float (&hk)[HL]) {
        w.vjp(l, hk, cot, ub, dl_);
        const size_t r = row0 + stage;
#pragma unroll
        for (int c = 0; c < DL; ++c) { p.sa[r * D + l + 32 * c] = cot[c]; p.su[r * D + l + 32 * c] = uin[c]; }
#pragma unroll
        for (int c = 0; c < HL; ++c) { p.sh[r * H + l + 32 * c] = hk[c]; p.sd[r * H + l + 32 * c] = dl_[c]; }
        __syncwarp();
      };
      // last stage: k[NS] = f(u_last), u_last == y1 (FSAL) is still in u
#pragma unroll
      for (int c = 0; c < DL; ++c) cot[c] = fsign * kb[NS][c];
      back(NS, u, h[NS]);
#pragma unroll
      for (int c = 0; c < DL; ++c) {
        ub[c] += ybar[c];
        y0b[c] += ub[c];
#pragma unroll
        for (int j = 0; j < NS; ++j) kb[j][c] = fmaf(TB::beta(NS - 1, j) * dt32, ub[c], kb[j][c]);
      }
#pragma unroll
      for (int i = NS - 2; i >= 0; --i) {
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          float sum = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], TB::beta(i, j) * dt32, sum);
          u[c] = y0[c] + sum;
          cot[c] = fsign * kb[i + 1][c];
        }
        back(i + 1, u, h[i + 1]);
#pragma unroll
        for (int c = 0; c < DL; ++c) {
          y0b[c] += ub[c];
#pragma unroll
          for (int j = 0; j <= i; ++j) kb[j][c] = fmaf(TB::beta(i, j) * dt32, ub[c], kb[j][c]);
        }
      }
      if (s > 0) {
        // k[0] of this step IS f1 = k[NS] of the previous step (same autograd node): hand its cotangent back; this step has
        // no stage-0 evaluation of its own, its row stays out of the sums (zero cotangent)
#pragma unroll
        for (int c = 0; c < DL; ++c) { ybar[c] = y0b[c]; fbar[c] = kb[0][c]; cot[c] = 0.f; }
        const size_t r = row0;
#pragma unroll
        for (int c = 0; c < DL; ++c) { p.sa[r * D + l + 32 * c] = 0.f; p.su[r * D + l + 32 * c] = 0.f; }
#pragma unroll
        for (int c = 0; c < HL; ++c) { p.sh[r * H + l + 32 * c] = 0.f; p.sd[r * H + l + 32 * c] = 0.f; }
      } else {
#pragma unroll
        for (int c = 0; c < DL; ++c) cot[c] = fsign * kb[0][c];
        back(0, y0, h[0]);
#pragma unroll
        for (int c = 0; c < DL; ++c) ybar[c] = y0b[c] + ub[c];
      }
    }
#pragma unroll
    for (int c = 0; c < DL; ++c)
      p.grad_y0[(size_t)b * D + l + 32 * c] = p.grad_traj[w_off(p.layout, 0, b, p.B, p.T, D) + l + 32 * c] + ybar[c];
  }
}

// ---- contraction of the valid rows: C[M][N] = sum_r A[r][M] * Bm[r][N], column sums of A; row count from the device log --------
// A FIXED number of slices (kWideSlices CTAs) shares whatever rows the solve recorded: slice s takes rows [s*per, (s+1)*per),
// per = ceil(rows / slices) rounded up to 16 — the host cannot size the grid by n_accepted without a sync, and sizing it by the
// checkpoint capacity left a handful of CTAs with all the work (3 recorded steps of 64: 25 busy CTAs of 518).  The second pass
// adds the slices in order: deterministic for a given (batch, n_accepted).
constexpr int kWideSlices = 592;

__device__ __forceinline__ long long wide_dp5_rows_per_slice(long long rows, int slices) {
  const long long per = (rows + slices - 1) / slices;
  return ((per + 15) / 16) * 16;
}

template <int M, int N>
__global__ void __launch_bounds__(256) wide_dp5_wgrad_partial_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                     const GodeStepLog* __restrict__ log, int ckpt_capacity,
                                                                     long long rows_per_step, float* __restrict__ partial) {
  constexpr int OUT = M * N / 256;
  static_assert((M * N) % 256 == 0, "tile must divide over 256 threads");
  __shared__ __align__(16) float sA[16][M];
  __shared__ __align__(16) float sB[16][N];
  const int t = threadIdx.x;
  const long long rows = (long long)min(log->n_accepted, ckpt_capacity) * rows_per_step;
  const long long per = wide_dp5_rows_per_slice(rows, gridDim.x);
  const long long r0 = (long long)blockIdx.x * per;
  if (r0 >= rows) return;   // beyond the recorded steps: the reduction below does not read this slice
  const long long r1 = r0 + per < rows ? r0 + per : rows;
  float acc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) acc[o] = 0.f;
  float colsum = 0.f;
  for (long long r = r0; r < r1; r += 16) {
    const int nr = (int)(r1 - r < 16 ? r1 - r : 16);
    for (int e = t; e < 16 * M; e += 256) sA[e / M][e % M] = (e / M) < nr ? A[(r + e / M) * M + e % M] : 0.f;
    for (int e = t; e < 16 * N; e += 256) sB[e / N][e % N] = (e / N) < nr ? Bm[(r + e / N) * N + e % N] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (t < M) colsum += sA[k][t];
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const int idx = o * 256 + t;
        acc[o] = fmaf(sA[k][idx / N], sB[k][idx % N], acc[o]);
      }
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.x * (M * N + M);
#pragma unroll
  for (int o = 0; o < OUT; ++o) out[o * 256 + t] = acc[o];
  if (t < M) out[M * N + t] = colsum;
}

__global__ void wide_dp5_wgrad_reduce_kernel(const float* __restrict__ partial, const GodeStepLog* __restrict__ log,
                                             int ckpt_capacity, long long rows_per_step, int n_slices, int len,
                                             float* __restrict__ out_w, int len_w, float* __restrict__ out_b) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  const long long rows = (long long)min(log->n_accepted, ckpt_capacity) * rows_per_step;
  const long long per = wide_dp5_rows_per_slice(rows, n_slices);
  const int slices = rows > 0 ? (int)((rows + per - 1) / per) : 0;
  float s = log->status != 0 ? __int_as_float(0x7fc00000) : 0.f;
  for (int k = 0; k < slices; ++k) s += partial[(size_t)k * len + e];
  if (e < len_w) out_w[e] = s;
  else out_b[e - len_w] = s;
}

// ---- host ---------------------------------------------------------------------------------------------------------------
static size_t wide_dp5_rows(int B, int kc) { return (size_t)B * (size_t)kc * 7; }

size_t wide_dopri5_workspace_bytes(int B, int D, int H, int ckpt_capacity, int backward) {
  if (!backward) return (size_t)GODE_SYNC_REGION_BYTES + sizeof(float) * (size_t)B * 4 * D + 256;
  const size_t rows = wide_dp5_rows(B, ckpt_capacity);
  const size_t slices = kWideSlices;
  return (size_t)GODE_SYNC_REGION_BYTES + sizeof(float) * (rows * (size_t)(2 * D + 2 * H) +
                                                          slices * (size_t)(H * D + (H > D ? H : D))) + 1024;
}

template <int D, int H>
static int launch_wide_dp5_fwd(WideDp5Args& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  auto kern = wide_dopri5_fwd_kernel<D, H>;
  const size_t smem = Wide<D, H>::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(1000 + (int)e);
  static int limit_cache = 0;
  const int cap = coop_limit(kern, kWideWarps * 32, smem, limit_cache);
  if (cap <= 0) return GODE_ERR_COOP;
  int grid = (a.B + kWideWarps - 1) / kWideWarps;
  if (grid > cap) grid = cap;
  if (grid > kSyncMaxGrid) grid = kSyncMaxGrid;
  if (ws_bytes < wide_dopri5_workspace_bytes(a.B, D, H, 0, 0)) return GODE_ERR_WORKSPACE;
  grid_sync_bind(a.gs, workspace);
  a.state = reinterpret_cast<float*>(ws_scratch(workspace));
  void* args[] = {(void*)&a};
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kWideWarps * 32), args, smem, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

template <int D, int H>
static int launch_wide_dp5_bwd(WideDp5Args& a, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  const int kc = a.o.ckpt_capacity;
  if (ws_bytes < wide_dopri5_workspace_bytes(a.B, D, H, kc, 1)) return GODE_ERR_WORKSPACE;
  const size_t rows = wide_dp5_rows(a.B, kc);
  float* base = reinterpret_cast<float*>(ws_scratch(workspace));
  a.sa = base; a.su = a.sa + rows * D; a.sh = a.su + rows * D; a.sd = a.sh + rows * H;
  float* partial = a.sd + rows * H;
  auto kern = wide_dopri5_backprop_kernel<D, H>;
  const size_t smem = Wide<D, H>::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(1000 + (int)e);
  int grid = (a.B + kWideWarps - 1) / kWideWarps;
  const int cap = sm_count() * (smem > 110 * 1024 ? 1 : 2);
  if (grid > cap) grid = cap;
  kern<<<grid, kWideWarps * 32, smem, st>>>(a);
  if (int rc = launch_status()) return rc;
  const int slices = kWideSlices;
  const long long rps = (long long)a.B * 7;
  // dW1 (H x D) = delta^T u ; db1 = colsum(delta)     flat layout [W1 | b1 | W2 | b2]
  wide_dp5_wgrad_partial_kernel<H, D><<<slices, 256, 0, st>>>(a.sd, a.su, a.log, kc, rps, partial);
  wide_dp5_wgrad_reduce_kernel<<<(H * D + H + 255) / 256, 256, 0, st>>>(partial, a.log, kc, rps, slices, H * D + H, grad_params,
                                                                       H * D, grad_params + H * D);
  // dW2 (D x H) = cot^T h ; db2 = colsum(cot)
  wide_dp5_wgrad_partial_kernel<D, H><<<slices, 256, 0, st>>>(a.sa, a.sh, a.log, kc, rps, partial);
  wide_dp5_wgrad_reduce_kernel<<<(D * H + D + 255) / 256, 256, 0, st>>>(partial, a.log, kc, rps, slices, D * H + D,
                                                                       grad_params + H * D + H, D * H,
                                                                       grad_params + H * D + H + D * H);
  return launch_status();
}

int wide_dopri5_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                    const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                    float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                    float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (T > kWideMaxT) return GODE_ERR_T_TOO_LONG;
  if (opts->tableau != GODE_TAB_DOPRI5) return GODE_ERR_SHAPE;   // bosh3 / adaptive_heun exist for the reference shape only
  WideDp5Args a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.log = log; a.mailbox = status_mailbox();
  a.att_t0 = att_t0; a.att_dt = att_dt; a.att_er = att_er; a.att_acc = att_acc;
  a.ckpt = ckpt; a.acc_t0 = acc_t0; a.acc_dt = acc_dt;
  a.o = *opts; a.B = B; a.T = T; a.layout = out_layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  if (D == 64 && H == 256) return launch_wide_dp5_fwd<64, 256>(a, workspace, ws_bytes, st);
  if (D == 32 && H == 32) return launch_wide_dp5_fwd<32, 32>(a, workspace, ws_bytes, st);
  if (D == 32 && H == 64) return launch_wide_dp5_fwd<32, 64>(a, workspace, ws_bytes, st);
  return GODE_ERR_SHAPE;
}

int wide_dopri5_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2, const float* b2,
                             const double* t_host, int B, int D, int H, int T, int layout, const GodeStepLog* log,
                             const float* ckpt, const double* acc_t0, const double* acc_dt, int ckpt_capacity, float fsign,
                             float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (T > kWideMaxT) return GODE_ERR_T_TOO_LONG;
  WideDp5Args a{};
  a.grad_traj = grad_traj; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.log = const_cast<GodeStepLog*>(log);
  a.ckpt = const_cast<float*>(ckpt); a.acc_t0 = const_cast<double*>(acc_t0); a.acc_dt = const_cast<double*>(acc_dt);
  a.grad_y0 = grad_y0;
  a.o.ckpt_capacity = ckpt_capacity; a.o.fsign = fsign;
  a.B = B; a.T = T; a.layout = layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  if (D == 64 && H == 256) return launch_wide_dp5_bwd<64, 256>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 32) return launch_wide_dp5_bwd<32, 32>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 64) return launch_wide_dp5_bwd<32, 64>(a, grad_params, workspace, ws_bytes, st);
  return GODE_ERR_SHAPE;
}

}  // namespace gode
