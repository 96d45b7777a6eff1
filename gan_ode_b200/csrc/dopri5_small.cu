// dopri5_small.cu — adaptive Dormand–Prince 5(4) for the small-field shapes, FP32 state / FP64 time.
//   dopri5_fwd_kernel           torchdiffeq RKAdaptiveStepsizeODESolver, batch-global RMS error norm:
//                               ONE (t, dt) for the whole batch, so every attempted step ends in a grid-wide
//                               reduction.  The kernel is launched cooperatively (all CTAs co-resident), keeps every
//                               trajectory's y0/f0/k1..k7 in registers for the whole solve and replaces torchdiffeq's
//                               per-attempt device->host sync (`if error_ratio <= 1`) by an in-kernel barrier.
//   dopri5_backprop_bwd_kernel  reverse-mode through the accepted steps + dense-output interpolation,
//                               dt sequence read from the device-side step log (no host round trip).
#include <stdlib.h>
#include <cuda_runtime.h>
#ifdef GODE_TRACE
// Developer build only (python -m gan_ode_b200.build --trace -> libgode_trace.so): thread 0 of CTA 0 stamps
// (%globaltimer ns, clock64) at fixed points of the two headline kernels; scripts/headline_trace.py reads them back.
__device__ unsigned long long g_gode_trace[2 * 64 * 2];
#define GODE_TP(kern, slot)                                                                   \
  if (blockIdx.x == 0 && threadIdx.x == 0) {                                                  \
    unsigned long long gt_;                                                                   \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                   \
    g_gode_trace[((kern) * 64 + (slot)) * 2] = gt_;                                           \
    g_gode_trace[((kern) * 64 + (slot)) * 2 + 1] = (unsigned long long)clock64();             \
  }
extern "C" int gode_debug_trace_read(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_gode_trace, sizeof(g_gode_trace));
}
#define GODE_TP_R(slot) GODE_TP(1, slot)   // trace points inside reduce_param_grads (small_field.cuh)
// latest time ANY CTA passes this point (the global timer only grows, so the maximum belongs to the last launch)
#define GODE_TP_LAST(slot)                                                                    \
  if (threadIdx.x == 0) {                                                                     \
    unsigned long long gt_;                                                                   \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                   \
    atomicMax(&g_gode_trace[(64 + (slot)) * 2], gt_);                                         \
  }
#else
#define GODE_TP(kern, slot)
#endif
#include "dopri5_common.cuh"
#include "gru_cell.cuh"


namespace gode {

constexpr int kMaxT = GODE_ADAPTIVE_MAX_T;  // output times passed by value
constexpr int kDp5StageT = 32;  // output grids up to this length have their upstream gradients staged in shared memory

struct Dp5Args {
  const float *y0, *W1, *b1, *W2, *b2;
  const float* grad_traj;
  float* traj;
  float* grad_y0;
  float* grad_params;
  ReduceWs ws;
  GodeStepLog* log;
  int32_t* mailbox;           // mapped host int (gode_set_status_mailbox) or null: receives a non-zero status
  double* att_t0; double* att_dt; float* att_er; uint8_t* att_acc;
  float* ckpt; double* acc_t0; double* acc_dt;
  GridSyncWs gs;              // persistent grid all-reduce region (grid_sync.cuh)
  GodeAdaptiveOpts o;
  int B, T, layout;
  // world-scope norm (gode_dopri5_fwd_world): off when w_world <= 1
  int w_rank, w_world;
  long long w_total_B;
  unsigned long long* const* w_slots;
  unsigned int* w_launch_ctr;
  // persistent ODE-RNN sampler (RNN instantiation, models/mocogan_ode_rnn.py:45-52): F (solve over [0,1] -> GRU jump) pairs in
  // ONE cooperative launch.  Frame f: traj = seg + f*2*B*D ([input copy, h']), log block at (char*)log + f*log_stride
  // ([GodeStepLog | attempt arrays] as gode_dopri5_fwd lays them out), ckpt + f*kc*B*D, acc_t0/acc_dt + f*2*kc
  int F;
  size_t log_stride;
  const float* eps;                           // (F,B,D) the per-frame noise inputs e_t
  const float *g_wih, *g_whh, *g_bih, *g_bhh; // nn.GRUCell parameters
  float* codes;                               // (F,B,D) h_f
  double t[kMaxT];
};

__device__ __forceinline__ size_t toff(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

// ---- world-scope norm: exchange of the per-rank totals over NVLink peer memory, inside the solver kernel ----------------
// Every thread of the grid enters with the same per-rank totals v[] (grid_allreduce_sum) and leaves with the totals over
// all ranks.  CTA 0 stores one tagged word {fp32 value | tag} per value into slot [rank] of EVERY rank's exchange buffer
// (st.relaxed.sys: an aligned 8-byte store is single-copy atomic, value and tag cannot be seen torn, so no fence); warp 0
// of every CTA polls its OWN GPU's buffer — remote writes land there — until the `world` words carry this epoch's tag, and
// adds them in rank order.  The tag is the epoch number, counted ACROSS launches (the counter lives in device memory; every
// rank runs the same sequence of world-scope solves with the same number of reductions, because they share the step
// control), so tags stay unique under CUDA-graph replay and the parity double-buffering argument carries over launch
// boundaries: a peer publishes epoch e+2 only after it finished e+1, which needs this rank's e+1 word, which is stored only
// after every CTA here has passed its epoch-e poll.  A peer that never shows up trips a 10 s timeout instead of hanging.
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int NV>
__device__ __forceinline__ void world_allreduce_sum(double (&v)[NV], const Dp5Args& p, unsigned int& wepoch, int& status,
                                                    int lane, int warp) {
  __shared__ double s_w[kGsMaxVals];
  __shared__ int s_late;
  ++wepoch;
  const unsigned int tag = wepoch;
  const int par = (int)(wepoch & 1u), W = p.w_world;
  if (blockIdx.x == 0 && warp == 0) {
    for (int idx = lane; idx < W * NV; idx += 32) {
      const int r = idx / NV, k = idx % NV;
      float val = 0.f;  // (no dynamic index into v: that would put the caller's array into local memory)
#pragma unroll
      for (int kk = 0; kk < NV; ++kk) val = k == kk ? (float)v[kk] : val;
      st_relaxed_sys_u64(p.w_slots[r] + (size_t)(par * kGsMaxVals + k) * W + p.w_rank,
                         (unsigned long long)__float_as_uint(val) | ((unsigned long long)tag << 32));
    }
  }
  if (warp == 0) {
    const unsigned long long* mine = p.w_slots[p.w_rank] + (size_t)par * kGsMaxVals * W;
    const unsigned long long t_start = global_ns();
    bool late = false;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float mine_v = 0.f;
      if (lane < W) {
        unsigned long long x = ld_relaxed_sys_u64(mine + (size_t)k * W + lane);
        while ((unsigned int)(x >> 32) != tag) {
          if (global_ns() - t_start > 10000000000ull) { late = true; break; }
          x = ld_relaxed_sys_u64(mine + (size_t)k * W + lane);
        }
        mine_v = __uint_as_float((unsigned int)x);
      }
      double s = 0.0;
      for (int r = 0; r < W; ++r) s += (double)__shfl_sync(0xffffffffu, mine_v, r);
      if (lane == 0) s_w[k] = s;
    }
    const bool any_late = __any_sync(0xffffffffu, late);
    if (lane == 0) s_late = any_late ? 1 : 0;
  }
  __syncthreads();  // (the next write to s_w is separated from these reads by the barriers of the next grid reduction)
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_w[k];
  if (s_late) status |= GODE_ST_PEER_TIMEOUT;
}

// WORLD: world-scope norm compiled in (gode_dopri5_fwd_world).  A template parameter, not a runtime flag: the mere presence
// of the exchange code cost the ordinary solve 4 us of 30 at the bench shape (measured), so the default instantiation is
// kept free of it.
template <int D, int H, int L, int WARPS, bool WORLD, int TAB = GODE_TAB_DOPRI5, bool RNN = false>
__global__ void __launch_bounds__(WARPS * 32) dopri5_fwd_kernel(const __grid_constant__ Dp5Args p) {
  using S = Shape<D, H, L>;
  using TB = Tableau<TAB>;
  constexpr int NS = TB::NS;
  __shared__ __align__(16) float s_lines[WARPS * FwdLines<D, H, L>::kFloatsPerWarp];
  __shared__ float s_f[WARPS * kGsMaxVals];
  __shared__ double s_d[kGsMaxVals];
  __shared__ float s_gru_raw[RNN ? sizeof(GruW) / sizeof(float) : 1];   // GRU weights (RNN instantiation only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  GODE_TP(0, 0);
  griddep_launch_dependents();   // a backward launched with PDL may stage its weights under this kernel's tail
  SyncState ss;
  ss.begin(p.gs);                // persistent tags: continue where the previous launch on this workspace stopped
  FwdLines<D, H, L> ln;
  ln.bind(s_lines + warp * FwdLines<D, H, L>::kFloatsPerWarp, g);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  GruW& s_gru = *reinterpret_cast<GruW*>(s_gru_raw);
  if constexpr (RNN) {
    load_w(s_gru, p.g_wih, p.g_whh, p.g_bih, p.g_bhh, tid, WARPS * 32);
    __syncthreads();
  }
  const int b = (blockIdx.x * WARPS + warp) * S::G + g;
  const bool valid = b < p.B;
  const bool logger = (blockIdx.x == 0 && tid == 0);
  constexpr bool world = WORLD;
  const double n_elem = (double)(world ? p.w_total_B : (long long)p.B) * (double)D;
  const double inv_n = 1.0 / n_elem;   // the norms below are rms values rounded to fp32 (torch computes them in fp32): one
                                       // fp64 multiply + fp32 sqrt instead of an fp64 divide + fp64 sqrt on the critical path
  const float rtol32 = (float)p.o.rtol, atol32 = (float)p.o.atol;
  // cumulative world epoch: read before the first grid-wide reduction; CTA 0 writes it back after the last one
  unsigned int wepoch = 0;
  if constexpr (WORLD) wepoch = *reinterpret_cast<volatile unsigned int*>(p.w_launch_ctr);

  float y0[S::DL], k[NS + 1][S::DL], hk[S::HL];
#pragma unroll
  for (int c = 0; c < S::DL; ++c) y0[c] = 0.f;
  if (valid) load_frag<S::DL>(p.y0 + (size_t)b * D + l * S::DL, y0);
  int status = 0;
  const int n_frames = RNN ? p.F : 1;
  for (int fr = 0; fr < n_frames; ++fr) {   // one pass for odeint; the ODE-RNN sampler runs its frames here
  // this frame's outputs (the plain solve: the caller's buffers)
  float* traj = p.traj;
  GodeStepLog* log = p.log;
  double* att_t0 = p.att_t0; double* att_dt = p.att_dt; float* att_er = p.att_er; uint8_t* att_acc = p.att_acc;
  float* ckpt = p.ckpt; double* acc_t0 = p.acc_t0; double* acc_dt = p.acc_dt;
  if constexpr (RNN) {
    const size_t cap = (size_t)p.o.log_capacity, kc = (size_t)p.o.ckpt_capacity;
    unsigned char* lg = reinterpret_cast<unsigned char*>(p.log) + (size_t)fr * p.log_stride;
    traj = p.traj + (size_t)fr * 2 * p.B * D;
    log = reinterpret_cast<GodeStepLog*>(lg);
    att_t0 = reinterpret_cast<double*>(lg + 64); att_dt = reinterpret_cast<double*>(lg + 64 + 8 * cap);
    att_er = reinterpret_cast<float*>(lg + 64 + 16 * cap); att_acc = lg + 64 + 20 * cap;
    ckpt = p.ckpt ? p.ckpt + (size_t)fr * kc * p.B * D : nullptr;
    acc_t0 = p.acc_t0 ? p.acc_t0 + (size_t)fr * 2 * kc : nullptr;
    acc_dt = p.acc_t0 ? p.acc_t0 + (size_t)fr * 2 * kc + kc : nullptr;
  }
  float hp[S::DL];   // RNN: the solve's last output h' (what the jump consumes)
#pragma unroll
  for (int c = 0; c < S::DL; ++c) hp[c] = 0.f;
  if (valid) store_frag<S::DL>(traj + toff(p.layout, 0, b, p.B, p.T, D) + l * S::DL, y0);
  GODE_TP(0, 1);
  field<D, H, L>(w, ln, l, p.o.fsign, y0, k[0], hk);  // f0
  GODE_TP(0, 2);
  int nfe = 1;
  double t0 = p.t[0];
  double dt;

  // ---- misc.py::_select_initial_step (order = 4) -------------------------------------------------------------
  {
    float scale[S::DL];
    double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < S::DL; ++c) {
      scale[c] = atol32 + fabsf(y0[c]) * rtol32;
      const float r0 = y0[c] / scale[c], r1 = k[0][c] / scale[c];
      if (valid) {
        v[0] += (double)r0 * (double)r0;
        v[1] += (double)r1 * (double)r1;
        if (!isfinite(y0[c])) v[2] += 1.0;
      }
    }
    grid_allreduce_sum<3, WARPS>(v, s_f, s_d, p.gs, ss, lane, warp);
    if constexpr (WORLD) world_allreduce_sum<3>(v, p, wepoch, status, lane, warp);
    if (v[2] > 0.0) status |= GODE_ST_NONFINITE;
    if (p.o.first_step > 0.0) {
      dt = p.o.first_step;
    } else {
      const float d0 = sqrtf((float)(v[0] * inv_n)), d1 = sqrtf((float)(v[1] * inv_n));
      const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
      float u[S::DL], f1[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) u[c] = y0[c] + h0 * k[0][c];
      field<D, H, L>(w, ln, l, p.o.fsign, u, f1, hk);
      nfe++;
      double v2[1] = {0.0};
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        const float r = (f1[c] - k[0][c]) / scale[c];
        if (valid) v2[0] += (double)r * (double)r;
      }
      grid_allreduce_sum<1, WARPS>(v2, s_f, s_d, p.gs, ss, lane, warp);
      if constexpr (WORLD) world_allreduce_sum<1>(v2, p, wepoch, status, lane, warp);
      const float d2 = sqrtf((float)(v2[0] * inv_n)) / h0;
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
      else h1 = powf(0.01f / fmaxf(d1, d2), 1.f / (float)TB::ORDER);   // _select_initial_step(order = ORDER - 1)
      dt = (double)fminf(100.f * h0, h1);
    }
  }
  if (logger) log->dt0 = dt;
  GODE_TP(0, 3);

  // ---- solvers.py::AdaptiveStepsizeODESolver.integrate / rk_common.py::_adaptive_step ---------------------------
  int iout = 1, n_att = 0, n_acc = 0, n_steps = 0;
  while (iout < p.T && status == 0) {
    if (n_steps >= p.o.max_num_steps) { status |= GODE_ST_MAX_STEPS; break; }
    if (!(t0 + dt > t0)) { status |= GODE_ST_DT_UNDERFLOW; break; }
    const double t1 = t0 + dt;
    const float dt32 = (float)dt;
    float u[S::DL];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        float s = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
        for (int j = 1; j <= i; ++j) s = fmaf(k[j][c], TB::beta(i, j) * dt32, s);
        u[c] = y0[c] + s;
      }
      field<D, H, L>(w, ln, l, p.o.fsign, u, k[i + 1], hk);
    }
    nfe += NS;
    GODE_TP(0, 4 + 3 * min(n_att, 8));
    // FSAL (c_sol == beta[NS-1], dopri5 / bosh3): u is y1.  Otherwise y1 comes from c_sol.  Either way k[NS] is f1.
    if constexpr (!TB::FSAL) {
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        float s = k[0][c] * (dt32 * TB::csol(0));
#pragma unroll
        for (int j = 1; j <= NS; ++j) s = fmaf(k[j][c], dt32 * TB::csol(j), s);
        u[c] = y0[c] + s;
      }
    }
    double v[1] = {0.0};
#pragma unroll
    for (int c = 0; c < S::DL; ++c) {
      float e = k[0][c] * (dt32 * TB::cerr(0));
#pragma unroll
      for (int j = 1; j <= NS; ++j)
        if (TB::cerr(j) != 0.f) e = fmaf(k[j][c], dt32 * TB::cerr(j), e);
      const float tol = atol32 + rtol32 * fmaxf(fabsf(y0[c]), fabsf(u[c]));
      const float r = e / tol;
      if (valid) v[0] += (double)r * (double)r;
    }
    grid_allreduce_sum<1, WARPS>(v, s_f, s_d, p.gs, ss, lane, warp);
    if constexpr (WORLD) world_allreduce_sum<1>(v, p, wepoch, status, lane, warp);
    GODE_TP(0, 5 + 3 * min(n_att, 8));
    const float er = sqrtf((float)(v[0] * inv_n));
    bool accept = er <= 1.f;
    if (dt > p.o.max_step) accept = false;
    if (dt <= p.o.min_step) accept = true;
    if (logger && n_att < p.o.log_capacity) {
      att_t0[n_att] = t0; att_dt[n_att] = dt; att_er[n_att] = er; att_acc[n_att] = accept ? 1 : 0;
    }
    if (accept) {
      if (p.o.ckpt_capacity > 0) {
        if (n_acc < p.o.ckpt_capacity) {
          if (valid) store_frag<S::DL>(ckpt + ((size_t)n_acc * p.B + b) * D + l * S::DL, y0);
          // not FSAL: the step's f0 is the PREVIOUS step's last stage derivative, not f(y0) — the replay cannot recompute
          // it from y0, so it is checkpointed as well (second half of the checkpoint buffer: 2 * ckpt_capacity rows)
          if constexpr (!TB::FSAL)
            if (valid) store_frag<S::DL>(ckpt + ((size_t)(p.o.ckpt_capacity + n_acc) * p.B + b) * D + l * S::DL, k[0]);
          if (logger) { acc_t0[n_acc] = t0; acc_dt[n_acc] = dt; }
        } else {
          status |= GODE_ST_CKPT_OVERFLOW;
        }
      }
      if (iout < p.T && p.t[iout] <= t1) {
        // interp.py::_interp_fit (dense output), only when an output lands in this step
        float ca[S::DL], cb[S::DL], cc[S::DL], cd[S::DL];
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
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
        // 1 / (t1 - t0) to fp64 accuracy without an fp64 division: fp32 reciprocal + two Newton steps (6e-8 -> 4e-15 -> 0)
        const double span = t1 - t0;
        double inv_span = (double)__frcp_rn((float)span);
        inv_span = inv_span * (2.0 - span * inv_span);
        inv_span = inv_span * (2.0 - span * inv_span);
        // Which outputs land in (t0, t1] and their normalised positions: lane q examines output iout + q, so the fp64
        // subtract / multiply / convert (slow, dependent) of all of them run side by side instead of once per loop trip;
        // the output times increase, so the hits are a prefix.  Identical in every warp of the grid.
        while (iout < p.T) {
          const int io = iout + lane;
          const bool in = io < p.T && p.t[io < p.T ? io : p.T - 1] <= t1;
          const float xq = in ? (float)((p.t[io < p.T ? io : p.T - 1] - t0) * inv_span) : 0.f;
          const int cnt = __popc(__ballot_sync(0xffffffffu, in));
          for (int q = 0; q < cnt; ++q) {
            const float x = __shfl_sync(0xffffffffu, xq, q);
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
            if (valid) store_frag<S::DL>(traj + toff(p.layout, iout + q, b, p.B, p.T, D) + l * S::DL, o);
            if constexpr (RNN) {
#pragma unroll
              for (int c = 0; c < S::DL; ++c) hp[c] = o[c];
            }
          }
          iout += cnt;
          if (cnt > 0) n_steps = -1;  // max_num_steps is counted per output interval (per _advance call)
          if (cnt < 32) break;
        }
      }
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { y0[c] = u[c]; k[0][c] = k[NS][c]; }
      t0 = t1;
      ++n_acc;
    }
    dt = optimal_step<TB::ORDER>(dt, er, p.o);
    dt = fmin(fmax(dt, p.o.min_step), p.o.max_step);
    GODE_TP(0, 6 + 3 * min(n_att, 8));
    ++n_att;
    ++n_steps;
  }
  if (logger) {
    log->status = status;
    if (status != 0 && p.mailbox) *reinterpret_cast<volatile int32_t*>(p.mailbox) = status;
    log->n_attempts = n_att;
    log->n_accepted = n_acc;
    log->nfe = nfe;
    log->t_final = t0;
  }
  if constexpr (RNN) {
    if (status != 0) {   // a failed frame ends the sampler: the remaining frames' logs carry the status too
      if (logger)
        for (int f2 = fr + 1; f2 < n_frames; ++f2)
          reinterpret_cast<GodeStepLog*>(reinterpret_cast<unsigned char*>(p.log) + (size_t)f2 * p.log_stride)->status = status;
      break;
    }
    // the jump h_f = GRUCell(e_f, h'_f) (models/mocogan_ode_rnn.py:49): per trajectory, no grid-wide step
    float xe[S::DL];
#pragma unroll
    for (int c = 0; c < S::DL; ++c) xe[c] = 0.f;
    if (valid) load_frag<S::DL>(p.eps + ((size_t)fr * p.B + b) * D + l * S::DL, xe);
    __syncwarp();
    store_frag<S::DL>(ln.y + l * S::DL, xe);
    store_frag<S::DL>(ln.h + l * S::DL, hp);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < S::DL; ++c) {
      const Gates gt = gates(s_gru, ln.y, ln.h, l * S::DL + c);
      y0[c] = (1.f - gt.z) * gt.n + gt.z * hp[c];
    }
    if (valid) store_frag<S::DL>(p.codes + ((size_t)fr * p.B + b) * D + l * S::DL, y0);
    __syncwarp();
  }
  }   // frames
  if (logger) {
    if constexpr (WORLD) *p.w_launch_ctr = wepoch;
    ss.finish(p.gs);   // every CTA has arrived at the last reduction, hence has read the bases
  }
  GODE_TP(0, 40);
}

// ------------------------------------------------------------------------------------------------------------
template <int D, int H, int L, int WARPS, int TAB = GODE_TAB_DOPRI5>
__global__ void __launch_bounds__(WARPS * 32) dopri5_backprop_bwd_kernel(const __grid_constant__ Dp5Args p) {
  using S = Shape<D, H, L>;
  using BL = BwdLines<D, H, L>;
  using TB = Tableau<TAB>;
  constexpr int NS = TB::NS;
  extern __shared__ __align__(16) float smem[];
  float* s_lines = smem;
  float* s_cw = s_lines + WARPS * BL::kFloatsPerWarp;
  float* s_red = s_cw + ColWeights<D, H, L>::kFloats;
  // upstream gradients of this warp's trajectories for ALL output times, staged once per trajectory group with every load
  // in flight (T <= kDp5StageT): inside the replay loop they used to be T dependent global loads, each a full L2/HBM latency
  float* s_gr = s_red + WARPS * S::P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  GODE_TP(1, 0);
  const bool staged = p.T <= kDp5StageT;
  float* my_gr = s_gr + (size_t)warp * p.T * (S::G * D);
  // (griddep_wait first: a no-op for an ordinary launch.  Overlapping this prologue with the forward's tail by a programmatic
  // dependent launch was built and measured, profiles/README.md round 2: it does not pay for this pair of kernels.)
  griddep_wait();
  // the step log and the sync bases are read first so that their latency overlaps the weight loads below (all cold reads)
  const int log_n_accepted = p.log->n_accepted, log_status = p.log->status;
  SyncState ss;
  ss.begin(p.ws.gs);
  const int n4 = p.T * (S::G * D / 4);   // float4 elements of one trajectory group's upstream gradients
  // float4 e: output time e / (G*D/4), trajectory, 4 components.  Asynchronous global -> shared copies (no registers, all in
  // flight at once); rows past the batch are zero-filled.  The first group's copies are issued before the weight loads.
  auto stage_grads = [&](int base) {
    for (int e = lane; e < n4; e += 32) {
      const int so = e / (S::G * D / 4), r = e % (S::G * D / 4), gg = r / (D / 4), q4 = r % (D / 4);
      const bool in = base + gg < p.B;
      const float* src = in ? p.grad_traj + toff(p.layout, so, base + gg, p.B, p.T, D) + 4 * q4 : p.grad_traj;
      cp_async16(my_gr + (size_t)so * (S::G * D) + gg * D + 4 * q4, src, in);
    }
  };
  const int base_first = (blockIdx.x * WARPS + warp) * S::G;
  GODE_TP(1, 2);
  if (staged && base_first < p.B) stage_grads(base_first);
  GODE_TP(1, 50);
  BL ln;
  ln.bind(s_lines + warp * BL::kFloatsPerWarp, g);
  ColWeights<D, H, L> cw;
  cw.bind(s_cw);
  cw.stage(p.W1, p.W2, tid, WARPS * 32);
  GODE_TP(1, 51);
  RowWeights<D, H, L> w;
  w.load(p.W1, p.b1, p.W2, p.b2, l);
  GradAcc<D, H, L> acc;
  GODE_TP(1, 1);
  __syncthreads();
  GODE_TP(1, 3);
  const int n_acc = min(log_n_accepted, p.o.ckpt_capacity);
  // A forward that failed (dt underflow, non-finite state, step budget, checkpoint overflow) has no valid replay:
  // return NaN gradients instead of silently truncated ones (the host does not sync to look at the status).
  const float poison = log_status != 0 ? __int_as_float(0x7fc00000) : 0.f;
  acc.fill(poison);
  GODE_TP(1, 52);
  const int stride = gridDim.x * WARPS * S::G;
  for (int base = (blockIdx.x * WARPS + warp) * S::G; base < p.B; base += stride) {
    const int b = base + g;
    const bool valid = b < p.B;
    const float sc = valid ? 1.f : 0.f;
    float ybar[S::DL], fbar[S::DL];  // cotangents of the next step's (y0, f0) == this step's (y1, f1)
#pragma unroll
    for (int c = 0; c < S::DL; ++c) { ybar[c] = poison; fbar[c] = 0.f; }
    int iout = p.T - 1;
    if (staged) {
      if (base != base_first) {
        __syncwarp();
        stage_grads(base);
      }
    }
    // checkpoint and step table of the step about to be replayed are fetched one step ahead; the first fetch is issued
    // before waiting for the staged gradients so that the two cold reads overlap
    float y0n[S::DL];
    double t0n = 0.0, dtn = 0.0;
#pragma unroll
    for (int c = 0; c < S::DL; ++c) y0n[c] = 0.f;
    if (n_acc > 0) {
      t0n = p.acc_t0[n_acc - 1]; dtn = p.acc_dt[n_acc - 1];
      if (valid) load_frag<S::DL>(p.ckpt + ((size_t)(n_acc - 1) * p.B + b) * D + l * S::DL, y0n);
    }
    GODE_TP(1, 53);
    if (staged) {
      cp_async_wait_all();
      __syncwarp();
    }
    GODE_TP(1, 4);
    for (int s = n_acc - 1; s >= 0; --s) {
      GODE_TP(1, 5 + 2 * min(s, 8));
      const double t0 = t0n, dtd = dtn, t1 = t0 + dtd;
      const float dt32 = (float)dtd;
      float y0[S::DL], k[NS + 1][S::DL], h[NS + 1][S::HL], u[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) y0[c] = y0n[c];
      if (s > 0) {
        t0n = p.acc_t0[s - 1]; dtn = p.acc_dt[s - 1];
        if (valid) load_frag<S::DL>(p.ckpt + ((size_t)(s - 1) * p.B + b) * D + l * S::DL, y0n);
      }
      // recompute the step exactly as the forward did
      if (TB::FSAL || s == 0) {
        field<D, H, L>(w, ln, l, p.o.fsign, y0, k[0], h[0]);
      } else {   // not FSAL: f0 of this step is the previous step's k[NS] (checkpointed by the forward)
#pragma unroll
        for (int c = 0; c < S::DL; ++c) k[0][c] = 0.f;
        if (valid) load_frag<S::DL>(p.ckpt + ((size_t)(p.o.ckpt_capacity + s) * p.B + b) * D + l * S::DL, k[0]);
#pragma unroll
        for (int c = 0; c < S::HL; ++c) h[0][c] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float sum = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], TB::beta(i, j) * dt32, sum);
          u[c] = y0[c] + sum;
        }
        field<D, H, L>(w, ln, l, p.o.fsign, u, k[i + 1], h[i + 1]);
      }
      GODE_TP(1, 6 + 2 * min(s, 8));
      // cotangents of the interpolation inputs (y0, y1, ymid, f0, f1) from every output inside (t0, t1]
      float y0b[S::DL], ymb[S::DL], f0b[S::DL], kb[NS + 1][S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) { y0b[c] = 0.f; ymb[c] = 0.f; f0b[c] = 0.f; }
      const double inv_span = 1.0 / (t1 - t0);
      // lane q examines output iout - q (see the forward: the fp64 position arithmetic of all outputs of the step in parallel)
      while (iout >= 1) {
        const int io = iout - lane;
        const bool in = io >= 1 && p.t[io >= 1 ? io : 1] > t0;
        const float xq = in ? (float)((p.t[io >= 1 ? io : 1] - t0) * inv_span) : 0.f;
        const int cnt = __popc(__ballot_sync(0xffffffffu, in));
        for (int q = 0; q < cnt; ++q) {
          const int ic = iout - q;
          float gout[S::DL];
#pragma unroll
          for (int c = 0; c < S::DL; ++c) gout[c] = 0.f;
          if (staged) load_frag<S::DL>(my_gr + (size_t)ic * (S::G * D) + g * D + l * S::DL, gout);
          else if (valid) load_frag<S::DL>(p.grad_traj + toff(p.layout, ic, b, p.B, p.T, D) + l * S::DL, gout);
          const float x = __shfl_sync(0xffffffffu, xq, q);
          const float p2 = x * x, p3 = p2 * x, p4 = p3 * x;
          const float cy0 = 1.f - 11.f * p2 + 18.f * p3 - 8.f * p4;
          const float cy1 = -5.f * p2 + 14.f * p3 - 8.f * p4;
          const float cym = 16.f * p2 - 32.f * p3 + 16.f * p4;
          const float cf0 = dt32 * (x - 4.f * p2 + 5.f * p3 - 2.f * p4);
          const float cf1 = dt32 * (p2 - 3.f * p3 + 2.f * p4);
#pragma unroll
          for (int c = 0; c < S::DL; ++c) {
            y0b[c] = fmaf(cy0, gout[c], y0b[c]);
            ybar[c] = fmaf(cy1, gout[c], ybar[c]);
            ymb[c] = fmaf(cym, gout[c], ymb[c]);
            f0b[c] = fmaf(cf0, gout[c], f0b[c]);
            fbar[c] = fmaf(cf1, gout[c], fbar[c]);
          }
        }
        iout -= cnt;
        if (cnt < 32) break;
      }
      // ymid = y0 + sum_j k_j dt cmid_j ; f1 = k[NS] ; f0 = k[0]
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        y0b[c] += ymb[c];
#pragma unroll
        for (int j = 0; j <= NS; ++j) kb[j][c] = (dt32 * TB::cmid(j)) * ymb[c];
        kb[NS][c] += fbar[c];
        kb[0][c] += f0b[c];
        if constexpr (!TB::FSAL) {   // y1 = y0 + dt sum_j c_sol[j] k[j]: its cotangent goes to y0 and to every k
          y0b[c] += ybar[c];
#pragma unroll
          for (int j = 0; j <= NS; ++j) kb[j][c] = fmaf(TB::csol(j) * dt32, ybar[c], kb[j][c]);
        }
      }
      // last stage: k[NS] = f(u_last) (the lines still hold u_last / h[NS] from the recompute); FSAL: u_last == y1
      // k_j = fsign * f(u_j): the cotangent reaching f is fsign * kb_j
      float ub[S::DL], cot[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) cot[c] = p.o.fsign * kb[NS][c];
      mlp_vjp<D, H, L>(cw, ln, l, h[NS], cot, sc, ub, acc);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        if constexpr (TB::FSAL) ub[c] += ybar[c];
        y0b[c] += ub[c];
#pragma unroll
        for (int j = 0; j < NS; ++j) kb[j][c] = fmaf(TB::beta(NS - 1, j) * dt32, ub[c], kb[j][c]);
      }
      // earlier stages, last to first
#pragma unroll
      for (int i = NS - 2; i >= 0; --i) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float sum = k[0][c] * (TB::beta(i, 0) * dt32);
#pragma unroll
          for (int j = 1; j <= i; ++j) sum = fmaf(k[j][c], TB::beta(i, j) * dt32, sum);
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
          for (int j = 0; j <= i; ++j) kb[j][c] = fmaf(TB::beta(i, j) * dt32, ub[c], kb[j][c]);
        }
      }
      if (s > 0) {
        // k[0] of this step IS f1 = k[NS] of the previous step (same autograd node) -> hand its cotangent back
#pragma unroll
        for (int c = 0; c < S::DL; ++c) { ybar[c] = y0b[c]; fbar[c] = kb[0][c]; }
      } else {
        regather<D, H, L>(ln, l, y0, h[0]);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) cot[c] = p.o.fsign * kb[0][c];
        mlp_vjp<D, H, L>(cw, ln, l, h[0], cot, sc, ub, acc);
#pragma unroll
        for (int c = 0; c < S::DL; ++c) ybar[c] = y0b[c] + ub[c];
      }
    }
    float g0[S::DL];
#pragma unroll
    for (int c = 0; c < S::DL; ++c) g0[c] = 0.f;
    if (valid) {
      if (staged) load_frag<S::DL>(my_gr + g * D + l * S::DL, g0);
      else load_frag<S::DL>(p.grad_traj + toff(p.layout, 0, b, p.B, p.T, D) + l * S::DL, g0);
#pragma unroll
      for (int c = 0; c < S::DL; ++c) g0[c] += ybar[c];
      store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, g0);
    }
  }
  GODE_TP(1, 30);
  reduce_param_grads<D, H, L, WARPS>(acc, s_red, p.ws, ss, p.grad_params, lane, warp, tid);
  if (blockIdx.x == 0 && tid == 0) ss.finish(p.ws.gs);
  GODE_TP(1, 31);
}

// ------------------------------------------------------------------------------------------------------------
constexpr int kDp5Warps = 4;

template <int D, int H, int L, int WARPS = kDp5Warps>
static int dp5_fwd_grid(int B) {
  const int per_cta = WARPS * Shape<D, H, L>::G;
  return (B + per_cta - 1) / per_cta;
}

size_t dopri5_small_workspace_bytes(int B, int D, int H) {
  (void)D; (void)H;
  (void)B;
  return (size_t)GODE_SYNC_REGION_BYTES;   // the persistent sync region only
}

template <int D, int H, int L, int WARPS, bool WORLD, int TAB = GODE_TAB_DOPRI5, bool RNN = false>
static int launch_dp5_fwd_k(Dp5Args& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  const int grid = dp5_fwd_grid<D, H, L, WARPS>(a.B);
  auto kern = dopri5_fwd_kernel<D, H, L, WARPS, WORLD, TAB, RNN>;
  static int limit_cache = 0;
  const int cap = coop_limit(kern, WARPS * 32, 0, limit_cache);
  if (cap <= 0 || grid > cap || grid > kSyncMaxGrid) return GODE_ERR_COOP;
  if (ws_bytes < grid_sync_bytes(grid)) return GODE_ERR_WORKSPACE;
  grid_sync_bind(a.gs, workspace);
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(WARPS * 32), args, 0, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

template <int D, int H, int L, int WARPS = kDp5Warps>
static int launch_dp5_fwd(Dp5Args& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (a.o.tableau == GODE_TAB_BOSH3) return launch_dp5_fwd_k<D, H, L, WARPS, false, GODE_TAB_BOSH3>(a, workspace, ws_bytes, st);
  if (a.o.tableau == GODE_TAB_ADAPTIVE_HEUN)
    return launch_dp5_fwd_k<D, H, L, WARPS, false, GODE_TAB_ADAPTIVE_HEUN>(a, workspace, ws_bytes, st);
  return a.w_world > 1 ? launch_dp5_fwd_k<D, H, L, WARPS, true>(a, workspace, ws_bytes, st)
                       : launch_dp5_fwd_k<D, H, L, WARPS, false>(a, workspace, ws_bytes, st);
}

template <int D, int H, int L, int WARPS = 4, int TAB = GODE_TAB_DOPRI5>
static int launch_dp5_bwd(Dp5Args& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  using S = Shape<D, H, L>;
  auto kern = dopri5_backprop_bwd_kernel<D, H, L, WARPS, TAB>;
  const size_t smem = sizeof(float) * (WARPS * BwdLines<D, H, L>::kFloatsPerWarp + ColWeights<D, H, L>::kFloats + WARPS * S::P +
                                       (a.T <= kDp5StageT ? (size_t)WARPS * a.T * S::G * D : 0));
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

int dopri5_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                     const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                     float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                     float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes, cudaStream_t st,
                     const GodeWorld* world) {
  if (T > kMaxT) return GODE_ERR_T_TOO_LONG;
  Dp5Args a{};
  if (opts->tableau < GODE_TAB_DOPRI5 || opts->tableau > GODE_TAB_ADAPTIVE_HEUN || (world && opts->tableau != GODE_TAB_DOPRI5))
    return GODE_ERR_ARG;
  if (world) {
    a.w_rank = world->rank; a.w_world = world->world; a.w_total_B = world->total_B;
    a.w_slots = reinterpret_cast<unsigned long long* const*>(world->slots_dev); a.w_launch_ctr = world->launch_ctr;
  }
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.log = log; a.mailbox = status_mailbox();
  a.att_t0 = att_t0; a.att_dt = att_dt; a.att_er = att_er; a.att_acc = att_acc;
  a.ckpt = ckpt; a.acc_t0 = acc_t0; a.acc_dt = acc_dt;
  a.o = *opts; a.B = B; a.T = T; a.layout = out_layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  if (D == 16 && H == 16) {
    // Every trajectory of the batch has to be resident at once (batch-global error norm), and every attempted step ends in
    // a grid-wide reduction whose cost grows with the number of CTAs.  Measured (scripts/dp5_lane_sweep.py, graph replay):
    // 8 lanes per trajectory (shortest dependent chain, 16 trajectories per CTA) wins below ~2048 trajectories; from there
    // 4 lanes (32 per CTA, half the CTAs) is faster — 24.6 vs 28.8 us at B = 4096, 29.7 vs 37.8 us at B = 7000 — and holds
    // 9472 trajectories.  Beyond that 2 and finally 1 lane per trajectory (the compiler spills part of the stage vectors)
    // hold 18 944 / 56 832.  Larger batches: shard them (multi-GPU) or use norm='trajectory'.
    const char* force = getenv("GODE_DP5_LANES");  // developer switch: '8' / '4' force that mapping first
    const bool first8 = force ? force[0] == '8' : B < 2048;
    // (fatter CTAs to shrink the grid reduction were measured and lost: 4-lane mapping at B = 4096, graph replay, 4 warps per
    // CTA 24.7 us, 8 warps 28.8 us, 16 warps 106 us)
    int rc = first8 ? launch_dp5_fwd<16, 16, 8>(a, workspace, ws_bytes, st) : GODE_ERR_COOP;
    if (rc != GODE_ERR_COOP) return rc;
    rc = launch_dp5_fwd<16, 16, 4>(a, workspace, ws_bytes, st);
    if (rc != GODE_ERR_COOP) return rc;
    rc = launch_dp5_fwd<16, 16, 2>(a, workspace, ws_bytes, st);
    if (rc != GODE_ERR_COOP) return rc;
    return launch_dp5_fwd<16, 16, 1>(a, workspace, ws_bytes, st);
  }
  return GODE_ERR_SHAPE;
}

// All F frames of the ODE-RNN sampler in ONE cooperative launch (models/mocogan_ode_rnn.py:45-52): per frame the same dopri5
// solve over [0, 1] as gode_dopri5_fwd (batch-global control, its own initial-step selection, step log and checkpoints),
// then the GRU jump per trajectory — no grid-wide step for the jump, the grid-sync tags just keep counting across frames.
int dopri5_small_odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2,
                            const float* b2, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                            int B, int D, int H, int F, const GodeAdaptiveOpts* opts, float* codes, float* seg,
                            unsigned char* logs, size_t log_stride, float* ckpt, double* acc, void* workspace,
                            size_t ws_bytes, cudaStream_t st) {
  if (!(D == 16 && H == 16) || opts->tableau != GODE_TAB_DOPRI5) return GODE_ERR_SHAPE;
  Dp5Args a{};
  a.y0 = h0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = seg; a.log = reinterpret_cast<GodeStepLog*>(logs);
  a.mailbox = status_mailbox();
  a.ckpt = ckpt; a.acc_t0 = acc; a.acc_dt = acc;
  a.o = *opts; a.B = B; a.T = 2; a.layout = GODE_LAYOUT_TBD;
  a.t[0] = 0.0; a.t[1] = 1.0;
  a.F = F; a.log_stride = log_stride; a.eps = eps; a.g_wih = w_ih; a.g_whh = w_hh; a.g_bih = b_ih; a.g_bhh = b_hh;
  a.codes = codes;
  const bool first8 = B < 2048;
  int rc = first8 ? launch_dp5_fwd_k<16, 16, 8, kDp5Warps, false, GODE_TAB_DOPRI5, true>(a, workspace, ws_bytes, st) : GODE_ERR_COOP;
  if (rc != GODE_ERR_COOP) return rc;
  rc = launch_dp5_fwd_k<16, 16, 4, kDp5Warps, false, GODE_TAB_DOPRI5, true>(a, workspace, ws_bytes, st);
  if (rc != GODE_ERR_COOP) return rc;
  rc = launch_dp5_fwd_k<16, 16, 2, kDp5Warps, false, GODE_TAB_DOPRI5, true>(a, workspace, ws_bytes, st);
  if (rc != GODE_ERR_COOP) return rc;
  return launch_dp5_fwd_k<16, 16, 1, kDp5Warps, false, GODE_TAB_DOPRI5, true>(a, workspace, ws_bytes, st);
}

int dopri5_small_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                              const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                              const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                              int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                              size_t ws_bytes, cudaStream_t st, const GodeWorld* xchg, int tableau) {
  if (T > kMaxT) return GODE_ERR_T_TOO_LONG;
  if (tableau != GODE_TAB_DOPRI5 && xchg) return GODE_ERR_ARG;
  Dp5Args a{};
  if (xchg) {
    a.ws.w_rank = xchg->rank; a.ws.w_world = xchg->world; a.ws.w_ctr = xchg->launch_ctr;
    a.ws.w_slots = reinterpret_cast<unsigned long long* const*>(xchg->slots_dev);
  }
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.grad_traj = grad_traj; a.log = const_cast<GodeStepLog*>(log);
  a.ckpt = const_cast<float*>(ckpt); a.acc_t0 = const_cast<double*>(acc_t0); a.acc_dt = const_cast<double*>(acc_dt);
  a.o.ckpt_capacity = ckpt_capacity; a.o.fsign = fsign; a.grad_y0 = grad_y0; a.grad_params = grad_params;
  a.B = B; a.T = T; a.layout = layout;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  if (D == 16 && H == 16 && tableau == GODE_TAB_BOSH3) return launch_dp5_bwd<16, 16, 8, 4, GODE_TAB_BOSH3>(a, workspace, ws_bytes, st);
  if (D == 16 && H == 16 && tableau == GODE_TAB_ADAPTIVE_HEUN)
    return launch_dp5_bwd<16, 16, 8, 4, GODE_TAB_ADAPTIVE_HEUN>(a, workspace, ws_bytes, st);
  if (D == 16 && H == 16) {
    // 4 trajectories per warp.  With 4-warp CTAs, 148 < CTAs <= 296 means 108 SMs carry two CTAs and 40 carry one: the kernel
    // is bound by shared-memory bandwidth, so the doubled SMs finish last and everyone waits for them in the final reduction
    // (trace in profiles/README.md).  In that range 7-warp CTAs put ONE CTA on every SM (B = 4096: 147 CTAs x 28 trajectories).
    const int sms = sm_count();
    const int grid4 = (B + 15) / 16, grid7 = (B + 27) / 28;
    const char* force = getenv("GODE_DP5_BWD_WARPS");
    const bool seven = force ? force[0] == '7' : (grid4 > sms && grid7 <= sms);
    if (seven) return launch_dp5_bwd<16, 16, 8, 7>(a, workspace, ws_bytes, st);
    return launch_dp5_bwd<16, 16, 8>(a, workspace, ws_bytes, st);  // (8-warp CTAs measured: slower)
  }
  return GODE_ERR_SHAPE;
}

}  // namespace gode
