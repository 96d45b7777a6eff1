// odernn_small.cu — the ODE-RNN sampler of models/mocogan_ode_rnn.py:40-54 behind ONE C-ABI call per direction.
//   per frame f:  h'_f = odeint(ode_fn, h_{f-1}, [0, 1])[-1]        (dopri5, the reference passes no solver kwargs)
//                 h_f  = GRUCell(e_f, h'_f)                          (input = noise, hidden = the ODE-evolved state)
// The solves are the fused dopri5 kernels of dopri5_small.cu (one launch per frame, batch-global step control exactly as
// each of the reference's 16 odeint calls); this file adds the fused GRU jump (forward and backward, PyTorch nn.GRUCell
// semantics: gate order r, z, n; n = tanh(W_in x + b_in + r (W_hn h + b_hn)); h_out = (1 - z) n + z h) and the frame
// loop on the host side of the C ABI, so a whole 16-frame forward (or backward) is enqueued without the Python
// interpreter, with every intermediate staying on the device:
//   * h_f is written straight into the output codes (F, B, D) and is the next frame's y0 (no copies);
//   * the backward walks the frames in reverse: GRU VJP -> discrete adjoint of that frame's solve
//     (dopri5_backprop_bwd_kernel over the frame's own device-side step log and checkpoints);
//   * parameter gradients: one slot per frame, summed in frame order at the end (deterministic).
#include <stdlib.h>
#include "launch.h"
#include "gru_cell.cuh"

namespace gode {

namespace {

// h_out = GRUCell(x, h); one thread per (trajectory, unit), 16 trajectories per 256-thread tile
__global__ void __launch_bounds__(256) gru_jump_fwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                           const float* w_ih, const float* w_hh, const float* b_ih,
                                                           const float* b_hh, int B, float* __restrict__ h_out) {
  __shared__ GruW w;
  __shared__ float sx[GT][GD + 1], sh[GT][GD + 1];
  const int tid = threadIdx.x, t = tid / GD, j = tid % GD;
  load_w(w, w_ih, w_hh, b_ih, b_hh, tid, 256);
  for (int base = blockIdx.x * GT; base < B; base += gridDim.x * GT) {
    const int b = base + t;
    __syncthreads();
    sx[t][j] = b < B ? x[(size_t)b * GD + j] : 0.f;
    sh[t][j] = b < B ? h[(size_t)b * GD + j] : 0.f;
    __syncthreads();
    const Gates g = gates(w, sx[t], sh[t], j);
    if (b < B) h_out[(size_t)b * GD + j] = (1.f - g.z) * g.n + g.z * sh[t][j];
  }
}

// VJP of the jump.  go = grad_out (+ carry), per tile: (t,j) threads form the gate cotangents, (j,k) threads accumulate
// the weight gradients over the tile's trajectories in registers (persistent over tiles), (t,k) threads form dx and dh.
// Each CTA writes ONE partial row of GP floats; gru_reduce_kernel adds the rows in CTA order.
__global__ void __launch_bounds__(256) gru_jump_bwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                           const float* w_ih, const float* w_hh, const float* b_ih,
                                                           const float* b_hh, const float* __restrict__ grad_out,
                                                           const float* __restrict__ grad_carry, int B,
                                                           float* __restrict__ grad_x, float* __restrict__ grad_h,
                                                           float* __restrict__ partial) {
  __shared__ GruW w;
  __shared__ float sx[GT][GD + 1], sh[GT][GD + 1];
  __shared__ float sa[GT][4][GD + 1];  // cotangents of (a_r, a_z, a_n, hn) per trajectory and unit
  __shared__ float sz[GT][GD + 1];     // go * z (direct path to h)
  const int tid = threadIdx.x, t = tid / GD, j = tid % GD;  // also used as (j, k) = (tid / GD, tid % GD)
  load_w(w, w_ih, w_hh, b_ih, b_hh, tid, 256);
  float aih[3] = {0.f, 0.f, 0.f}, ahh[3] = {0.f, 0.f, 0.f}, abi[3] = {0.f, 0.f, 0.f}, abh[3] = {0.f, 0.f, 0.f};
  for (int base = blockIdx.x * GT; base < B; base += gridDim.x * GT) {
    const int b = base + t;
    __syncthreads();
    sx[t][j] = b < B ? x[(size_t)b * GD + j] : 0.f;
    sh[t][j] = b < B ? h[(size_t)b * GD + j] : 0.f;
    __syncthreads();
    {
      const Gates g = gates(w, sx[t], sh[t], j);
      float go = 0.f;
      if (b < B) go = grad_out[(size_t)b * GD + j] + (grad_carry ? grad_carry[(size_t)b * GD + j] : 0.f);
      const float dn = go * (1.f - g.z), dz = go * (sh[t][j] - g.n);
      const float dan = dn * (1.f - g.n * g.n);
      const float dr = dan * g.hn;
      sa[t][0][j] = dr * g.r * (1.f - g.r);
      sa[t][1][j] = dz * g.z * (1.f - g.z);
      sa[t][2][j] = dan;
      sa[t][3][j] = dan * g.r;
      sz[t][j] = go * g.z;
    }
    __syncthreads();
    {  // weight gradients: this thread owns (unit jj = t, input k = j) of all three gates of both matrices
      const int jj = t, k = j;
#pragma unroll
      for (int tt = 0; tt < GT; ++tt) {
        const float xv = sx[tt][k], hv = sh[tt][k];
        const float ar = sa[tt][0][jj], az = sa[tt][1][jj], an = sa[tt][2][jj], ah = sa[tt][3][jj];
        aih[0] = fmaf(ar, xv, aih[0]); aih[1] = fmaf(az, xv, aih[1]); aih[2] = fmaf(an, xv, aih[2]);
        ahh[0] = fmaf(ar, hv, ahh[0]); ahh[1] = fmaf(az, hv, ahh[1]); ahh[2] = fmaf(ah, hv, ahh[2]);
        if (k == 0) { abi[0] += ar; abi[1] += az; abi[2] += an; abh[0] += ar; abh[1] += az; abh[2] += ah; }
      }
    }
    {  // dx[t][k], dh[t][k]
      const int k = j;
      float dx = 0.f, dh = sz[t][k];
#pragma unroll
      for (int jj = 0; jj < GD; ++jj) {
        const float ar = sa[t][0][jj], az = sa[t][1][jj], an = sa[t][2][jj], ah = sa[t][3][jj];
        dx = fmaf(w.wih[jj * GS + k], ar, dx); dx = fmaf(w.wih[(GD + jj) * GS + k], az, dx); dx = fmaf(w.wih[(2 * GD + jj) * GS + k], an, dx);
        dh = fmaf(w.whh[jj * GS + k], ar, dh); dh = fmaf(w.whh[(GD + jj) * GS + k], az, dh); dh = fmaf(w.whh[(2 * GD + jj) * GS + k], ah, dh);
      }
      if (b < B) {
        if (grad_x) grad_x[(size_t)b * GD + k] = dx;
        grad_h[(size_t)b * GD + k] = dh;
      }
    }
  }
  float* out = partial + (size_t)blockIdx.x * GP;
  const int jj = t, k = j;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    out[(g * GD + jj) * GD + k] = aih[g];
    out[3 * GD * GD + (g * GD + jj) * GD + k] = ahh[g];
    if (k == 0) { out[6 * GD * GD + g * GD + jj] = abi[g]; out[6 * GD * GD + 3 * GD + g * GD + jj] = abh[g]; }
  }
}

// out[e] (+)= sum over rows k of partial[k][e], rows added in order
__global__ void rows_sum_kernel(const float* __restrict__ partial, int rows, int len, float* __restrict__ out, int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  float s = accumulate ? out[e] : 0.f;
  for (int k = 0; k < rows; ++k) s += partial[(size_t)k * len + e];
  out[e] = s;
}

int gru_grid(int B) {
  const int tiles = (B + GT - 1) / GT, cap = sm_count() * 4;
  return tiles < cap ? tiles : cap;
}

}  // namespace

size_t odernn_log_stride(int log_capacity) { return align256(64 + (size_t)21 * (size_t)(log_capacity > 0 ? log_capacity : 0)); }

// workspace head shared by the per-frame solver launches: forward, replay backward, continuous adjoint
static size_t odernn_solver_ws(int B, int D, int H) {
  size_t m = dopri5_small_workspace_bytes(B, D, H);
  const size_t bw = bwd_workspace_bytes(H * D + H + D * H + D), ad = dopri5_small_adjoint_workspace_bytes(B, D, H);
  if (bw > m) m = bw;
  if (ad > m) m = ad;
  return align256(m);
}

size_t odernn_workspace_bytes(int B, int D, int H) {
  return odernn_solver_ws(B, D, H) + sizeof(float) * (size_t)GP * (size_t)(sm_count() * 4) + 256;
}

int gru_jump_fwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, int B, int D, float* h_out, cudaStream_t st) {
  if (D != GD) return GODE_ERR_SHAPE;
  gru_jump_fwd_kernel<<<gru_grid(B), 256, 0, st>>>(x, h, w_ih, w_hh, b_ih, b_hh, B, h_out);
  return launch_status();
}

// grad_params: flat [w_ih|w_hh|b_ih|b_hh]; accumulate != 0 adds to it.  workspace: gru partial rows.
int gru_jump_bwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, const float* grad_out, const float* grad_carry, int B, int D, float* grad_x,
                 float* grad_h, float* grad_params, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (D != GD) return GODE_ERR_SHAPE;
  const int grid = gru_grid(B);
  if (ws_bytes < sizeof(float) * (size_t)GP * (size_t)grid) return GODE_ERR_WORKSPACE;
  float* partial = reinterpret_cast<float*>(workspace);
  gru_jump_bwd_kernel<<<grid, 256, 0, st>>>(x, h, w_ih, w_hh, b_ih, b_hh, grad_out, grad_carry, B, grad_x, grad_h, partial);
  rows_sum_kernel<<<(GP + 255) / 256, 256, 0, st>>>(partial, grid, GP, grad_params, accumulate);
  return launch_status();
}

int odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H, int F,
               const GodeAdaptiveOpts* opts, float* codes, float* seg, unsigned char* logs, float* ckpt, double* acc,
               int32_t* n_acc, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (D != GD || !small_field_shape(D, H)) return GODE_ERR_SHAPE;
  const double t01[2] = {0.0, 1.0};
  const int cap = opts->log_capacity, kc = opts->ckpt_capacity;
  const size_t ls = odernn_log_stride(cap), bd = (size_t)B * D;
  const bool per_traj = opts->norm_scope == GODE_NORM_TRAJ;
  GodeAdaptiveOpts ot = *opts;
  ot.log_capacity = 0;  // per-trajectory mode keeps no per-attempt logs here
  // batch-global control (torchdiffeq's): ONE persistent cooperative kernel runs all F (solve, jump) pairs
  // (dopri5_small.cu::dopri5_fwd_kernel<..., RNN>).  GODE_ODERNN_PERFRAME=1 (developer switch) keeps the round-1 path:
  // one solver launch + one jump launch per frame.
  {
    const char* pf = getenv("GODE_ODERNN_PERFRAME");
    if (!per_traj && !(pf && pf[0] == '1'))
      return dopri5_small_odernn_fwd(h0, eps, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh, B, D, H, F, opts, codes, seg, logs, ls,
                                     kc > 0 ? ckpt : nullptr, kc > 0 ? acc : nullptr, workspace, ws_bytes, st);
  }
  for (int f = 0; f < F; ++f) {
    const float* y0 = f == 0 ? h0 : codes + (size_t)(f - 1) * bd;
    unsigned char* lg = logs + (size_t)f * ls;
    float* tr = seg + (size_t)f * 2 * bd;
    if (per_traj) {  // every trajectory its own (t, dt): no grid-wide reduction, ordinary launch (dopri5_traj_small.cu)
      double* a0 = kc > 0 ? acc + (size_t)f * 2 * kc * B : nullptr;
      int rc = dopri5_traj_small_fwd(y0, W1, b1, W2, b2, t01, B, D, H, 2, &ot, GODE_LAYOUT_TBD, tr,
                                     reinterpret_cast<GodeStepLog*>(lg), n_acc + (size_t)f * B, n_acc + (size_t)F * B, nullptr,
                                     nullptr, nullptr, kc > 0 ? ckpt + (size_t)f * kc * bd : nullptr, a0,
                                     kc > 0 ? a0 + (size_t)kc * B : nullptr, st);
      if (rc) return rc;
      rc = gru_jump_fwd(eps + (size_t)f * bd, tr + bd, w_ih, w_hh, b_ih, b_hh, B, D, codes + (size_t)f * bd, st);
      if (rc) return rc;
      continue;
    }
    int rc = dopri5_small_fwd(y0, W1, b1, W2, b2, t01, B, D, H, 2, opts, GODE_LAYOUT_TBD, tr,
                              reinterpret_cast<GodeStepLog*>(lg), reinterpret_cast<double*>(lg + 64),
                              reinterpret_cast<double*>(lg + 64 + 8 * (size_t)cap), reinterpret_cast<float*>(lg + 64 + 16 * (size_t)cap),
                              lg + 64 + 20 * (size_t)cap, kc > 0 ? ckpt + (size_t)f * kc * bd : nullptr,
                              kc > 0 ? acc + (size_t)f * 2 * kc : nullptr, kc > 0 ? acc + (size_t)f * 2 * kc + kc : nullptr,
                              workspace, ws_bytes, st);
    if (rc) return rc;
    rc = gru_jump_fwd(eps + (size_t)f * bd, tr + bd, w_ih, w_hh, b_ih, b_hh, B, D, codes + (size_t)f * bd, st);
    if (rc) return rc;
  }
  return GODE_OK;
}

// grad_codes (F,B,D) -> grad_h0 (B,D), grad_eps (F,B,D) or null, grad_ode (P1) and grad_gru (GP) OVERWRITTEN.
// scratch: floats [carry (B,D) | gtraj (2,B,D) | ode slots (F, P1)]
int odernn_bwd(const float* grad_codes, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H, int F,
               int ckpt_capacity, const float* seg, const unsigned char* logs, size_t log_stride, const float* ckpt,
               const double* acc, const int32_t* n_acc, const GodeAdaptiveOpts* adjoint_opts, int adjoint_param_mask,
               float* grad_h0, float* grad_eps, float* grad_ode, float* grad_gru, float* scratch, void* workspace,
               size_t ws_bytes, cudaStream_t st) {
  if (D != GD || !small_field_shape(D, H)) return GODE_ERR_SHAPE;
  const double t01[2] = {0.0, 1.0};
  const int P1 = H * D + H + D * H + D, kc = ckpt_capacity;
  const size_t bd = (size_t)B * D;
  float* carry = scratch;
  float* gtraj = scratch + bd;
  float* slots = scratch + 3 * bd;
  const size_t dp_ws = odernn_solver_ws(B, D, H);
  if (ws_bytes < dp_ws + sizeof(float) * (size_t)GP * (size_t)gru_grid(B)) return GODE_ERR_WORKSPACE;
  void* gru_ws = reinterpret_cast<unsigned char*>(workspace) + dp_ws;
  cudaMemsetAsync(gtraj, 0, sizeof(float) * bd, st);  // no gradient reaches the solve's copy of its own input
  for (int f = F - 1; f >= 0; --f) {
    const float* tr = seg + (size_t)f * 2 * bd;
    int rc = gru_jump_bwd(eps + (size_t)f * bd, tr + bd, w_ih, w_hh, b_ih, b_hh, grad_codes + (size_t)f * bd,
                          f == F - 1 ? nullptr : carry, B, D, grad_eps ? grad_eps + (size_t)f * bd : nullptr, gtraj + bd,
                          grad_gru, f != F - 1, gru_ws, ws_bytes - dp_ws, st);
    if (rc) return rc;
    const unsigned char* lg = logs + (size_t)f * log_stride;
    if (adjoint_opts) {  // torchdiffeq's continuous adjoint of this frame's solve (dopri5_adj_small.cu); needs no checkpoints
      rc = dopri5_small_adjoint_bwd(tr, gtraj, W1, b1, W2, b2, t01, B, D, H, 2, GODE_LAYOUT_TBD, adjoint_opts, adjoint_param_mask,
                                    f == 0 ? grad_h0 : carry, slots + (size_t)f * P1, nullptr, nullptr, nullptr, nullptr,
                                    workspace, dp_ws, st);
      if (rc) return rc;
      continue;
    }
    if (n_acc) {  // per-trajectory step control: replay every trajectory's own accepted steps
      const double* a0 = acc + (size_t)f * 2 * kc * B;
      rc = dopri5_traj_small_bwd(gtraj, W1, b1, W2, b2, t01, B, D, H, 2, GODE_LAYOUT_TBD,
                                 reinterpret_cast<const GodeStepLog*>(lg), n_acc + (size_t)f * B, ckpt + (size_t)f * kc * bd, a0,
                                 a0 + (size_t)kc * B, kc, 1.0f, f == 0 ? grad_h0 : carry, slots + (size_t)f * P1, workspace,
                                 dp_ws, st);
      if (rc) return rc;
      continue;
    }
    rc = dopri5_small_backprop_bwd(gtraj, W1, b1, W2, b2, t01, B, D, H, 2, GODE_LAYOUT_TBD,
                                   reinterpret_cast<const GodeStepLog*>(lg), ckpt + (size_t)f * kc * bd,
                                   acc + (size_t)f * 2 * kc, acc + (size_t)f * 2 * kc + kc, kc, 1.0f,
                                   f == 0 ? grad_h0 : carry, slots + (size_t)f * P1, workspace, dp_ws, st);
    if (rc) return rc;
  }
  rows_sum_kernel<<<(P1 + 255) / 256, 256, 0, st>>>(slots, F, P1, grad_ode, 0);
  return launch_status();
}

}  // namespace gode
