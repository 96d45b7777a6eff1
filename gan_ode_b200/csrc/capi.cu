// capi.cu — the extern "C" boundary declared in include/gode.h.  Argument checking and dispatch only.
#include "launch.h"

using namespace gode;

namespace gode {
static int32_t* g_status_mailbox = nullptr;   // set once by the host before the first solve, read-only afterwards
int32_t* status_mailbox() { return g_status_mailbox; }
int& thread_launch_flags() {
  static thread_local int flags = 0;
  return flags;
}
}  // namespace gode

extern "C" {

const char* gode_version(void) { return "gode 0.2 (sm_100a)"; }

int gode_workspace_init(void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (!workspace || ws_bytes < (size_t)GODE_SYNC_REGION_BYTES) return GODE_ERR_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(workspace, 0, GODE_SYNC_REGION_BYTES, (cudaStream_t)stream);
  return e == cudaSuccess ? GODE_OK : -(1000 + (int)e);
}

int gode_stream_capture_id(gode_stream_t stream, unsigned long long* id_out) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  cudaError_t e = cudaStreamGetCaptureInfo((cudaStream_t)stream, &st, &id);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return -(1000 + (int)e); }
  if (id_out) *id_out = id;
  return st == cudaStreamCaptureStatusActive ? 1 : 0;
}

int gode_set_status_mailbox(int32_t* host_mapped) {
  gode::g_status_mailbox = host_mapped;
  return GODE_OK;
}

int gode_set_thread_launch_flags(int flags) {
  int& f = gode::thread_launch_flags();
  const int prev = f;
  f = flags;
  return prev;
}

const char* gode_strerror(int code) {
  switch (code) {
    case GODE_OK: return "ok";
    case GODE_ERR_SHAPE: return "no kernel compiled for this (D, H) at this precision";
    case GODE_ERR_ARG: return "invalid argument (null pointer, B <= 0, T < 2, bad layout)";
    case GODE_ERR_T_TOO_LONG:
      return "a time / step table exceeds its launch-parameter limit (include/gode.h: GODE_MAX_HOST_STEPS for a host dt table "
             "of the fixed-grid solvers — pass dt on the device instead; GODE_ADAPTIVE_MAX_T output times; GODE_SDE_MAX_* "
             "steps / frames / cells / reverse steps)";
    case GODE_ERR_WORKSPACE: return "workspace too small";
    case GODE_ERR_COOP: return "batch-global adaptive solve needs every CTA co-resident: batch too large";
    case GODE_ERR_PRECISION: return "precision mode not available for this entry point";
    default: return code <= -1000 ? cudaGetErrorString((cudaError_t)(-code - 1000)) : "unknown gode error";
  }
}

int gode_param_count(int D, int H) { return H * D + H + D * H + D; }

int gode_supported(int D, int H, int precision) {
  if (precision == GODE_PREC_FP32) return (small_field_shape(D, H) || wide_shape(D, H)) ? 1 : 0;
  if (precision == GODE_PREC_TF32) return tc_shape(D, H) ? 1 : 0;
  if (precision == GODE_PREC_BF16) return (tc_shape(D, H) || tc_wide_shape(D, H)) ? 1 : 0;
  return 0;
}

static bool bad_common(const void* a, const void* b, const void* c, const void* d, const void* e, int B, int T, int layout) {
  return !a || !b || !c || !d || !e || B <= 0 || T < 2 || (layout != GODE_LAYOUT_TBD && layout != GODE_LAYOUT_BTD);
}

int gode_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                 int dt_on_device, int B, int D, int H, int T, int precision, int out_layout, float* traj,
                 gode_stream_t stream) {
  if (bad_common(y0, W1, b1, W2, b2, B, T, out_layout) || !dt || !traj) return GODE_ERR_ARG;
  if (precision == GODE_PREC_BF16 && tc_wide_shape(D, H))
    return tc_rk4_fwd_wide(y0, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, out_layout, traj, (cudaStream_t)stream);
  if (precision == GODE_PREC_TF32 || precision == GODE_PREC_BF16) {
    if (!tc_shape(D, H)) return GODE_ERR_SHAPE;
    return tc_rk4_fwd(y0, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, precision, out_layout, traj, (cudaStream_t)stream);
  }
  if (precision != GODE_PREC_FP32) return GODE_ERR_PRECISION;
  if (wide_shape(D, H))
    return wide_rk4_fwd(y0, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, out_layout, traj, (cudaStream_t)stream);
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_fwd(y0, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, out_layout, traj, (cudaStream_t)stream);
}

size_t gode_bwd_workspace_bytes(int B, int D, int H) {
  (void)B;
  return bwd_workspace_bytes(gode_param_count(D, H));
}

// Kernels that use their workspace as plain scratch (no grid synchronisation) get the part BEHIND the persistent sync
// region, so that a workspace shared with the cooperative kernels never has its counters or tagged words overwritten.
static inline void* scratch_of(void* workspace) { return ws_scratch(workspace); }
static inline size_t scratch_bytes(size_t ws_bytes) {
  return ws_bytes > (size_t)GODE_SYNC_REGION_BYTES ? ws_bytes - (size_t)GODE_SYNC_REGION_BYTES : 0;
}

size_t gode_rk4_bwd_workspace_bytes(int B, int D, int H, int T) {
  if (wide_shape(D, H)) {
    const size_t a = wide_bwd_workspace_bytes(B, D, H, T), b = tc_wide_shape(D, H) ? tc_rk4_adj_wide_workspace_bytes(B) : 0;
    return (size_t)GODE_SYNC_REGION_BYTES + (a > b ? a : b);
  }
  {
    const size_t a = bwd_workspace_bytes(gode_param_count(D, H)),
                 b = tc_shape(D, H) ? (size_t)GODE_SYNC_REGION_BYTES + tc_rk4_adj_small_workspace_bytes(B) : 0;
    return a > b ? a : b;
  }
}

static int rk4_bwd_common(bool adjoint, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                          const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H,
                          int T, int precision, int layout, float* grad_y0, float* grad_params, void* workspace,
                          size_t ws_bytes, gode_stream_t stream) {
  if (bad_common(traj, W1, b1, W2, b2, B, T, layout) || !grad_traj || !dt || !grad_y0 || !grad_params || !workspace)
    return GODE_ERR_ARG;
  if (precision == GODE_PREC_BF16 && tc_wide_shape(D, H) && adjoint)  // tensor-core continuous adjoint (wide field)
    return tc_rk4_adj_wide(traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                           scratch_of(workspace), scratch_bytes(ws_bytes), (cudaStream_t)stream);
  if (precision == GODE_PREC_BF16 && tc_shape(D, H) && adjoint)       // tensor-core continuous adjoint (reference shape)
    return tc_rk4_adj_small(traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                            scratch_of(workspace), scratch_bytes(ws_bytes), (cudaStream_t)stream);
  if (precision != GODE_PREC_FP32) return GODE_ERR_PRECISION;
  if (wide_shape(D, H)) {
    if (!adjoint)   // backprop through the solver for the wide fields (round 2)
      return wide_rk4_backprop_bwd(traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                                   scratch_of(workspace), scratch_bytes(ws_bytes), (cudaStream_t)stream);
    return wide_rk4_adjoint_bwd(traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0,
                                grad_params, scratch_of(workspace), scratch_bytes(ws_bytes), (cudaStream_t)stream);
  }
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_bwd(adjoint, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0,
                       grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_rk4_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                         const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int precision,
                         int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                         gode_stream_t stream) {
  return rk4_bwd_common(true, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, precision, layout,
                        grad_y0, grad_params, workspace, ws_bytes, stream);
}

int gode_rk4_backprop_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int precision,
                          int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                          gode_stream_t stream) {
  return rk4_bwd_common(false, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, precision, layout,
                        grad_y0, grad_params, workspace, ws_bytes, stream);
}

int gode_rk4_sampler_fwd(const float* pre_Wa, const float* pre_ba, const float* pre_Wb, const float* pre_bb, float pre_slope,
                         int pre_hidden, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                         int dt_on_device, int B, int D, int H, int T, uint64_t seed, int64_t traj_offset,
                         const int64_t* traj_ids, int out_layout, float* out, int ld_out, float* noise_out,
                         gode_stream_t stream) {
  if (!W1 || !b1 || !W2 || !b2 || !dt || !out || B <= 0 || T < 2 ||
      (out_layout != GODE_LAYOUT_TBD && out_layout != GODE_LAYOUT_BTD) || (pre_Wa && (!pre_ba || !pre_Wb || !pre_bb)) ||
      (ld_out != 0 && (out_layout != GODE_LAYOUT_BTD || ld_out < D || (ld_out & 1))) || (reinterpret_cast<uintptr_t>(out) & 7))
    return GODE_ERR_ARG;   // (rows are read / written as float2: even strides, 8-byte aligned base)
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_fused_sampler_fwd(pre_Wa, pre_ba, pre_Wb, pre_bb, pre_slope, pre_hidden, W1, b1, W2, b2, dt, dt_on_device, B,
                                     D, H, T, seed, traj_offset, reinterpret_cast<const long long*>(traj_ids), out_layout, out,
                                     ld_out, noise_out, (cudaStream_t)stream);
}

int gode_rk4_adjoint_bwd_strided(const float* traj, int ld_traj, const float* grad_traj, int ld_grad, const float* W1,
                                 const float* b1, const float* W2, const float* b2, const float* dt, int dt_on_device, int B,
                                 int D, int H, int T, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                                 gode_stream_t stream) {
  if (bad_common(traj, W1, b1, W2, b2, B, T, GODE_LAYOUT_BTD) || !grad_traj || !dt || !grad_y0 || !grad_params || !workspace ||
      ld_traj < D || ld_grad < D || (ld_traj & 1) || (ld_grad & 1) || (reinterpret_cast<uintptr_t>(traj) & 7) ||
      (reinterpret_cast<uintptr_t>(grad_traj) & 7))
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_bwd(true, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, GODE_LAYOUT_BTD, grad_y0,
                       grad_params, workspace, ws_bytes, (cudaStream_t)stream, GODE_METHOD_RK4, nullptr, ld_traj, ld_grad);
}

int gode_fixed_adjoint_bwd_substep(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                                   const float* W2, const float* b2, const float* sub_dt, const int32_t* sub_beg,
                                   const int32_t* sub_end, int B, int D, int H, int T, int layout, float* grad_y0,
                                   float* grad_params, void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (bad_common(traj, W1, b1, W2, b2, B, T, layout) || !grad_traj || !sub_dt || !sub_beg || !sub_end || !grad_y0 ||
      !grad_params || !workspace)
    return GODE_ERR_ARG;
  if (method != GODE_METHOD_RK4 && method != GODE_METHOD_EULER && method != GODE_METHOD_MIDPOINT) return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  SubSteps sub{sub_dt, sub_beg, sub_end};
  return rk4_small_bwd(true, traj, grad_traj, W1, b1, W2, b2, nullptr, 1, B, D, H, T, layout, grad_y0, grad_params, workspace,
                       ws_bytes, (cudaStream_t)stream, method, nullptr, 0, 0, &sub);
}

int gode_rk4_bwd_world(int adjoint, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                       const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                       int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                       const GodeWorld* exchange, gode_stream_t stream) {
  if (bad_common(traj, W1, b1, W2, b2, B, T, layout) || !grad_traj || !dt || !grad_y0 || !grad_params || !workspace ||
      !exchange)
    return GODE_ERR_ARG;
  if (exchange->world < 1 || exchange->world > 64 || exchange->rank < 0 || exchange->rank >= exchange->world ||
      !exchange->slots_dev || !exchange->launch_ctr)
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_bwd(adjoint != 0, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0,
                       grad_params, workspace, ws_bytes, (cudaStream_t)stream, GODE_METHOD_RK4, exchange);
}

size_t gode_dopri5_workspace_bytes(int B, int D, int H) {
  if (wide_shape(D, H)) return wide_dopri5_workspace_bytes(B, D, H, 0, 0);
  const size_t a = dopri5_small_workspace_bytes(B, D, H), b = bwd_workspace_bytes(gode_param_count(D, H));
  return a > b ? a : b;
}

size_t gode_dopri5_backprop_workspace_bytes(int B, int D, int H, int ckpt_capacity) {
  if (wide_shape(D, H)) return wide_dopri5_workspace_bytes(B, D, H, ckpt_capacity, 1);
  return gode_dopri5_workspace_bytes(B, D, H);
}

int gode_dopri5_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                    const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                    float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                    float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes,
                    gode_stream_t stream) {
  if (bad_common(y0, W1, b1, W2, b2, B, T, out_layout) || !t_host || !opts || !traj || !log || !workspace)
    return GODE_ERR_ARG;
  if (opts->log_capacity > 0 && (!att_t0 || !att_dt || !att_er || !att_acc)) return GODE_ERR_ARG;
  if (opts->ckpt_capacity > 0 && (!ckpt || !acc_t0 || !acc_dt)) return GODE_ERR_ARG;
  if (opts->norm_scope != GODE_NORM_BATCH) return GODE_ERR_ARG;
  for (int i = 1; i < T; ++i)
    if (!(t_host[i] > t_host[i - 1])) return GODE_ERR_ARG;
  if (wide_shape(D, H))
    return wide_dopri5_fwd(y0, W1, b1, W2, b2, t_host, B, D, H, T, opts, out_layout, traj, log, att_t0, att_dt, att_er,
                           att_acc, ckpt, acc_t0, acc_dt, workspace, ws_bytes, (cudaStream_t)stream);
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_small_fwd(y0, W1, b1, W2, b2, t_host, B, D, H, T, opts, out_layout, traj, log, att_t0, att_dt, att_er,
                          att_acc, ckpt, acc_t0, acc_dt, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_dopri5_fwd_world(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                          float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                          float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes,
                          const GodeWorld* world, gode_stream_t stream) {
  if (bad_common(y0, W1, b1, W2, b2, B, T, out_layout) || !t_host || !opts || !traj || !log || !workspace || !world)
    return GODE_ERR_ARG;
  if (world->world < 1 || world->world > 32 || world->rank < 0 || world->rank >= world->world || world->total_B < B ||
      !world->slots_dev || !world->launch_ctr)
    return GODE_ERR_ARG;
  if (opts->log_capacity > 0 && (!att_t0 || !att_dt || !att_er || !att_acc)) return GODE_ERR_ARG;
  if (opts->ckpt_capacity > 0 && (!ckpt || !acc_t0 || !acc_dt)) return GODE_ERR_ARG;
  if (opts->norm_scope != GODE_NORM_BATCH) return GODE_ERR_ARG;
  for (int i = 1; i < T; ++i)
    if (!(t_host[i] > t_host[i - 1])) return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_small_fwd(y0, W1, b1, W2, b2, t_host, B, D, H, T, opts, out_layout, traj, log, att_t0, att_dt, att_er,
                          att_acc, ckpt, acc_t0, acc_dt, workspace, ws_bytes, (cudaStream_t)stream, world);
}

int gode_dopri5_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                             const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                             const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                             int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                             size_t ws_bytes, gode_stream_t stream) {
  if (bad_common(grad_traj, W1, b1, W2, b2, B, T, layout) || !t_host || !log || !ckpt || !acc_t0 || !acc_dt ||
      ckpt_capacity <= 0 || !grad_y0 || !grad_params || !workspace)
    return GODE_ERR_ARG;
  if (wide_shape(D, H))
    return wide_dopri5_backprop_bwd(grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, log, ckpt, acc_t0, acc_dt,
                                    ckpt_capacity, fsign, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_small_backprop_bwd(grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, log, ckpt, acc_t0, acc_dt,
                                   ckpt_capacity, fsign, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_adaptive_backprop_bwd(int tableau, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                               const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                               const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                               int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                               size_t ws_bytes, gode_stream_t stream) {
  if (bad_common(grad_traj, W1, b1, W2, b2, B, T, layout) || !t_host || !log || !ckpt || !acc_t0 || !acc_dt ||
      ckpt_capacity <= 0 || !grad_y0 || !grad_params || !workspace || tableau < GODE_TAB_DOPRI5 ||
      tableau > GODE_TAB_ADAPTIVE_HEUN)
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_small_backprop_bwd(grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, log, ckpt, acc_t0, acc_dt,
                                   ckpt_capacity, fsign, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream,
                                   nullptr, tableau);
}

int gode_dopri5_backprop_bwd_world(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                                   const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                                   int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                                   size_t ws_bytes, const GodeWorld* exchange, gode_stream_t stream) {
  if (bad_common(grad_traj, W1, b1, W2, b2, B, T, layout) || !t_host || !log || !ckpt || !acc_t0 || !acc_dt ||
      ckpt_capacity <= 0 || !grad_y0 || !grad_params || !workspace || !exchange)
    return GODE_ERR_ARG;
  if (exchange->world < 1 || exchange->world > 64 || exchange->rank < 0 || exchange->rank >= exchange->world ||
      !exchange->slots_dev || !exchange->launch_ctr)
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_small_backprop_bwd(grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, log, ckpt, acc_t0, acc_dt,
                                   ckpt_capacity, fsign, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream,
                                   exchange);
}

size_t gode_dopri5_adjoint_workspace_bytes(int B, int D, int H) { return dopri5_small_adjoint_workspace_bytes(B, D, H); }

int gode_dopri5_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                            const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                            const GodeAdaptiveOpts* opts, int param_mask, float* grad_y0, float* grad_params,
                            GodeStepLog* log, double* att_dt, float* att_er, uint8_t* att_acc, void* workspace,
                            size_t ws_bytes, gode_stream_t stream) {
  if (bad_common(traj, W1, b1, W2, b2, B, T, layout) || !grad_traj || !t_host || !opts || !grad_y0 || !grad_params ||
      !workspace)
    return GODE_ERR_ARG;
  if (att_dt && (!att_er || !att_acc)) return GODE_ERR_ARG;
  for (int i = 1; i < T; ++i)
    if (!(t_host[i] > t_host[i - 1])) return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  if (param_mask < 0 || param_mask > 15) return GODE_ERR_ARG;
  return dopri5_small_adjoint_bwd(traj, grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, opts, param_mask, grad_y0, grad_params,
                                  log, att_dt, att_er, att_acc, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_dopri5_traj_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                         const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts,
                         int out_layout, float* traj, GodeStepLog* log, int32_t* n_acc, int32_t* n_att,
                         double* att_dt, float* att_er, uint8_t* att_acc, float* ckpt, double* acc_t0,
                         double* acc_dt, gode_stream_t stream) {
  if (bad_common(y0, W1, b1, W2, b2, B, T, out_layout) || !t_host || !opts || !traj || !log || !n_acc || !n_att)
    return GODE_ERR_ARG;
  if (opts->log_capacity > 0 && (!att_dt || !att_er || !att_acc)) return GODE_ERR_ARG;
  if (opts->ckpt_capacity > 0 && (!ckpt || !acc_t0 || !acc_dt)) return GODE_ERR_ARG;
  for (int i = 1; i < T; ++i)
    if (!(t_host[i] > t_host[i - 1])) return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_traj_small_fwd(y0, W1, b1, W2, b2, t_host, B, D, H, T, opts, out_layout, traj, log, n_acc, n_att,
                               opts->log_capacity > 0 ? att_dt : nullptr, att_er, att_acc, ckpt, acc_t0, acc_dt,
                               (cudaStream_t)stream);
}

int gode_dopri5_traj_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                                  const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                                  const GodeStepLog* log, const int32_t* n_acc, const float* ckpt,
                                  const double* acc_t0, const double* acc_dt, int ckpt_capacity, float fsign,
                                  float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                                  gode_stream_t stream) {
  if (bad_common(grad_traj, W1, b1, W2, b2, B, T, layout) || !t_host || !log || !n_acc || !ckpt || !acc_t0 || !acc_dt ||
      ckpt_capacity <= 0 || !grad_y0 || !grad_params || !workspace)
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return dopri5_traj_small_bwd(grad_traj, W1, b1, W2, b2, t_host, B, D, H, T, layout, log, n_acc, ckpt, acc_t0, acc_dt,
                               ckpt_capacity, fsign, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

size_t gode_sde_workspace_bytes(int B, int D, int H) {
  (void)B;
  return sde_small_workspace_bytes(D, H);
}

static bool bad_nets(const float* const* a, const float* const* b) {
  if (!a || !b) return true;
  for (int i = 0; i < 4; ++i)
    if (!a[i] || !b[i]) return true;
  return false;
}

int gode_sde_em_fwd(const float* y0, const float* const* drift, const float* const* diffusion, const float* h_host,
                    int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host, int B, int D,
                    int H, int T, const float* dW, uint64_t seed, int64_t traj_offset, int out_layout, float* frames,
                    float* states, gode_stream_t stream) {
  if (!y0 || bad_nets(drift, diffusion) || !h_host || !out_step_host || !w0_host || !w1_host || !frames || B <= 0 ||
      T < 2 || n_steps < 1 || (out_layout != GODE_LAYOUT_TBD && out_layout != GODE_LAYOUT_BTD))
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return sde_small_fwd(y0, drift, diffusion, h_host, n_steps, out_step_host, w0_host, w1_host, B, D, H, T, dW, seed,
                       traj_offset, out_layout, frames, states, (cudaStream_t)stream);
}

int gode_sde_em_fwd_cells(const float* y0, const float* const* drift, const float* const* diffusion, const float* h_host,
                          int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host,
                          const int* fwd_lo_host, const float* cell_sqrt_host, int R, int B, int D, int H, int T,
                          const float* dW_cells, uint64_t seed, int64_t traj_offset, int out_layout, float* frames,
                          gode_stream_t stream) {
  if (!y0 || bad_nets(drift, diffusion) || !h_host || !out_step_host || !w0_host || !w1_host || !fwd_lo_host ||
      !cell_sqrt_host || !frames || B <= 0 || T < 2 || n_steps < 1 || R < 1 ||
      (out_layout != GODE_LAYOUT_TBD && out_layout != GODE_LAYOUT_BTD))
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return sde_small_fwd(y0, drift, diffusion, h_host, n_steps, out_step_host, w0_host, w1_host, B, D, H, T, dW_cells, seed,
                       traj_offset, out_layout, frames, nullptr, (cudaStream_t)stream, fwd_lo_host, cell_sqrt_host, R);
}

int gode_sde_adjoint_bwd(const float* frames, const float* grad_frames, const float* const* drift,
                         const float* const* diffusion, int n_rev, const float* h_rev_host, const int* rev_lo_host,
                         const int* rev_hi_host, const int* ibeg_host, const int* iend_host, const float* cell_sqrt_host,
                         int R, int B, int D, int H, int T, const float* dW_cells, uint64_t seed, int64_t traj_offset,
                         int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                         gode_stream_t stream) {
  if (!frames || !grad_frames || bad_nets(drift, diffusion) || !h_rev_host || !rev_lo_host || !rev_hi_host || !ibeg_host ||
      !iend_host || !cell_sqrt_host || !grad_y0 || !grad_params || !workspace || B <= 0 || T < 2 || n_rev < 1 || R < 1 ||
      (layout != GODE_LAYOUT_TBD && layout != GODE_LAYOUT_BTD))
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return sde_small_adjoint_bwd(frames, grad_frames, drift, diffusion, n_rev, h_rev_host, rev_lo_host, rev_hi_host, ibeg_host,
                               iend_host, cell_sqrt_host, R, B, D, H, T, dW_cells, seed, traj_offset, layout, grad_y0,
                               grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_sde_em_bwd(const float* states, const float* grad_frames, const float* const* drift,
                    const float* const* diffusion, const float* h_host, int n_steps, const int* out_step_host,
                    const float* w0_host, const float* w1_host, int B, int D, int H, int T, const float* dW,
                    uint64_t seed, int64_t traj_offset, int layout, float* grad_y0, float* grad_params,
                    void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (!states || !grad_frames || bad_nets(drift, diffusion) || !h_host || !out_step_host || !w0_host || !w1_host ||
      !grad_y0 || !grad_params || !workspace || B <= 0 || T < 2 || n_steps < 1 ||
      (layout != GODE_LAYOUT_TBD && layout != GODE_LAYOUT_BTD))
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return sde_small_bwd(states, grad_frames, drift, diffusion, h_host, n_steps, out_step_host, w0_host, w1_host, B, D, H,
                       T, dW, seed, traj_offset, layout, grad_y0, grad_params, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_gru_param_count(int D) { return 2 * 3 * D * D + 2 * 3 * D; }

int gode_gru_jump_fwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, int B, int D, float* h_out, gode_stream_t stream) {
  if (!x || !h || !w_ih || !w_hh || !b_ih || !b_hh || !h_out || B <= 0) return GODE_ERR_ARG;
  return gru_jump_fwd(x, h, w_ih, w_hh, b_ih, b_hh, B, D, h_out, (cudaStream_t)stream);
}

int gode_gru_jump_bwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, const float* grad_out, int B, int D, float* grad_x, float* grad_h,
                      float* grad_params, void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (!x || !h || !w_ih || !w_hh || !b_ih || !b_hh || !grad_out || !grad_h || !grad_params || !workspace || B <= 0)
    return GODE_ERR_ARG;
  return gru_jump_bwd(x, h, w_ih, w_hh, b_ih, b_hh, grad_out, nullptr, B, D, grad_x, grad_h, grad_params, 0, workspace,
                      ws_bytes, (cudaStream_t)stream);
}

size_t gode_odernn_log_stride(int log_capacity) { return odernn_log_stride(log_capacity); }
size_t gode_odernn_workspace_bytes(int B, int D, int H) { return odernn_workspace_bytes(B, D, H); }

int gode_odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H,
                    int F, const GodeAdaptiveOpts* opts, float* codes, float* seg, void* logs, float* ckpt, double* acc,
                    int32_t* n_acc, void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (!h0 || !eps || !W1 || !b1 || !W2 || !b2 || !w_ih || !w_hh || !b_ih || !b_hh || !opts || !codes || !seg || !logs ||
      !workspace || B <= 0 || F < 1)
    return GODE_ERR_ARG;
  if (opts->ckpt_capacity > 0 && (!ckpt || !acc)) return GODE_ERR_ARG;
  if (opts->log_capacity < 0 || (opts->norm_scope == GODE_NORM_TRAJ && !n_acc)) return GODE_ERR_ARG;
  return odernn_fwd(h0, eps, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh, B, D, H, F, opts, codes, seg,
                    reinterpret_cast<unsigned char*>(logs), ckpt, acc, n_acc, workspace, ws_bytes, (cudaStream_t)stream);
}

int gode_odernn_bwd(const float* grad_codes, const float* eps, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B,
                    int D, int H, int F, int log_capacity, int ckpt_capacity, const float* seg, const void* logs,
                    const float* ckpt, const double* acc, const int32_t* n_acc, const GodeAdaptiveOpts* adjoint_opts,
                    int adjoint_param_mask, float* grad_h0, float* grad_eps, float* grad_ode, float* grad_gru,
                    float* scratch, void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (!grad_codes || !eps || !W1 || !b1 || !W2 || !b2 || !w_ih || !w_hh || !b_ih || !b_hh || !seg || !grad_h0 ||
      !grad_ode || !grad_gru || !scratch || !workspace || B <= 0 || F < 1)
    return GODE_ERR_ARG;
  if (!adjoint_opts && (!logs || !ckpt || !acc || ckpt_capacity <= 0)) return GODE_ERR_ARG;
  if (adjoint_opts && (adjoint_param_mask < 0 || adjoint_param_mask > 15)) return GODE_ERR_ARG;
  return odernn_bwd(grad_codes, eps, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh, B, D, H, F, ckpt_capacity, seg,
                    reinterpret_cast<const unsigned char*>(logs), odernn_log_stride(log_capacity), ckpt, acc, n_acc,
                    adjoint_opts, adjoint_param_mask, grad_h0, grad_eps, grad_ode, grad_gru, scratch, workspace, ws_bytes,
                    (cudaStream_t)stream);
}

int gode_allreduce_p2p(float* data, int n, void* const* bufs_dev, void* const* pads_dev, int rank, int world, int cap,
                       uint32_t* epoch_ctr, gode_stream_t stream) {
  if (!data || n <= 0 || !bufs_dev || !pads_dev || !epoch_ctr) return GODE_ERR_ARG;
  return p2p_allreduce(data, n, reinterpret_cast<float* const*>(bufs_dev), reinterpret_cast<unsigned int* const*>(pads_dev),
                       rank, world, cap, epoch_ctr, (cudaStream_t)stream);
}

static bool bad_method(int m) { return m != GODE_METHOD_RK4 && m != GODE_METHOD_EULER && m != GODE_METHOD_MIDPOINT; }

int gode_fixed_fwd(int method, const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                   const float* dt, int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj,
                   gode_stream_t stream) {
  if (bad_method(method) || bad_common(y0, W1, b1, W2, b2, B, T, out_layout) || !dt || !traj) return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_fwd(y0, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, out_layout, traj, (cudaStream_t)stream, method);
}

static int fixed_bwd(bool adjoint, int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                     const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                     int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, gode_stream_t stream) {
  if (bad_method(method) || bad_common(traj, W1, b1, W2, b2, B, T, layout) || !grad_traj || !dt || !grad_y0 || !grad_params ||
      !workspace)
    return GODE_ERR_ARG;
  if (!small_field_shape(D, H)) return GODE_ERR_SHAPE;
  return rk4_small_bwd(adjoint, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                       workspace, ws_bytes, (cudaStream_t)stream, method);
}

int gode_fixed_adjoint_bwd(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                           const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                           int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                           gode_stream_t stream) {
  return fixed_bwd(true, method, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                   workspace, ws_bytes, stream);
}

int gode_fixed_backprop_bwd(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                            const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                            int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                            gode_stream_t stream) {
  return fixed_bwd(false, method, traj, grad_traj, W1, b1, W2, b2, dt, dt_on_device, B, D, H, T, layout, grad_y0, grad_params,
                   workspace, ws_bytes, stream);
}

}  // extern "C"
