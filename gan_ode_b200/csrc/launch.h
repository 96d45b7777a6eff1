// launch.h — host-side helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/gode.h"

namespace gode {

// Device properties are queried per call (cached per device id in a small immutable table; no mutable
// global state that a second thread could observe half-written: entries are written once with the same value).
inline int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GODE_OK : -(1000 + (int)e);
}

// backward kernels keep parameter-gradient accumulators in registers and loop over their trajectories; they are
// launched cooperatively (grid barrier before the final cross-CTA reduction), at most two 128-thread CTAs per SM.
inline int bwd_grid_cap() { return sm_count() * 2; }

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// [persistent sync region (GODE_SYNC_REGION_BYTES, zero-filled once by the owner) | per-CTA partial rows]
inline size_t bwd_workspace_bytes(int P) {
  const int cap = bwd_grid_cap();
  return (size_t)GODE_SYNC_REGION_BYTES + sizeof(float) * (size_t)P * (size_t)cap;
}
inline char* ws_scratch(void* workspace) { return reinterpret_cast<char*>(workspace) + GODE_SYNC_REGION_BYTES; }

// launch flags of the calling thread (gode_set_thread_launch_flags)
int& thread_launch_flags();
// process-wide status mailbox in mapped host memory (gode_set_status_mailbox); null = off
int32_t* status_mailbox();

// Cooperative launch — or, with `pdl`, an ORDINARY launch with programmatic stream serialisation (the kernel must call
// griddep_wait() before it touches anything the previous kernel on the stream wrote).  Measured on B200: a launch that is
// both cooperative and programmatic is accepted but never starts early (the backward's CTA 0 entered 0.25 us after the
// forward's exit) and the graph edge costs ~4 us more than a plain one, so the PDL variant gives the cooperative attribute
// up.  That is sound for these kernels: their grid is capped at the co-resident capacity of an EMPTY device
// (coop_limit), their grid barrier only waits for CTAs of the same grid, and the kernel they overlap (the forward) never
// waits for them — so every CTA gets an SM at the latest when the forward's CTAs retire.  If the driver refuses the
// attribute, the cooperative launch is used (remembered per kernel).
template <class K>
inline cudaError_t coop_launch(K kern, int grid, int block, void** args, size_t smem, cudaStream_t st, bool pdl) {
  static bool pdl_ok = true;
  if (pdl && pdl_ok) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();
    pdl_ok = false;
  }
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), args, smem, st);
}

// co-resident CTA limit of a kernel (cached per instantiation; racing writers store the same value)
template <class K>
inline int coop_limit(K kern, int threads, size_t smem, int& cache) {
  if (cache > 0) return cache;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm <= 0) return 0;
  cache = per_sm * sm_count();
  return cache;
}

// sub-step tables of a sub-stepped adjoint (options['step_size']), device resident
struct SubSteps {
  const float* dt;
  const int *beg, *end;
};

// entry points of the per-family translation units
int rk4_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                  int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st,
                  int method = GODE_METHOD_RK4);
int rk4_small_bwd(bool adjoint, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                  const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                  int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st,
                  int method = GODE_METHOD_RK4, const GodeWorld* xchg = nullptr, int ld_traj = 0, int ld_grad = 0,
                  const struct SubSteps* sub = nullptr);
int rk4_small_fused_sampler_fwd(const float* pre_Wa, const float* pre_ba, const float* pre_Wb, const float* pre_bb,
                                float pre_slope, int pre_hidden, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                                unsigned long long seed, long long traj_offset, const long long* traj_ids, int out_layout,
                                float* out, int ld_out, float* noise_out, cudaStream_t st);

size_t dopri5_small_workspace_bytes(int B, int D, int H);
int dopri5_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                     const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                     float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                     float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes, cudaStream_t st,
                     const GodeWorld* world = nullptr);
int dopri5_small_odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2,
                            const float* b2, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                            int B, int D, int H, int F, const GodeAdaptiveOpts* opts, float* codes, float* seg,
                            unsigned char* logs, size_t log_stride, float* ckpt, double* acc, void* workspace,
                            size_t ws_bytes, cudaStream_t st);
// dopri5_adj_small.cu — torchdiffeq's continuous adjoint with the adaptive solver
size_t dopri5_small_adjoint_workspace_bytes(int B, int D, int H);
int dopri5_small_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                             const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                             const GodeAdaptiveOpts* opts, int param_mask, float* grad_y0, float* grad_params,
                             GodeStepLog* log, double* att_dt, float* att_er, uint8_t* att_acc, void* workspace, size_t ws_bytes,
                             cudaStream_t st);
int dopri5_small_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                              const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                              const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                              int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                              size_t ws_bytes, cudaStream_t st, const GodeWorld* xchg = nullptr,
                              int tableau = GODE_TAB_DOPRI5);

int tc_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
               int dt_on_device, int B, int D, int H, int T, int precision, int out_layout, float* traj,
               cudaStream_t st);

size_t sde_small_workspace_bytes(int D, int H);
int sde_small_fwd(const float* y0, const float* const* fw, const float* const* gw, const float* h_host, int n_steps,
                  const int* out_step_host, const float* w0_host, const float* w1_host, int B, int D, int H, int T,
                  const float* dW, unsigned long long seed, long long traj_offset, int layout, float* out, float* states,
                  cudaStream_t st, const int* fwd_lo_host = nullptr, const float* cell_sqrt_host = nullptr, int R = 0);
int sde_small_adjoint_bwd(const float* frames, const float* grad_out, const float* const* fw, const float* const* gw,
                          int n_rev, const float* h_rev_host, const int* rev_lo_host, const int* rev_hi_host,
                          const int* ibeg_host, const int* iend_host, const float* cell_sqrt_host, int R, int B, int D, int H,
                          int T, const float* dW, unsigned long long seed, long long traj_offset, int layout, float* grad_y0,
                          float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);
int sde_small_bwd(const float* states, const float* grad_out, const float* const* fw, const float* const* gw,
                  const float* h_host, int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host,
                  int B, int D, int H, int T, const float* dW, unsigned long long seed, long long traj_offset, int layout,
                  float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);

int dopri5_traj_small_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                          float* traj, GodeStepLog* log, int32_t* n_acc, int32_t* n_att, double* att_dt, float* att_er,
                          uint8_t* att_acc, float* ckpt, double* acc_t0, double* acc_dt, cudaStream_t st);
int dopri5_traj_small_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, int layout, const GodeStepLog* log,
                          const int32_t* n_acc, const float* ckpt, const double* acc_t0, const double* acc_dt,
                          int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                          size_t ws_bytes, cudaStream_t st);

bool wide_shape(int D, int H);
size_t wide_bwd_workspace_bytes(int B, int D, int H, int T);
int wide_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                 int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st);
int wide_rk4_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                         const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                         float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);

int wide_rk4_backprop_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                          float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);

// wide_dopri5.cu: dopri5 with the batch-global controller for the wide shapes (any batch size: state in global memory)
size_t wide_dopri5_workspace_bytes(int B, int D, int H, int ckpt_capacity, int backward);
int wide_dopri5_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                    const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts, int out_layout,
                    float* traj, GodeStepLog* log, double* att_t0, double* att_dt, float* att_er, uint8_t* att_acc,
                    float* ckpt, double* acc_t0, double* acc_dt, void* workspace, size_t ws_bytes, cudaStream_t st);
int wide_dopri5_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2, const float* b2,
                             const double* t_host, int B, int D, int H, int T, int layout, const GodeStepLog* log,
                             const float* ckpt, const double* acc_t0, const double* acc_dt, int ckpt_capacity, float fsign,
                             float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);

inline bool small_field_shape(int D, int H) { return D == 16 && H == 16; }
int tc_rk4_fwd_wide(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                    int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st);
size_t tc_rk4_adj_wide_workspace_bytes(int B);
int tc_rk4_adj_wide(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                    float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);
size_t odernn_log_stride(int log_capacity);
size_t odernn_workspace_bytes(int B, int D, int H);
int gru_jump_fwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, int B, int D, float* h_out, cudaStream_t st);
int gru_jump_bwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, const float* grad_out, const float* grad_carry, int B, int D, float* grad_x,
                 float* grad_h, float* grad_params, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t st);
int odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H, int F,
               const GodeAdaptiveOpts* opts, float* codes, float* seg, unsigned char* logs, float* ckpt, double* acc,
               int32_t* n_acc, void* workspace, size_t ws_bytes, cudaStream_t st);
int odernn_bwd(const float* grad_codes, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H, int F,
               int ckpt_capacity, const float* seg, const unsigned char* logs, size_t log_stride, const float* ckpt,
               const double* acc, const int32_t* n_acc, const GodeAdaptiveOpts* adjoint_opts, int adjoint_param_mask,
               float* grad_h0, float* grad_eps, float* grad_ode, float* grad_gru, float* scratch, void* workspace,
               size_t ws_bytes, cudaStream_t st);
int p2p_allreduce(float* data, int n, float* const* bufs_dev, unsigned int* const* pads_dev, int rank, int world, int cap,
                  unsigned int* epoch_ctr, cudaStream_t st);
size_t tc_rk4_adj_small_workspace_bytes(int B);
int tc_rk4_adj_small(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                     const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                     float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st);
inline bool tc_shape(int D, int H) { return D == 16 && H == 16; }
inline bool tc_wide_shape(int D, int H) { return D == 64 && H == 256; }  // BF16 only

}  // namespace gode
