// wide_rk4.cu — FP32 fixed-grid RK4 (3/8 rule) forward + continuous adjoint for WIDE fields (D, H multiples of 32,
// e.g. the D=64 / H=256 "larger motion latent" of BASELINE.json configs[3]), where the lane-split kernels'
// register-resident weights and gradient accumulators no longer fit.
//
//   * one WARP per trajectory: lane l owns state components d = l + 32*dl and hidden units j = l + 32*jl (cyclic);
//   * W1 (H x (D+4)) and W2 (D x (H+4)) live once per CTA in shared memory (136 KB at 64/256).  The +4 padding makes all
//     four access patterns conflict-free: rows by lane with LDS.128 (forward layers), columns by lane with LDS.32
//     (g_h = a W2, vjp = delta W1);
//   * full vectors (u, h, a, delta) are exchanged through per-warp shared-memory lines (broadcast reads);
//   * parameter gradients: each (trajectory, interval, stage) writes its RK-weighted cotangent/activation rows
//     (c*a, h, c*delta, u) to a scratch buffer and a second kernel contracts them over all rows
//     (dW2 = (c a)^T h, dW1 = (c delta)^T u, db = column sums) with a fixed split-K order -> deterministic.
// Numerics: same operation order as the small-field kernels; parity bar <= 1e-5 relative.
#include "wide_field.cuh"

namespace gode {

constexpr float kWThird = 0.33333334f;

struct WideArgs {
  const float *y0, *W1, *b1, *W2, *b2;
  const float* traj_in;
  const float* grad_traj;
  float* traj;
  float* grad_y0;
  float *sa, *sh, *sd, *su;  // scratch rows: (N, D), (N, H), (N, H), (N, D) with N = B*(T-1)*4
  const float* dt_dev;
  int B, T, layout;
  float dt_val[GODE_MAX_HOST_STEPS];
};

template <int D, int H>
__global__ void __launch_bounds__(kWideWarps * 32) wide_rk4_fwd_kernel(const __grid_constant__ WideArgs p) {
  using W = Wide<D, H>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, l = tid & 31, warp = tid >> 5;
  W::stage(smem, p.W1, p.b1, p.W2, p.b2, tid, kWideWarps * 32);
  W w;
  w.bind(smem, warp);
  __syncthreads();
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  for (int b = blockIdx.x * kWideWarps + warp; b < p.B; b += gridDim.x * kWideWarps) {
    float y[W::DL], k1[W::DL], k2[W::DL], k3[W::DL], k4[W::DL], u[W::DL], hk[W::HL];
#pragma unroll
    for (int dl = 0; dl < W::DL; ++dl) {
      y[dl] = p.y0[(size_t)b * D + l + 32 * dl];
      p.traj[w_off(p.layout, 0, b, p.B, p.T, D) + l + 32 * dl] = y[dl];
    }
    for (int s = 0; s + 1 < p.T; ++s) {
      const float dt = dtp[s];
      w.forward(l, y, k1, hk);
#pragma unroll
      for (int i = 0; i < W::DL; ++i) u[i] = y[i] + dt * k1[i] * kWThird;
      w.forward(l, u, k2, hk);
#pragma unroll
      for (int i = 0; i < W::DL; ++i) u[i] = y[i] + dt * (k2[i] - k1[i] * kWThird);
      w.forward(l, u, k3, hk);
#pragma unroll
      for (int i = 0; i < W::DL; ++i) u[i] = y[i] + dt * (k1[i] - k2[i] + k3[i]);
      w.forward(l, u, k4, hk);
#pragma unroll
      for (int i = 0; i < W::DL; ++i) {
        y[i] = y[i] + (k1[i] + 3.f * (k2[i] + k3[i]) + k4[i]) * dt * 0.125f;
        p.traj[w_off(p.layout, s + 1, b, p.B, p.T, D) + l + 32 * i] = y[i];
      }
    }
  }
}

// continuous adjoint, one 3/8 step of the augmented system per interval (same algebra as rk4_adjoint_bwd_kernel);
// the parameter part is emitted as scratch rows instead of register accumulators
template <int D, int H>
__global__ void __launch_bounds__(kWideWarps * 32) wide_rk4_adjoint_kernel(const __grid_constant__ WideArgs p) {
  using W = Wide<D, H>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, l = tid & 31, warp = tid >> 5;
  W::stage(smem, p.W1, p.b1, p.W2, p.b2, tid, kWideWarps * 32);
  W w;
  w.bind(smem, warp);
  __syncthreads();
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  for (int b = blockIdx.x * kWideWarps + warp; b < p.B; b += gridDim.x * kWideWarps) {
    float a[W::DL];
#pragma unroll
    for (int i = 0; i < W::DL; ++i) a[i] = p.grad_traj[w_off(p.layout, p.T - 1, b, p.B, p.T, D) + l + 32 * i];
    for (int i = p.T - 1; i >= 1; --i) {
      const float dt = dtp[i - 1];
      const float cs[4] = {dt * 0.125f, 3.f * dt * 0.125f, 3.f * dt * 0.125f, dt * 0.125f};
      float y[W::DL], gprev[W::DL], uy[W::DL], ua[W::DL], f[W::DL], v[W::DL], hk[W::HL], dl_[W::HL];
      float k1y[W::DL], k1a[W::DL], k2y[W::DL], k2a[W::DL], k3y[W::DL], k3a[W::DL];
#pragma unroll
      for (int c = 0; c < W::DL; ++c) {
        y[c] = p.traj_in[w_off(p.layout, i, b, p.B, p.T, D) + l + 32 * c];
        gprev[c] = p.grad_traj[w_off(p.layout, i - 1, b, p.B, p.T, D) + l + 32 * c];
        uy[c] = y[c]; ua[c] = a[c];
      }
      const size_t row0 = ((size_t)b * (p.T - 1) + (i - 1)) * 4;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        w.forward(l, uy, f, hk);
        w.vjp(l, hk, ua, v, dl_);
        const size_t r = row0 + s;
#pragma unroll
        for (int c = 0; c < W::DL; ++c) { p.sa[r * D + l + 32 * c] = cs[s] * ua[c]; p.su[r * D + l + 32 * c] = uy[c]; }
#pragma unroll
        for (int c = 0; c < W::HL; ++c) { p.sh[r * H + l + 32 * c] = hk[c]; p.sd[r * H + l + 32 * c] = cs[s] * dl_[c]; }
#pragma unroll
        for (int c = 0; c < W::DL; ++c) {
          const float ky = -f[c], ka = v[c];
          if (s == 0) { k1y[c] = ky; k1a[c] = ka; uy[c] = y[c] + dt * ky * kWThird; ua[c] = a[c] + dt * ka * kWThird; }
          if (s == 1) { k2y[c] = ky; k2a[c] = ka; uy[c] = y[c] + dt * (ky - k1y[c] * kWThird); ua[c] = a[c] + dt * (ka - k1a[c] * kWThird); }
          if (s == 2) { k3y[c] = ky; k3a[c] = ka; uy[c] = y[c] + dt * (k1y[c] - k2y[c] + ky); ua[c] = a[c] + dt * (k1a[c] - k2a[c] + ka); }
          if (s == 3) a[c] = a[c] + (k1a[c] + 3.f * (k2a[c] + k3a[c]) + ka) * dt * 0.125f + gprev[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < W::DL; ++c) p.grad_y0[(size_t)b * D + l + 32 * c] = a[c];
  }
}

// Backprop through the solver (autograd-through-odeint semantics) for the wide fields: per step the four stage
// inputs and tanh vectors are recomputed from the stored y_s exactly as the forward computed them, then the stages are
// walked 4..1 (same algebra as rk4_backprop_bwd_kernel); every VJP emits its rows (cotangent on k_s, h_s, delta_s, u_s) to
// the same scratch layout as the adjoint, so the same contraction kernels produce the parameter gradients.
template <int D, int H>
__global__ void __launch_bounds__(kWideWarps * 32) wide_rk4_backprop_kernel(const __grid_constant__ WideArgs p) {
  using W = Wide<D, H>;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, l = tid & 31, warp = tid >> 5;
  W::stage(smem, p.W1, p.b1, p.W2, p.b2, tid, kWideWarps * 32);
  W w;
  w.bind(smem, warp);
  __syncthreads();
  const float* __restrict__ dtp = p.dt_dev ? p.dt_dev : p.dt_val;
  for (int b = blockIdx.x * kWideWarps + warp; b < p.B; b += gridDim.x * kWideWarps) {
    float yb[W::DL];  // cotangent of y_{s+1}
#pragma unroll
    for (int c = 0; c < W::DL; ++c) yb[c] = p.grad_traj[w_off(p.layout, p.T - 1, b, p.B, p.T, D) + l + 32 * c];
    for (int s = p.T - 2; s >= 0; --s) {
      const float dt = dtp[s];
      float y[W::DL], gs[W::DL], k1[W::DL], k2[W::DL], k3[W::DL], k4[W::DL], u2[W::DL], u3[W::DL], u4[W::DL];
      float h1[W::HL], h2[W::HL], h3[W::HL], h4[W::HL];
#pragma unroll
      for (int c = 0; c < W::DL; ++c) {
        y[c] = p.traj_in[w_off(p.layout, s, b, p.B, p.T, D) + l + 32 * c];
        gs[c] = p.grad_traj[w_off(p.layout, s, b, p.B, p.T, D) + l + 32 * c];
      }
      // recompute (same expressions as wide_rk4_fwd_kernel)
      w.forward(l, y, k1, h1);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) u2[c] = y[c] + dt * k1[c] * kWThird;
      w.forward(l, u2, k2, h2);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) u3[c] = y[c] + dt * (k2[c] - k1[c] * kWThird);
      w.forward(l, u3, k3, h3);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) u4[c] = y[c] + dt * (k1[c] - k2[c] + k3[c]);
      w.forward(l, u4, k4, h4);
      (void)k4;
      // reverse
      const float c18 = dt * 0.125f, c38 = 3.f * c18, dt3 = dt * kWThird;
      float kb1[W::DL], kb2[W::DL], kb3[W::DL], kb4[W::DL], ub[W::DL], dl_[W::HL];
#pragma unroll
      for (int c = 0; c < W::DL; ++c) { kb1[c] = c18 * yb[c]; kb2[c] = c38 * yb[c]; kb3[c] = c38 * yb[c]; kb4[c] = c18 * yb[c]; }
      const size_t row0 = ((size_t)b * (p.T - 1) + s) * 4;
      auto emit = [&](int stage, const float (&cot)[W::DL], const float (&u)[W::DL], const float (&hk)[W::HL]) {
        const size_t r = row0 + stage;
#pragma unroll
        for (int c = 0; c < W::DL; ++c) { p.sa[r * D + l + 32 * c] = cot[c]; p.su[r * D + l + 32 * c] = u[c]; }
#pragma unroll
        for (int c = 0; c < W::HL; ++c) { p.sh[r * H + l + 32 * c] = hk[c]; p.sd[r * H + l + 32 * c] = dl_[c]; }
      };
      w.vjp(l, h4, kb4, ub, dl_);
      emit(3, kb4, u4, h4);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) { yb[c] += ub[c]; kb1[c] += dt * ub[c]; kb2[c] -= dt * ub[c]; kb3[c] += dt * ub[c]; }
      __syncwarp();
      w.vjp(l, h3, kb3, ub, dl_);
      emit(2, kb3, u3, h3);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) { yb[c] += ub[c]; kb2[c] += dt * ub[c]; kb1[c] -= dt3 * ub[c]; }
      __syncwarp();
      w.vjp(l, h2, kb2, ub, dl_);
      emit(1, kb2, u2, h2);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) { yb[c] += ub[c]; kb1[c] += dt3 * ub[c]; }
      __syncwarp();
      w.vjp(l, h1, kb1, ub, dl_);
      emit(0, kb1, y, h1);
#pragma unroll
      for (int c = 0; c < W::DL; ++c) yb[c] += ub[c] + gs[c];
      __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < W::DL; ++c) p.grad_y0[(size_t)b * D + l + 32 * c] = yb[c];
  }
}

// ---- contraction of the scratch rows: C[M][N] = sum_n A[n][M] * Bm[n][N], plus column sums of A -------------------------
// split-K: CTA s handles rows [s*KS, (s+1)*KS); 256 threads, each an 8 x (N/32)... kept simple: thread t owns output
// columns n = t % N_T .. and rows m in a strided set; partials go to [slices][M*N + M]; a second pass adds the slices
// in order (deterministic).
constexpr int kKS = 1024;

template <int M, int N>
__global__ void __launch_bounds__(256) wgrad_partial_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                            long long rows, float* __restrict__ partial) {
  // outputs per thread: M*N/256 ; thread t -> n = t % N' pattern below
  constexpr int OUT = M * N / 256;
  static_assert((M * N) % 256 == 0, "tile must divide over 256 threads");
  __shared__ __align__(16) float sA[16][M];
  __shared__ __align__(16) float sB[16][N];
  const int t = threadIdx.x;
  float acc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) acc[o] = 0.f;
  float colsum = 0.f;  // thread t < M accumulates column t of A
  const long long r0 = (long long)blockIdx.x * kKS;
  const long long r1 = r0 + kKS < rows ? r0 + kKS : rows;
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
        const int idx = o * 256 + t;  // flat output index m*N + n
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

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int slices, int len, float* __restrict__ out_w,
                                    int len_w, float* __restrict__ out_b) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  float s = 0.f;
  for (int k = 0; k < slices; ++k) s += partial[(size_t)k * len + e];
  if (e < len_w) out_w[e] = s;
  else out_b[e - len_w] = s;
}

// ---- host -----------------------------------------------------------------------------------------------------------------
bool wide_shape(int D, int H) { return (D == 64 && H == 256) || (D == 32 && H == 32) || (D == 32 && H == 64); }

static size_t wide_rows(int B, int T) { return (size_t)B * (T - 1) * 4; }

size_t wide_bwd_workspace_bytes(int B, int D, int H, int T) {
  const size_t rows = wide_rows(B, T);
  const size_t slices = (rows + kKS - 1) / kKS;
  const size_t scratch = rows * (size_t)(2 * D + 2 * H);
  const size_t part = slices * (size_t)(H * D + (H > D ? H : D));
  return sizeof(float) * (scratch + part) + 1024;
}

template <int D, int H>
static int launch_wide_fwd(WideArgs& a, cudaStream_t st) {
  auto kern = wide_rk4_fwd_kernel<D, H>;
  const size_t smem = Wide<D, H>::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(1000 + (int)e);
  int grid = (a.B + kWideWarps - 1) / kWideWarps;
  const int cap = sm_count() * (smem > 110 * 1024 ? 1 : 2);
  if (grid > cap) grid = cap;
  kern<<<grid, kWideWarps * 32, smem, st>>>(a);
  return launch_status();
}

template <int D, int H, bool ADJOINT = true>
static int launch_wide_bwd(WideArgs& a, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (ws_bytes < wide_bwd_workspace_bytes(a.B, D, H, a.T)) return GODE_ERR_WORKSPACE;
  const size_t rows = wide_rows(a.B, a.T);
  float* base = reinterpret_cast<float*>(workspace);
  a.sa = base; a.su = a.sa + rows * D; a.sh = a.su + rows * D; a.sd = a.sh + rows * H;
  float* partial = a.sd + rows * H;
  auto kern = ADJOINT ? wide_rk4_adjoint_kernel<D, H> : wide_rk4_backprop_kernel<D, H>;
  const size_t smem = Wide<D, H>::smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(1000 + (int)e);
  int grid = (a.B + kWideWarps - 1) / kWideWarps;
  const int cap = sm_count() * (smem > 110 * 1024 ? 1 : 2);
  if (grid > cap) grid = cap;
  kern<<<grid, kWideWarps * 32, smem, st>>>(a);
  if (int rc = launch_status()) return rc;
  const int slices = (int)((rows + kKS - 1) / kKS);
  // dW1 (H x D) = (c delta)^T u ; db1 = colsum(c delta)     flat layout [W1 | b1 | W2 | b2]
  wgrad_partial_kernel<H, D><<<slices, 256, 0, st>>>(a.sd, a.su, (long long)rows, partial);
  wgrad_reduce_kernel<<<(H * D + H + 255) / 256, 256, 0, st>>>(partial, slices, H * D + H, grad_params, H * D, grad_params + H * D);
  // dW2 (D x H) = (c a)^T h ; db2 = colsum(c a)
  wgrad_partial_kernel<D, H><<<slices, 256, 0, st>>>(a.sa, a.sh, (long long)rows, partial);
  wgrad_reduce_kernel<<<(D * H + D + 255) / 256, 256, 0, st>>>(partial, slices, D * H + D, grad_params + H * D + H, D * H,
                                                              grad_params + H * D + H + D * H);
  return launch_status();
}

static int fill_wide_dt(WideArgs& a, const float* dt, int dt_on_device, int T) {
  if (dt_on_device) { a.dt_dev = dt; return GODE_OK; }
  if (T - 1 > GODE_MAX_HOST_STEPS) return GODE_ERR_T_TOO_LONG;
  for (int i = 0; i < T - 1; ++i) a.dt_val[i] = dt[i];
  return GODE_OK;
}

int wide_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                 int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj, cudaStream_t st) {
  WideArgs a{};
  a.y0 = y0; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj = traj; a.B = B; a.T = T; a.layout = out_layout;
  if (int rc = fill_wide_dt(a, dt, dt_on_device, T)) return rc;
  if (D == 64 && H == 256) return launch_wide_fwd<64, 256>(a, st);
  if (D == 32 && H == 32) return launch_wide_fwd<32, 32>(a, st);
  if (D == 32 && H == 64) return launch_wide_fwd<32, 64>(a, st);
  return GODE_ERR_SHAPE;
}

int wide_rk4_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                         const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                         float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  WideArgs a{};
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj_in = traj; a.grad_traj = grad_traj; a.grad_y0 = grad_y0;
  a.B = B; a.T = T; a.layout = layout;
  if (int rc = fill_wide_dt(a, dt, dt_on_device, T)) return rc;
  if (D == 64 && H == 256) return launch_wide_bwd<64, 256>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 32) return launch_wide_bwd<32, 32>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 64) return launch_wide_bwd<32, 64>(a, grad_params, workspace, ws_bytes, st);
  return GODE_ERR_SHAPE;
}

int wide_rk4_backprop_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T, int layout,
                          float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes, cudaStream_t st) {
  WideArgs a{};
  a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.traj_in = traj; a.grad_traj = grad_traj; a.grad_y0 = grad_y0;
  a.B = B; a.T = T; a.layout = layout;
  if (int rc = fill_wide_dt(a, dt, dt_on_device, T)) return rc;
  if (D == 64 && H == 256) return launch_wide_bwd<64, 256, false>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 32) return launch_wide_bwd<32, 32, false>(a, grad_params, workspace, ws_bytes, st);
  if (D == 32 && H == 64) return launch_wide_bwd<32, 64, false>(a, grad_params, workspace, ws_bytes, st);
  return GODE_ERR_SHAPE;
}

}  // namespace gode
