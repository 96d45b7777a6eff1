// gru_cell.cuh — nn.GRUCell arithmetic of the ODE-RNN jump (models/mocogan_ode_rnn.py:49, models/mocogan.py:198) shared by the
// stand-alone jump kernels (odernn_small.cu) and the persistent all-frames sampler kernel (dopri5_small.cu): gate order
// r, z, n; n = tanh(W_in x + b_in + r (W_hn h + b_hn)); h_out = (1 - z) n + z h.  One function, so both paths round identically.
#pragma once
#include <cuda_runtime.h>

namespace gode {

constexpr int GD = 16;             // dim_z_motion of the reference (models/mocogan.py:198: GRUCell(16, 16))
constexpr int GT = 16;             // trajectories per tile
constexpr int GP = 2 * 3 * GD * GD + 2 * 3 * GD;  // flat [w_ih (3D,D) | w_hh (3D,D) | b_ih (3D) | b_hh (3D)] = 1632

constexpr int GS = GD + 1;         // padded row stride of the weights in shared memory (rows by lane: conflict-free)
struct GruW {
  float wih[3 * GD * GS], whh[3 * GD * GS], bih[3 * GD], bhh[3 * GD];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void load_w(GruW& w, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                       int tid, int n) {
  for (int e = tid; e < 3 * GD * GD; e += n) { w.wih[(e / GD) * GS + e % GD] = w_ih[e]; w.whh[(e / GD) * GS + e % GD] = w_hh[e]; }
  for (int e = tid; e < 3 * GD; e += n) { w.bih[e] = b_ih[e]; w.bhh[e] = b_hh[e]; }
}

// gates of unit j of trajectory t (x, h rows in shared memory)
struct Gates { float r, z, n, hn; };
__device__ __forceinline__ Gates gates(const GruW& w, const float* x, const float* h, int j) {
  float ir = w.bih[j], iz = w.bih[GD + j], in = w.bih[2 * GD + j];
  float hr = w.bhh[j], hz = w.bhh[GD + j], hn = w.bhh[2 * GD + j];
#pragma unroll
  for (int k = 0; k < GD; ++k) {
    const float xv = x[k], hv = h[k];
    ir = fmaf(w.wih[j * GS + k], xv, ir);
    iz = fmaf(w.wih[(GD + j) * GS + k], xv, iz);
    in = fmaf(w.wih[(2 * GD + j) * GS + k], xv, in);
    hr = fmaf(w.whh[j * GS + k], hv, hr);
    hz = fmaf(w.whh[(GD + j) * GS + k], hv, hz);
    hn = fmaf(w.whh[(2 * GD + j) * GS + k], hv, hn);
  }
  Gates g;
  g.r = sigmoidf_(ir + hr);
  g.z = sigmoidf_(iz + hz);
  g.hn = hn;
  g.n = tanhf(in + g.r * hn);
  return g;
}

}  // namespace gode
