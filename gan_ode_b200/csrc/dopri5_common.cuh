// dopri5_common.cuh — Dormand–Prince tableau, fp64 controller and the sign-carrying field wrapper shared by the
// batch-global and per-trajectory dopri5 kernels.
#pragma once
#include "small_field.cuh"
#include "launch.h"

namespace gode {

// dopri5.py tableau, built in fp64 and cast to fp32 exactly as torchdiffeq casts it to y0.dtype
#define F32(x) ((float)(x))
__device__ constexpr float kBeta[6][6] = {
    {F32(1.0 / 5), 0, 0, 0, 0, 0},
    {F32(3.0 / 40), F32(9.0 / 40), 0, 0, 0, 0},
    {F32(44.0 / 45), F32(-56.0 / 15), F32(32.0 / 9), 0, 0, 0},
    {F32(19372.0 / 6561), F32(-25360.0 / 2187), F32(64448.0 / 6561), F32(-212.0 / 729), 0, 0},
    {F32(9017.0 / 3168), F32(-355.0 / 33), F32(46732.0 / 5247), F32(49.0 / 176), F32(-5103.0 / 18656), 0},
    {F32(35.0 / 384), 0, F32(500.0 / 1113), F32(125.0 / 192), F32(-2187.0 / 6784), F32(11.0 / 84)}};
__device__ constexpr float kCErr[7] = {F32(35.0 / 384 - 1951.0 / 21600),
                                        0,
                                        F32(500.0 / 1113 - 22642.0 / 50085),
                                        F32(125.0 / 192 - 451.0 / 720),
                                        F32(-2187.0 / 6784 - -12231.0 / 42400),
                                        F32(11.0 / 84 - 649.0 / 6300),
                                        F32(-1.0 / 60)};
__device__ constexpr float kCMid[7] = {F32(6025192743.0 / 30085553152.0 / 2),
                                        0,
                                        F32(51252292925.0 / 65400821598.0 / 2),
                                        F32(-2691868925.0 / 45128329728.0 / 2),
                                        F32(187940372067.0 / 1594534317056.0 / 2),
                                        F32(-1776094331.0 / 19743644256.0 / 2),
                                        F32(11237099.0 / 235043384.0 / 2)};
#undef F32

// misc.py::_optimal_step_size in fp64
// er^(-1/5) without fp64 exp/log (they were ~1.2 k cycles per attempted step on the solver's critical path, trace in
// profiles/README.md): MUFU-based fp32 guess (relative error ~1e-6), then two Newton steps for F(y) = y^-5 - er,
// y <- y (6 - er y^5) / 5 — division-free, quadratic: 1e-6 -> 3e-12 -> 3e-23, i.e. correctly rounded to within an ulp.
__device__ __forceinline__ double inv_fifth_root(float er32) {
  double y = (double)__powf(er32, -0.2f);
  const double er = (double)er32;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double y2 = y * y, y5 = y2 * y2 * y;
    y = y * ((6.0 - er * y5) * 0.2);
  }
  return y;
}

__device__ __forceinline__ double optimal_step(double dt, float er32, const GodeAdaptiveOpts& o) {
  if (er32 == 0.f) return dt * o.ifactor;
  if (er32 != er32) return (double)er32;                       // torch.min/max propagate NaN; fmin/fmax do not
  const double dfactor = er32 < 1.f ? 1.0 : o.dfactor;
  // below 1e-30 (and for +inf) the factor is pinned by ifactor / dfactor whatever the root is; keep the guess in range
  const float erc = fminf(fmaxf(er32, 1e-30f), 1e30f);
  const double factor = fmin(o.ifactor, fmax(o.safety * inv_fifth_root(erc), dfactor));
  return dt * factor;
}

// the (possibly time-reversed) field: out = fsign * f(u)
template <int D, int H, int L, class Lines>
__device__ __forceinline__ void field(const RowWeights<D, H, L>& w, const Lines& ln, int l, float fsign,
                                      const float (&u)[Shape<D, H, L>::DL], float (&out)[Shape<D, H, L>::DL],
                                      float (&hk)[Shape<D, H, L>::HL]) {
  mlp_forward<D, H, L>(w, ln, l, u, out, hk);
#pragma unroll
  for (int c = 0; c < Shape<D, H, L>::DL; ++c) out[c] *= fsign;
}

}  // namespace gode
