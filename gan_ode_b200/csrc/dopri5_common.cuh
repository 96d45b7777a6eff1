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
// bosh3.py — Bogacki–Shampine 3(2) (FSAL) and adaptive_heun.py — Heun–Euler 2(1) (not FSAL), same construction
__device__ constexpr float kBsBeta[3][3] = {{F32(1.0 / 2), 0, 0}, {0, F32(3.0 / 4), 0}, {F32(2.0 / 9), F32(1.0 / 3), F32(4.0 / 9)}};
__device__ constexpr float kBsCErr[4] = {F32(2.0 / 9 - 7.0 / 24), F32(1.0 / 3 - 1.0 / 4), F32(4.0 / 9 - 1.0 / 3), F32(-1.0 / 8)};
__device__ constexpr float kBsCMid[4] = {0, F32(0.5), 0, 0};
__device__ constexpr float kAhBeta[1][1] = {{1.f}};
__device__ constexpr float kAhCSol[2] = {0.5f, 0.5f};
__device__ constexpr float kAhCErr[2] = {0.5f, -0.5f};
__device__ constexpr float kAhCMid[2] = {0.5f, 0.f};
#undef F32

// Tableau of an adaptive solver (torchdiffeq RKAdaptiveStepsizeODESolver subclasses): NS stage evaluations per step (k has
// NS + 1 entries, k[0] = f0), ORDER for the controller / the initial-step heuristic, FSAL: y1 is the last stage input
// (c_sol[:-1] == beta[-1], c_sol[-1] == 0), otherwise y1 = y0 + dt sum_j c_sol[j] k[j].  In both cases f1 = k[NS] goes to
// the next step as its f0 (rk_common.py::_runge_kutta_step) — for adaptive_heun that is f(t1, y0 + dt k0), not f(t1, y1).
template <int TAB>
struct Tableau;
template <>
struct Tableau<GODE_TAB_DOPRI5> {
  static constexpr int NS = 6, ORDER = 5;
  static constexpr bool FSAL = true;
  __device__ static constexpr float beta(int i, int j) { return kBeta[i][j]; }
  __device__ static constexpr float cerr(int j) { return kCErr[j]; }
  __device__ static constexpr float cmid(int j) { return kCMid[j]; }
  __device__ static constexpr float csol(int j) { return j < 6 ? kBeta[5][j] : 0.f; }
};
template <>
struct Tableau<GODE_TAB_BOSH3> {
  static constexpr int NS = 3, ORDER = 3;
  static constexpr bool FSAL = true;
  __device__ static constexpr float beta(int i, int j) { return kBsBeta[i][j]; }
  __device__ static constexpr float cerr(int j) { return kBsCErr[j]; }
  __device__ static constexpr float cmid(int j) { return kBsCMid[j]; }
  __device__ static constexpr float csol(int j) { return j < 3 ? kBsBeta[2][j] : 0.f; }
};
template <>
struct Tableau<GODE_TAB_ADAPTIVE_HEUN> {
  static constexpr int NS = 1, ORDER = 2;
  static constexpr bool FSAL = false;
  __device__ static constexpr float beta(int i, int j) { return kAhBeta[i][j]; }
  __device__ static constexpr float cerr(int j) { return kAhCErr[j]; }
  __device__ static constexpr float cmid(int j) { return kAhCMid[j]; }
  __device__ static constexpr float csol(int j) { return kAhCSol[j]; }
};

// misc.py::_optimal_step_size in fp64
// er^(-1/5) without fp64 exp/log (they were ~1.2 k cycles per attempted step on the solver's critical path, trace in
// profiles/README.md): MUFU-based fp32 guess (relative error ~1e-6), then a series correction of its residual (below) —
// division-free.
template <int ORDER = 5>
__device__ __forceinline__ double inv_fifth_root(float er32) {   // er^(-1/ORDER)
  // With the guess y and its residual r = 1 - er y^ORDER (|r| < ~1e-4), the root is y (1 - r)^(-1/ORDER)
  //   = y (1 + a r + a(a+1)/2 r^2 + a(a+1)(a+2)/6 r^3 + O(r^4)),  a = 1/ORDER:
  // one pass of ~ORDER+4 dependent fp64 operations, within 2 ulp of the correctly rounded value (checked against a 200-bit
  // reference over er in [1e-8, 1e4] with the guess perturbed by 1e-5).  It replaced two Newton steps (twice the chain).
  const double y = (double)__powf(er32, -1.f / (float)ORDER);
  const double er = (double)er32;
  double yn = y;
#pragma unroll
  for (int q = 1; q < ORDER; ++q) yn *= y;
  constexpr double a = 1.0 / (double)ORDER, c2 = a * (a + 1.0) / 2.0, c3 = a * (a + 1.0) * (a + 2.0) / 6.0;
  const double r = fma(-er, yn, 1.0);
  return fma(y, r * fma(r, fma(r, c3, c2), a), y);
}

template <int ORDER = 5>
__device__ __forceinline__ double optimal_step(double dt, float er32, const GodeAdaptiveOpts& o) {
  if (er32 == 0.f) return dt * o.ifactor;
  if (er32 != er32) return (double)er32;                       // torch.min/max propagate NaN; fmin/fmax do not
  const double dfactor = er32 < 1.f ? 1.0 : o.dfactor;
  // below 1e-30 (and for +inf) the factor is pinned by ifactor / dfactor whatever the root is; keep the guess in range
  const float erc = fminf(fmaxf(er32, 1e-30f), 1e30f);
  const double factor = fmin(o.ifactor, fmax(o.safety * inv_fifth_root<ORDER>(erc), dfactor));
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
