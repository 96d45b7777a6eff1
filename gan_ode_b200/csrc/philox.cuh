// philox.cuh — counter-based Philox4x32-10 normals shared by the SDE kernels and the fused latent sampler.
// Contract (oracle/philox.py::normals): key = (seed lo, seed hi); counter = (global trajectory, index, d_block, stream)
// -> 4 uint32 -> 2x Box–Muller -> 4 standard normals ordered (sin01, cos01, sin23, cos23) for components 4*d_block .. +3.
//   stream 0: one draw per Euler–Maruyama step (sdeint)            index = step
//   stream 1: Brownian cells of the stochastic adjoint             index = cell
//   stream 2: the motion noise x of sample_z_m (fused sampler)     index = 0
#pragma once
#include <cuda_runtime.h>

namespace gode {

// ---- Philox4x32-10 ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned int x) { return (float)x * 2.3283064e-10f + 1.1641532e-10f; }

// the lane's DL standard normals for (trajectory, step): components l*DL .. l*DL+DL-1
template <int DL>
__device__ __forceinline__ void philox_normals(unsigned long long seed, unsigned long long traj, int step, int l, float (&z)[DL],
                                               unsigned int stream = 0u) {
  static_assert(DL == 1 || DL == 2 || DL == 4, "lane slice must tile a 4-wide Philox block");
  const int d0 = l * DL;
  const uint4 r = philox4x32_10(make_uint4((unsigned int)traj, (unsigned int)step, (unsigned int)(d0 >> 2), stream),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  // Box–Muller on the pair(s) this lane needs only (a lane owning 1 or 2 components needs one of the block's two pairs),
  // with the hardware transcendental units: log2 (MUFU.LG2), rsqrt/sqrt, sin/cos of 2*pi*u via sinpi/cospi-style exact
  // range (u in (0,1) -> the argument of MUFU.SIN/COS stays in (0, 2*pi)).  Differences from the CPU contract
  // (oracle/philox.py, numpy float32) are ~1e-6 relative, far inside the 2e-5 the stream test allows.
  auto pair = [](unsigned int a, unsigned int b, float& n0, float& n1) {
    const float rad = sqrtf(-1.3862944f * __log2f(u01(a)));  // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
    float sn, cs;
    __sincosf(6.2831855f * u01(b), &sn, &cs);
    n0 = rad * sn;
    n1 = rad * cs;
  };
  if constexpr (DL == 4) {
    pair(r.x, r.y, z[0], z[1]);
    pair(r.z, r.w, z[2], z[3]);
  } else {
    const bool second = (d0 & 2) != 0;
    float n0, n1;
    pair(second ? r.z : r.x, second ? r.w : r.y, n0, n1);
    if constexpr (DL == 2) { z[0] = n0; z[1] = n1; }
    else z[0] = (d0 & 1) ? n1 : n0;
  }
}

}  // namespace gode
