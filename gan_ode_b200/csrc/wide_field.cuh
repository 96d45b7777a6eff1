// wide_field.cuh — the ODEFunc MLP for WIDE fields (D, H multiples of 32), one WARP per trajectory, weights once per CTA in
// shared memory; shared by the fixed-grid kernels (wide_rk4.cu) and the adaptive ones (wide_dopri5.cu).  Layout and access
// patterns: see the header comment of wide_rk4.cu.
#pragma once
#include "launch.h"
#include "small_field.cuh"

namespace gode {

// 8 warps per CTA.  16 were measured (round 2: the weights fill 136 KB of shared memory at D=64 / H=256, so a fatter CTA is the
// only way to more warps per scheduler): no gain at B >= 4096 (0.541 vs 0.549 ms for the dopri5 forward) and half the SMs idle
// at B = 1184 — the mat-vecs are bound by shared-memory operand BANDWIDTH (every warp streams all weights per evaluation), not
// by its latency.
constexpr int kWideWarps = 8;

template <int D, int H>
struct Wide {
  static_assert(D % 32 == 0 && H % 32 == 0, "wide kernels need D, H multiples of 32");
  static constexpr int DL = D / 32, HL = H / 32;
  static constexpr int DS = D + 4, HS = H + 4;
  static constexpr int kWeightFloats = H * DS + D * HS + H + D;
  static constexpr int kLineFloats = 2 * (D + H);  // y, h, a, delta
  static constexpr size_t smem_bytes() { return sizeof(float) * (kWeightFloats + kWideWarps * kLineFloats); }

  const float *w1, *w2, *b1, *b2;  // shared
  float *ly, *lh, *la, *ld;        // this warp's lines

  __device__ __forceinline__ void bind(float* smem, int warp) {
    w1 = smem; w2 = smem + H * DS; b1 = smem + H * DS + D * HS; b2 = b1 + H;
    float* lines = smem + kWeightFloats + warp * kLineFloats;
    ly = lines; lh = lines + D; la = lines + D + H; ld = lines + 2 * D + H;
  }
  __device__ static void stage(float* smem, const float* W1, const float* B1, const float* W2, const float* B2, int tid, int nthr) {
    for (int e = tid; e < H * D; e += nthr) smem[(e / D) * DS + (e % D)] = W1[e];
    for (int e = tid; e < D * H; e += nthr) smem[H * DS + (e / H) * HS + (e % H)] = W2[e];
    for (int e = tid; e < H; e += nthr) smem[H * DS + D * HS + e] = B1[e];
    for (int e = tid; e < D; e += nthr) smem[H * DS + D * HS + H + e] = B2[e];
  }

  // out = W2 tanh(W1 u + b1) + b2 ; leaves u in ly, tanh vector in lh
  __device__ __forceinline__ void forward(int l, const float (&u)[DL], float (&out)[DL], float (&hk)[HL]) const {
    __syncwarp();
#pragma unroll
    for (int dl = 0; dl < DL; ++dl) ly[l + 32 * dl] = u[dl];
    __syncwarp();
#pragma unroll
    for (int jl = 0; jl < HL; ++jl) {
      const float* row = w1 + (size_t)(l + 32 * jl) * DS;
      float s0 = b1[l + 32 * jl], s1 = 0.f;
#pragma unroll 4
      for (int i = 0; i < D; i += 8) {
        const float4 wa = *reinterpret_cast<const float4*>(row + i), va = *reinterpret_cast<const float4*>(ly + i);
        const float4 wb = *reinterpret_cast<const float4*>(row + i + 4), vb = *reinterpret_cast<const float4*>(ly + i + 4);
        s0 = fmaf(wa.x, va.x, s0); s0 = fmaf(wa.y, va.y, s0); s0 = fmaf(wa.z, va.z, s0); s0 = fmaf(wa.w, va.w, s0);
        s1 = fmaf(wb.x, vb.x, s1); s1 = fmaf(wb.y, vb.y, s1); s1 = fmaf(wb.z, vb.z, s1); s1 = fmaf(wb.w, vb.w, s1);
      }
      hk[jl] = tanhf(s0 + s1);
    }
#pragma unroll
    for (int jl = 0; jl < HL; ++jl) lh[l + 32 * jl] = hk[jl];
    __syncwarp();
#pragma unroll
    for (int dl = 0; dl < DL; ++dl) {
      const float* row = w2 + (size_t)(l + 32 * dl) * HS;
      float s0 = b2[l + 32 * dl], s1 = 0.f;
#pragma unroll 4
      for (int j = 0; j < H; j += 8) {
        const float4 wa = *reinterpret_cast<const float4*>(row + j), va = *reinterpret_cast<const float4*>(lh + j);
        const float4 wb = *reinterpret_cast<const float4*>(row + j + 4), vb = *reinterpret_cast<const float4*>(lh + j + 4);
        s0 = fmaf(wa.x, va.x, s0); s0 = fmaf(wa.y, va.y, s0); s0 = fmaf(wa.z, va.z, s0); s0 = fmaf(wa.w, va.w, s0);
        s1 = fmaf(wb.x, vb.x, s1); s1 = fmaf(wb.y, vb.y, s1); s1 = fmaf(wb.z, vb.z, s1); s1 = fmaf(wb.w, vb.w, s1);
      }
      out[dl] = s0 + s1;
    }
  }

  // right after forward(): vjp = (a W2 ⊙ (1-h^2)) W1 ; delta returned for the gradient rows
  __device__ __forceinline__ void vjp(int l, const float (&hk)[HL], const float (&a)[DL], float (&out)[DL], float (&delta)[HL]) const {
#pragma unroll
    for (int dl = 0; dl < DL; ++dl) la[l + 32 * dl] = a[dl];
    __syncwarp();
#pragma unroll
    for (int jl = 0; jl < HL; ++jl) {
      float s0 = 0.f, s1 = 0.f;
      const float* col = w2 + (l + 32 * jl);
#pragma unroll 8
      for (int d = 0; d < D; d += 2) {
        s0 = fmaf(la[d], col[(size_t)d * HS], s0);
        s1 = fmaf(la[d + 1], col[(size_t)(d + 1) * HS], s1);
      }
      delta[jl] = (s0 + s1) * (1.f - hk[jl] * hk[jl]);
      ld[l + 32 * jl] = delta[jl];
    }
    __syncwarp();
#pragma unroll
    for (int dl = 0; dl < DL; ++dl) {
      float s0 = 0.f, s1 = 0.f;
      const float* col = w1 + (l + 32 * dl);
#pragma unroll 8
      for (int j = 0; j < H; j += 2) {
        s0 = fmaf(ld[j], col[(size_t)j * DS], s0);
        s1 = fmaf(ld[j + 1], col[(size_t)(j + 1) * DS], s1);
      }
      out[dl] = s0 + s1;
    }
  }
};

__device__ __forceinline__ size_t w_off(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}


}  // namespace gode
