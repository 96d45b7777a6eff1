// dopri5_adj_small.cu — torchdiffeq's CONTINUOUS adjoint for the adaptive solver (adjoint.py::OdeintAdjointMethod.backward
// with method = adjoint_method = 'dopri5'): the call the ODE-RNN sampler makes (models/mocogan_ode_rnn.py:47-48).
//
// Per output interval [t_i, t_{i-1}] torchdiffeq runs a fresh dopri5 solve of the augmented state
//     z = (vjp_t, y, a, theta_bar)         dz/dt = (., f(y), -a^T df/dy, -a^T df/dtheta)
// backwards in time (odeint negates time and field), with the default adjoint norm
//     max(|vjp_t|, rms(y), rms(a), max_k rms(theta_bar_k))        k over W1, b1, W2, b2
// in the initial-step heuristic and in every error ratio, then replaces y by the stored forward value and adds the
// upstream gradient of that output time to a.  The field is autonomous, so the vjp_t component is identically zero and is
// not carried.  What makes this different from the forward kernel: theta_bar's derivative is a sum over the WHOLE batch, and
// its error estimate enters the step controller, so every attempted step needs the batch-reduced theta-parts of the stage
// derivatives — as three fixed linear combinations (5th-order solution, error estimate, mid-point for the dense output)
// plus the last stage alone (FSAL: it is the first stage of the next step).
//
// One cooperative launch for the whole backward.  A trajectory is split over 8 lanes as in the other small-field kernels;
// y, a and their seven stage derivatives stay in registers.  The theta-part of a stage is the outer products
// delta (x) u, delta, cot (x) h, cot summed over the warp's 4 trajectories; their gather lines lie side by side in shared
// memory, so each lane forms one weight row (+ its bias) of the warp's sum directly — 17 of the warp's 544 values, no
// exchange — and folds it with the three tableau weights into per-warp accumulators in shared memory.  At the end of an attempt: warps -> CTA partial row in global memory,
// grid barrier, the float4 columns of the partial matrix are dealt to the warps of the grid and summed in a fixed order,
// grid barrier, every CTA reads the 2184 totals and evaluates the theta-part of the norm itself (same code on the same data:
// bit-identical in every CTA, so the accept/reject branch is uniform).  Deterministic: no atomics anywhere.
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include "dopri5_common.cuh"

namespace gode {

constexpr int kAdjMaxT = 256;
constexpr int kAdjWarps = 8;
constexpr int kAdjNV = 4;  // reduced theta-shaped vectors: 0 = last stage (K7), 1 = solution combo, 2 = error combo, 3 = mid-point combo
constexpr int kAdjNS = 8;  // reduced scalars (norm sums of the y / a parts, non-finite count)
constexpr int kAdjStages = 6;  // stage derivatives k_2..k_7 evaluated per attempted step

struct Dp5AdjArgs {
  const float *traj, *grad_traj, *W1, *b1, *W2, *b2;
  float *grad_y0, *grad_params;
  GodeStepLog* log;
  int32_t* mailbox;  // mapped host int (gode_set_status_mailbox) or null
  double* att_dt; float* att_er; uint8_t* att_acc;  // optional per-attempt log, all intervals concatenated
  GridSyncWs gs;
  GodeAdaptiveOpts o;
  int B, T, layout;
  int param_mask;  // bit k: parameter tensor k (W1, b1, W2, b2) is an adjoint parameter, i.e. enters the norm (adjoint.py)
  double t[kAdjMaxT];  // forward grid, increasing (already negated by the host for a decreasing t, o.fsign = -1)
};

__device__ __forceinline__ size_t adj_toff(int layout, int s, int b, int B, int T, int D) {
  return layout == GODE_LAYOUT_TBD ? ((size_t)s * B + b) * D : ((size_t)b * T + s) * D;
}

// interp.py::_interp_fit + _interp_evaluate for one element
__device__ __forceinline__ float interp_at(float y0, float y1, float ymid, float f0, float f1, float dt32, float x) {
  const float ca = 2.f * dt32 * (f1 - f0) - 8.f * (y1 + y0) + 16.f * ymid;
  const float cb = dt32 * (5.f * f0 - 3.f * f1) + 18.f * y0 + 14.f * y1 - 32.f * ymid;
  const float cc = dt32 * (f1 - 4.f * f0) - 11.f * y0 - 5.f * y1 + 16.f * ymid;
  const float cd = dt32 * f0;
  float tot = y0 + x * cd;
  float xp = x * x;
  tot = tot + xp * cc;
  xp = xp * x;
  tot = tot + xp * cb;
  xp = xp * x;
  tot = tot + xp * ca;
  return tot;
}

// Row-owned weights in SHARED memory (the other small-field kernels keep them in registers: here the registers go to the two
// state vectors' stage derivatives).  Physical row order (jl*L + l) as in ColWeights: the L lanes of a trajectory read
// consecutive padded rows (conflict-free LDS.128), lanes of other trajectories broadcast.
template <int D, int H, int L>
struct SmemRowWeights {
  using S = Shape<D, H, L>;
  static constexpr int kFloats = H * S::YS + H + D * S::HS + D;
  float *w1, *b1, *w2, *b2;
  __device__ __forceinline__ void bind(float* smem) { w1 = smem; b1 = w1 + H * S::YS; w2 = b1 + H; b2 = w2 + D * S::HS; }
  __device__ __forceinline__ void stage(const float* __restrict__ W1, const float* __restrict__ B1,
                                        const float* __restrict__ W2, const float* __restrict__ B2, int tid, int nthreads) {
    for (int e = tid; e < H * D; e += nthreads) {
      { const int j = e / D, i = e % D; w1[((j % S::HL) * L + j / S::HL) * S::YS + i] = W1[e]; }
      { const int d = e / H, j = e % H; w2[((d % S::DL) * L + d / S::DL) * S::HS + j] = W2[e]; }
    }
    for (int e = tid; e < H; e += nthreads) b1[e] = B1[e];
    for (int e = tid; e < D; e += nthreads) b2[e] = B2[e];
  }
};

template <int N>
__device__ __forceinline__ float dot_smem_init(const float* __restrict__ wrow, const float* __restrict__ line, float init) {
  float2 s = make_float2(init, 0.f);
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + i);
    const float4 v = *reinterpret_cast<const float4*>(line + i);
    s = __ffma2_rn(make_float2(w.x, w.y), make_float2(v.x, v.y), s);
    s = __ffma2_rn(make_float2(w.z, w.w), make_float2(v.z, v.w), s);
  }
  return s.x + s.y;
}

template <int D, int H, int L, int WARPS>
struct AdjLayout {
  using S = Shape<D, H, L>;
  static constexpr int P = S::P;
  static constexpr int Q = D + 1;                    // values of the warp's theta sum owned by one lane: a weight row + its bias
  static constexpr int VT = kAdjNV * P + kAdjNS;
  static_assert(D == H && H == 16 && 32 * Q == P, "lanes 0..15 own (W1 row j | b1[j]), lanes 16..31 own (W2 row d | b2[d])");
  static_assert(VT % 4 == 0 && P % 4 == 0, "float4 columns");
  static constexpr int kSmemFloats = WARPS * BwdLines<D, H, L>::kFloatsPerWarp + ColWeights<D, H, L>::kFloats +
                                     SmemRowWeights<D, H, L>::kFloats +
                                     (L == 8 ? kAdjStages : kAdjNV) * WARPS * P + 2 * P + VT + WARPS * kAdjNS + WARPS * 4;
  // native index n = lane*Q + q -> index in the flat [W1|b1|W2|b2] vector, tensor id 0..3
  __device__ static __forceinline__ int canon(int n, int& tensor) {
    const int lane = n / Q, q = n % Q, r = lane & 15;
    if (lane < 16) {
      if (q < D) { tensor = 0; return r * D + q; }
      tensor = 1;
      return H * D + r;
    }
    if (q < H) { tensor = 2; return H * D + H + r * H + q; }
    tensor = 3;
    return H * D + H + D * H + r;
  }
};

template <int D, int H, int L, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) dopri5_adjoint_bwd_kernel(const __grid_constant__ Dp5AdjArgs p) {
  using S = Shape<D, H, L>;
  using BL = BwdLines<D, H, L>;
  using A = AdjLayout<D, H, L, WARPS>;
  constexpr int P = A::P, Q = A::Q, VT = A::VT, NT = WARPS * 32;
  extern __shared__ __align__(16) float smem[];
  float* s_lines = smem;
  float* s_cw = s_lines + WARPS * BL::kFloatsPerWarp;
  float* s_rw = s_cw + ColWeights<D, H, L>::kFloats;
  // Theta-part of the augmented stage derivatives, summed over the warp's trajectories.  Two layouts:
  //   kSlots (8-lane mapping): one slot per stage k_2..k_7, [stage][WARPS][P]; the tableau combinations (solution, error,
  //     mid-point) are formed once per attempt when the CTA's row is written — 17 stores per lane and stage instead of 51 loads
  //     + 51 stores (stage glue 8.0k -> 3.9k cycles per attempt at B = 1024);
  //   else (4- and 2-lane mappings): the combinations are accumulated stage by stage, [kAdjNV][WARPS][P].  With eight
  //     trajectories per warp the slot layout measured SLOWER in total (+4 % at B = 8192: what the glue saved came back in the
  //     stage evaluations), so it is kept for the mapping where it pays.
  constexpr bool kSlots = L == 8;
  float* s_k = s_rw + SmemRowWeights<D, H, L>::kFloats;
  float* s_th = s_k + (kSlots ? kAdjStages : kAdjNV) * WARPS * P;   // theta_bar at the start of the current step (native order)
  float* s_k1 = s_th + P;                              // theta-part of the first stage (FSAL)
  float* s_tot = s_k1 + P;                             // [VT] grid totals of the last reduction
  float* s_sc = s_tot + VT;                            // [WARPS][kAdjNS]
  float* s_r4 = s_sc + WARPS * kAdjNS;                 // [WARPS][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / L, l = lane % L;
  BL ln;
  ln.bind(s_lines + warp * BL::kFloatsPerWarp, g);
  ColWeights<D, H, L> cw;
  cw.bind(s_cw);
  cw.stage(p.W1, p.W2, tid, NT);
  SmemRowWeights<D, H, L> rw;
  rw.bind(s_rw);
  rw.stage(p.W1, p.b1, p.W2, p.b2, tid, NT);
  for (int n = tid; n < P; n += NT) s_th[n] = 0.f;
  __syncthreads();

  const int b = (blockIdx.x * WARPS + warp) * S::G + g;
  const bool valid = b < p.B;
  const bool logger = (blockIdx.x == 0 && tid == 0);
  const float n_elem = (float)p.B * (float)D;
  const float rtol32 = (float)p.o.rtol, atol32 = (float)p.o.atol;
  const float asign = -p.o.fsign;  // the adjoint runs against the forward direction
  SyncState ss;
  ss.begin(p.gs);
  float* my_k = s_k + (size_t)warp * P + lane * Q;  // + (stage or vector) * WARPS * P

#ifdef GODE_ADJ_TIMING
  long long tc_stage = 0, tc_red = 0, tc_rest = 0, tc_f = 0, tc_v = 0, tc_t = 0, tc_a = 0, tc_mark = clock64();
#define GADJ_TICK(acc) { const long long now_ = clock64(); acc += now_ - tc_mark; tc_mark = now_; }
#else
#define GADJ_TICK(acc)
#endif
  // (dy/ds, da/ds) and the reduce-scattered theta-part of the augmented field at (u, ua)
  auto eval_aug = [&](const float (&u)[S::DL], const float (&ua)[S::DL], float (&ky)[S::DL], float (&ka)[S::DL],
                      float (&r2)[Q]) {
    float hk[S::HL], f[S::DL], cot[S::DL], delta[S::HL];
    GADJ_TICK(tc_stage);
    store_frag<S::DL>(ln.y + l * S::DL, u);
    __syncwarp();
#pragma unroll
    for (int jl = 0; jl < S::HL; ++jl)
      hk[jl] = tanhf(dot_smem_init<D>(rw.w1 + (jl * L + l) * S::YS, ln.y, rw.b1[l * S::HL + jl]));
    store_frag<S::HL>(ln.h + l * S::HL, hk);
    __syncwarp();
#pragma unroll
    for (int dl = 0; dl < S::DL; ++dl) f[dl] = dot_smem_init<H>(rw.w2 + (dl * L + l) * S::HS, ln.h, rw.b2[l * S::DL + dl]);
    GADJ_TICK(tc_f);
#pragma unroll
    for (int c = 0; c < S::DL; ++c) { ky[c] = asign * f[c]; cot[c] = -asign * ua[c]; }
    store_frag<S::DL>(ln.a + l * S::DL, cot);
    __syncwarp();
#pragma unroll
    for (int jl = 0; jl < S::HL; ++jl) {
      const float gh = dot_smem<D>(cw.w2t + (jl * L + l) * S::YS, ln.a);
      delta[jl] = gh * (1.f - hk[jl] * hk[jl]);
    }
    store_frag<S::HL>(ln.dl + l * S::HL, delta);
    __syncwarp();
#pragma unroll
    for (int dl = 0; dl < S::DL; ++dl) ka[dl] = dot_smem<H>(cw.w1t + (dl * L + l) * S::HS, ln.dl);
    GADJ_TICK(tc_v);
    // theta-part, summed over the warp's trajectories without any exchange: the lines of all G trajectories sit side by side
    // in this warp's shared-memory block, so lane r < 16 forms row r of dW1 = sum_g delta_g[r] u_g (and db1[r]), lane 16 + r
    // row r of dW2 = sum_g cot_g[r] h_g (and db2[r]) with broadcast reads — 17 of the warp's 544 values per lane.
    {
      const float* wl = s_lines + warp * BL::kFloatsPerWarp;
      const bool lower = lane < 16;
      const int r = lane & 15;
      const float* sv = lower ? wl + S::G * (2 * S::YS + S::HS) + r : wl + S::G * (S::YS + S::HS) + r;  // delta_g[r] : cot_g[r]
      const float* lv = lower ? wl : wl + S::G * S::YS;                                              // u_g : h_g
      const int svs = lower ? S::HS : S::YS, lvs = lower ? S::YS : S::HS;
#pragma unroll
      for (int q = 0; q < Q; ++q) r2[q] = 0.f;
#pragma unroll
      for (int gg = 0; gg < S::G; ++gg) {
        const float sg = sv[gg * svs];
#pragma unroll
        for (int i = 0; i < D; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(lv + gg * lvs + i);
          r2[i] = fmaf(sg, v.x, r2[i]); r2[i + 1] = fmaf(sg, v.y, r2[i + 1]);
          r2[i + 2] = fmaf(sg, v.z, r2[i + 2]); r2[i + 3] = fmaf(sg, v.w, r2[i + 3]);
        }
        r2[D] += sg;
      }
    }
    __syncwarp();  // all lanes are done with ln.y / ln.h before the next evaluation overwrites them
    GADJ_TICK(tc_t);
  };

  // Sum `nvec` accumulator vectors and the per-lane scalars over the whole grid; totals land in s_tot.
  // No barrier: every CTA stores its VT sums as tagged 64-bit words {fp32 | tag} into ITS row of the persistent region, the
  // warp that owns a float4 column polls that column in all rows (rows added in CTA order, then a fixed shuffle tree: the same
  // order as before), stores the four totals as tagged words, and every CTA polls the totals.  Two L2 write -> read latencies
  // per reduction instead of two fence + atomic + poll + fence barriers around the same traffic.  Reuse is safe without
  // double buffering: a CTA writes its row for reduction n+1 only after it has seen ALL totals of n, i.e. after every column
  // owner has finished reading the rows of n; an owner writes total n+1 only after it has seen every CTA's row n+1, each
  // written after that CTA consumed the totals of n.  Tags count across launches (SyncState), rows are only ever written
  // as tagged words.
  auto grid_reduce = [&](int nvec, float (&sc)[kAdjNS], float dt32) {
#pragma unroll
    for (int k = 0; k < kAdjNS; ++k) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) sc[k] += __shfl_xor_sync(0xffffffffu, sc[k], off);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kAdjNS; ++k) s_sc[warp * kAdjNS + k] = sc[k];
    }
    __syncthreads();
    const unsigned int tag = ++ss.epoch;
    const unsigned long long tag_hi = (unsigned long long)tag << 32;
    unsigned long long* mine = p.gs.rows + (size_t)blockIdx.x * VT;
    // tableau weights of the theta-part of k_j, j = st + 2 in 1-based stage numbering (k_1 carries zero weight in all three
    // combinations): solution = row 5 of beta, error, mid-point
    float wsol[kAdjStages], werr[kAdjStages], wmid[kAdjStages];
#pragma unroll
    for (int st = 0; st < kAdjStages; ++st) {
      wsol[st] = st < 5 ? kBeta[5][st + 1] * dt32 : 0.f;
      werr[st] = kCErr[st + 1] * dt32;
      wmid[st] = kCMid[st + 1] * dt32;
    }
    if constexpr (!kSlots) {
      for (int n = tid; n < P; n += NT) {
        for (int v = 0; v < nvec; ++v) {
          const float* col = s_k + (size_t)v * WARPS * P + n;
          float x = col[0];
#pragma unroll
          for (int q = 1; q < WARPS; ++q) x += col[q * P];
          st_relaxed_u64(mine + v * P + n, (unsigned long long)__float_as_uint(x) | tag_hi);
        }
      }
    } else {
    for (int n = tid; n < P; n += NT) {
      // per warp: the combinations in stage order (zero, then one fma per stage: the order the per-stage accumulation had),
      // then the warps in order
      float xk = 0.f, xs = 0.f, xe = 0.f, xm = 0.f;
#pragma unroll
      for (int q = 0; q < WARPS; ++q) {
        const float* col = s_k + (size_t)q * P + n;
        const float k7 = col[(size_t)(kAdjStages - 1) * WARPS * P];
        xk = q == 0 ? k7 : xk + k7;
        if (nvec > 1) {
          float as = 0.f, ae = 0.f, am = 0.f;
#pragma unroll
          for (int st = 0; st < kAdjStages; ++st) {
            const float v = st == kAdjStages - 1 ? k7 : col[(size_t)st * WARPS * P];
            as = fmaf(wsol[st], v, as);
            ae = fmaf(werr[st], v, ae);
            am = fmaf(wmid[st], v, am);
          }
          xs = q == 0 ? as : xs + as;
          xe = q == 0 ? ae : xe + ae;
          xm = q == 0 ? am : xm + am;
        }
      }
      st_relaxed_u64(mine + n, (unsigned long long)__float_as_uint(xk) | tag_hi);
      if (nvec > 1) {
        st_relaxed_u64(mine + P + n, (unsigned long long)__float_as_uint(xs) | tag_hi);
        st_relaxed_u64(mine + 2 * P + n, (unsigned long long)__float_as_uint(xe) | tag_hi);
        if (nvec > 3) st_relaxed_u64(mine + 3 * P + n, (unsigned long long)__float_as_uint(xm) | tag_hi);
      }
    }
    }
    if (tid < kAdjNS) {
      float x = s_sc[tid];
#pragma unroll
      for (int q = 1; q < WARPS; ++q) x += s_sc[q * kAdjNS + tid];
      st_relaxed_u64(mine + kAdjNV * P + tid, (unsigned long long)__float_as_uint(x) | tag_hi);
    }
    const int nb = gridDim.x, gw = blockIdx.x * WARPS + warp, nw = nb * WARPS;
    unsigned long long* totals = p.gs.rows + (size_t)nb * VT;
    for (int c4 = gw; c4 < VT / 4; c4 += nw) {
      if (c4 >= nvec * (P / 4) && c4 < kAdjNV * (P / 4)) continue;
      const unsigned long long* col = p.gs.rows + 4 * c4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      // kRows rows per lane with every load in flight (four where the registers allow it: the 8- and 4-lane instantiations); rows
      // added in order
      constexpr int kRows = L >= 4 ? 4 : 2;
      for (int r0 = lane; r0 < nb; r0 += 32 * kRows) {
        unsigned long long x[kRows][4];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int r = r0 + 32 * u;
          if (r < nb) {
            ld_relaxed_v2_u64(col + (size_t)r * VT, x[u][0], x[u][1]);
            ld_relaxed_v2_u64(col + (size_t)r * VT + 2, x[u][2], x[u][3]);
          }
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int r = r0 + 32 * u;
          if (r < nb) {
            while ((unsigned int)(x[u][0] >> 32) != tag || (unsigned int)(x[u][1] >> 32) != tag ||
                   (unsigned int)(x[u][2] >> 32) != tag || (unsigned int)(x[u][3] >> 32) != tag) {
              ld_relaxed_v2_u64(col + (size_t)r * VT, x[u][0], x[u][1]);
              ld_relaxed_v2_u64(col + (size_t)r * VT + 2, x[u][2], x[u][3]);
            }
            s.x += __uint_as_float((unsigned int)x[u][0]); s.y += __uint_as_float((unsigned int)x[u][1]);
            s.z += __uint_as_float((unsigned int)x[u][2]); s.w += __uint_as_float((unsigned int)x[u][3]);
          }
        }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, off); s.y += __shfl_xor_sync(0xffffffffu, s.y, off);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, off); s.w += __shfl_xor_sync(0xffffffffu, s.w, off);
      }
      if (lane < 4) {
        const float v = lane == 0 ? s.x : lane == 1 ? s.y : lane == 2 ? s.z : s.w;
        st_relaxed_u64(totals + 4 * c4 + lane, (unsigned long long)__float_as_uint(v) | tag_hi);
      }
    }
    for (int c4 = tid; c4 < VT / 4; c4 += NT) {
      if (c4 >= nvec * (P / 4) && c4 < kAdjNV * (P / 4)) continue;
      unsigned long long a0, a1, a2, a3;
      do {
        ld_relaxed_v2_u64(totals + 4 * c4, a0, a1);
        ld_relaxed_v2_u64(totals + 4 * c4 + 2, a2, a3);
      } while ((unsigned int)(a0 >> 32) != tag || (unsigned int)(a1 >> 32) != tag || (unsigned int)(a2 >> 32) != tag ||
               (unsigned int)(a3 >> 32) != tag);
      reinterpret_cast<float4*>(s_tot)[c4] = make_float4(__uint_as_float((unsigned int)a0), __uint_as_float((unsigned int)a1),
                                                         __uint_as_float((unsigned int)a2), __uint_as_float((unsigned int)a3));
    }
    __syncthreads();
  };

  // misc.py::_mixed_norm over the four parameter tensors of ratio(n), n in native order; identical in every thread of the grid
  auto theta_norm = [&](auto&& ratio) -> float {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = tid; n < P; n += NT) {
      int ts;
      A::canon(n, ts);
      const float r = ratio(n);
      const float r2 = r * r;
      s4[0] += ts == 0 ? r2 : 0.f; s4[1] += ts == 1 ? r2 : 0.f; s4[2] += ts == 2 ? r2 : 0.f; s4[3] += ts == 3 ? r2 : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s4[k] += __shfl_xor_sync(0xffffffffu, s4[k], off);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s_r4[warp * 4 + k] = s4[k];
    }
    __syncthreads();
    float t4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      t4[k] = s_r4[k];
#pragma unroll
      for (int q = 1; q < WARPS; ++q) t4[k] += s_r4[q * 4 + k];
    }
    __syncthreads();
    const int m = p.param_mask;
    return fmaxf(fmaxf(m & 1 ? sqrtf(t4[0] / (float)(H * D)) : 0.f, m & 2 ? sqrtf(t4[1] / (float)H) : 0.f),
                 fmaxf(m & 4 ? sqrtf(t4[2] / (float)(D * H)) : 0.f, m & 8 ? sqrtf(t4[3] / (float)D) : 0.f));
  };

  float y[S::DL], a[S::DL], ky[7][S::DL], ka[7][S::DL], r2[Q];
#pragma unroll
  for (int c = 0; c < S::DL; ++c) { y[c] = 0.f; a[c] = 0.f; }
  if (valid) {
    load_frag<S::DL>(p.traj + adj_toff(p.layout, p.T - 1, b, p.B, p.T, D) + l * S::DL, y);
    load_frag<S::DL>(p.grad_traj + adj_toff(p.layout, p.T - 1, b, p.B, p.T, D) + l * S::DL, a);
  }
  int status = 0, n_att = 0, n_acc = 0, nfe = 0;
  double dt = 0.0, dt0 = 0.0, t0 = 0.0;

  for (int i = p.T - 1; i >= 1 && status == 0; --i) {
    const double s_begin = -p.t[i], s_end = -p.t[i - 1];
    float sc[kAdjNS];
    float scy[S::DL], sca[S::DL];
#pragma unroll
    for (int c = 0; c < S::DL; ++c) { scy[c] = 1.f; sca[c] = 1.f; }
    // ---- _before_integrate: f0 and misc.py::_select_initial_step on the augmented state ---------------------------------
    // (the two field evaluations go through ONE loop body so that eval_aug is instantiated once here)
    float d0 = 0.f, d1 = 0.f, h0 = 0.f;
    float f1y[S::DL], f1a[S::DL];
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      float u[S::DL], ua[S::DL];
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        u[c] = ph == 0 ? y[c] : y[c] + h0 * ky[0][c];
        ua[c] = ph == 0 ? a[c] : a[c] + h0 * ka[0][c];
      }
      eval_aug(u, ua, f1y, f1a, r2);
      nfe++;
#pragma unroll
      for (int q = 0; q < Q; ++q) my_k[(size_t)(kSlots ? kAdjStages - 1 : 0) * WARPS * P + q] = r2[q];
#pragma unroll
      for (int k = 0; k < kAdjNS; ++k) sc[k] = 0.f;
      if (ph == 0) {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          ky[0][c] = f1y[c];
          ka[0][c] = f1a[c];
          scy[c] = atol32 + fabsf(y[c]) * rtol32;
          sca[c] = atol32 + fabsf(a[c]) * rtol32;
          if (valid) {
            const float r0 = y[c] / scy[c], r1 = a[c] / sca[c], r2y = ky[0][c] / scy[c], r3 = ka[0][c] / sca[c];
            sc[0] += r0 * r0; sc[1] += r1 * r1; sc[2] += r2y * r2y; sc[3] += r3 * r3;
            if (!isfinite(y[c]) || !isfinite(a[c])) sc[4] += 1.f;
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          const float ry = (f1y[c] - ky[0][c]) / scy[c], ra = (f1a[c] - ka[0][c]) / sca[c];
          if (valid) { sc[0] += ry * ry; sc[1] += ra * ra; }
        }
      }
      grid_reduce(1, sc, 0.f);
      if (ph == 0) {
        for (int n = tid; n < P; n += NT) s_k1[n] = s_tot[n];
        if (s_tot[kAdjNV * P + 4] > 0.f) status |= GODE_ST_NONFINITE;
        if (p.o.first_step > 0.0) {
          dt = p.o.first_step;
          break;
        }
        const float d0t = theta_norm([&](int n) { return s_th[n] / (atol32 + fabsf(s_th[n]) * rtol32); });
        const float d1t = theta_norm([&](int n) { return s_k1[n] / (atol32 + fabsf(s_th[n]) * rtol32); });
        d0 = fmaxf(fmaxf(sqrtf(s_tot[kAdjNV * P + 0] / n_elem), sqrtf(s_tot[kAdjNV * P + 1] / n_elem)), d0t);
        d1 = fmaxf(fmaxf(sqrtf(s_tot[kAdjNV * P + 2] / n_elem), sqrtf(s_tot[kAdjNV * P + 3] / n_elem)), d1t);
        h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : 0.01f * d0 / d1;
      } else {
        const float d2t = theta_norm([&](int n) { return (s_tot[n] - s_k1[n]) / (atol32 + fabsf(s_th[n]) * rtol32); });
        const float d2 =
            fmaxf(fmaxf(sqrtf(s_tot[kAdjNV * P + 0] / n_elem), sqrtf(s_tot[kAdjNV * P + 1] / n_elem)), d2t) / h0;
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
        else h1 = powf(0.01f / fmaxf(d1, d2), 0.2f);
        dt = (double)fminf(100.f * h0, h1);
      }
    }
    dt0 = dt;
    t0 = s_begin;

    // ---- rk_common.py::_adaptive_step until the step that reaches the end of the interval is accepted --------------------
    int n_steps = 0;
    bool done = false;
    while (!done && status == 0) {
      if (n_steps >= p.o.max_num_steps) { status |= GODE_ST_MAX_STEPS; break; }
      if (!(t0 + dt > t0)) { status |= GODE_ST_DT_UNDERFLOW; break; }
      const double t1 = t0 + dt;
      const float dt32 = (float)dt;
      const bool fin = !(s_end > t1);
      if constexpr (!kSlots) {
#pragma unroll
        for (int v = 1; v < kAdjNV; ++v) {
#pragma unroll
          for (int q = 0; q < Q; ++q) my_k[(size_t)v * WARPS * P + q] = 0.f;
        }
      }
      GADJ_TICK(tc_rest);
      float u[S::DL], ua[S::DL];
      // One copy of eval_aug in the instruction stream: the stage loop is NOT unrolled; the parts that index the stage
      // derivatives (registers) are dispatched on the stage number.
      auto stage_input = [&](auto ST) {
        constexpr int st = decltype(ST)::value;
#pragma unroll
        for (int c = 0; c < S::DL; ++c) {
          float sy = ky[0][c] * (kBeta[st][0] * dt32), sa = ka[0][c] * (kBeta[st][0] * dt32);
#pragma unroll
          for (int j = 1; j <= st; ++j) {
            sy = fmaf(ky[j][c], kBeta[st][j] * dt32, sy);
            sa = fmaf(ka[j][c], kBeta[st][j] * dt32, sa);
          }
          u[c] = y[c] + sy;
          ua[c] = a[c] + sa;
        }
      };
#pragma unroll 1
      for (int st = 0; st < 6; ++st) {
        switch (st) {
          case 0: stage_input(std::integral_constant<int, 0>{}); break;
          case 1: stage_input(std::integral_constant<int, 1>{}); break;
          case 2: stage_input(std::integral_constant<int, 2>{}); break;
          case 3: stage_input(std::integral_constant<int, 3>{}); break;
          case 4: stage_input(std::integral_constant<int, 4>{}); break;
          default: stage_input(std::integral_constant<int, 5>{}); break;
        }
        float kny[S::DL], kna[S::DL];
        eval_aug(u, ua, kny, kna, r2);
        // tableau weights of this stage's theta-part as compile-time constants (a dynamically indexed table would be loads)
        float ws = 0.f, we = 0.f, wm = 0.f;
        auto stage_store = [&](auto ST) {
          constexpr int j = decltype(ST)::value + 1;
#pragma unroll
          for (int c = 0; c < S::DL; ++c) { ky[j][c] = kny[c]; ka[j][c] = kna[c]; }
          ws = j <= 5 ? kBeta[5][j <= 5 ? j : 0] * dt32 : 0.f;
          we = kCErr[j] * dt32;
          wm = kCMid[j] * dt32;
        };
        switch (st) {
          case 0: stage_store(std::integral_constant<int, 0>{}); break;
          case 1: stage_store(std::integral_constant<int, 1>{}); break;
          case 2: stage_store(std::integral_constant<int, 2>{}); break;
          case 3: stage_store(std::integral_constant<int, 3>{}); break;
          case 4: stage_store(std::integral_constant<int, 4>{}); break;
          default: stage_store(std::integral_constant<int, 5>{}); break;
        }
        if constexpr (kSlots) {
          // the stage's theta-part (this lane's 17 of the warp's 544 values) into its slot; combined when the row is written
#pragma unroll
          for (int q = 0; q < Q; ++q) my_k[(size_t)st * WARPS * P + q] = r2[q];
        } else {
          // k_j = this stage's derivative (0-based j = st + 1); k_1 carries zero weight in all three combinations.  All loads
          // first, then the FMAs, then the stores: written as one += per element the chain was serialised
          const int j = st + 1;
          float t1[Q], t2[Q], t3[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            t1[q] = my_k[(size_t)1 * WARPS * P + q];
            t2[q] = my_k[(size_t)2 * WARPS * P + q];
            t3[q] = my_k[(size_t)3 * WARPS * P + q];
          }
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            t1[q] = fmaf(ws, r2[q], t1[q]);
            t2[q] = fmaf(we, r2[q], t2[q]);
            t3[q] = fmaf(wm, r2[q], t3[q]);
          }
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            my_k[(size_t)1 * WARPS * P + q] = t1[q];
            my_k[(size_t)2 * WARPS * P + q] = t2[q];
            my_k[(size_t)3 * WARPS * P + q] = t3[q];
          }
          if (j == 6) {
#pragma unroll
            for (int q = 0; q < Q; ++q) my_k[q] = r2[q];
          }
        }
      }
      nfe += 6;
      // u / ua are y1 / a1 (FSAL), k[6] their derivatives
#pragma unroll
      for (int k = 0; k < kAdjNS; ++k) sc[k] = 0.f;
#pragma unroll
      for (int c = 0; c < S::DL; ++c) {
        float ey = ky[0][c] * (dt32 * kCErr[0]), ea = ka[0][c] * (dt32 * kCErr[0]);
#pragma unroll
        for (int j = 2; j < 7; ++j) {
          ey = fmaf(ky[j][c], dt32 * kCErr[j], ey);
          ea = fmaf(ka[j][c], dt32 * kCErr[j], ea);
        }
        const float ry = ey / (atol32 + rtol32 * fmaxf(fabsf(y[c]), fabsf(u[c])));
        const float ra = ea / (atol32 + rtol32 * fmaxf(fabsf(a[c]), fabsf(ua[c])));
        if (valid) {
          sc[0] += ry * ry; sc[1] += ra * ra;
          if (!isfinite(u[c]) || !isfinite(ua[c])) sc[4] += 1.f;
        }
      }
      GADJ_TICK(tc_stage);
      grid_reduce(fin ? 4 : 3, sc, dt32);
      GADJ_TICK(tc_red);
      const float* tK7 = s_tot;
      const float* tSol = s_tot + P;
      const float* tErr = s_tot + 2 * P;
      const float* tMid = s_tot + 3 * P;
      const float ert = theta_norm([&](int n) {
        const float th1 = s_th[n] + fmaf(s_k1[n], kBeta[5][0] * dt32, tSol[n]);
        const float e = fmaf(s_k1[n], dt32 * kCErr[0], tErr[n]);
        return e / (atol32 + rtol32 * fmaxf(fabsf(s_th[n]), fabsf(th1)));
      });
      const float er = fmaxf(fmaxf(sqrtf(s_tot[kAdjNV * P + 0] / n_elem), sqrtf(s_tot[kAdjNV * P + 1] / n_elem)), ert);
      if (s_tot[kAdjNV * P + 4] > 0.f) status |= GODE_ST_NONFINITE;
      bool accept = er <= 1.f;
      if (dt > p.o.max_step) accept = false;
      if (dt <= p.o.min_step) accept = true;
      if (logger && n_att < p.o.log_capacity && p.att_dt) {
        p.att_dt[n_att] = dt; p.att_er[n_att] = er; p.att_acc[n_att] = accept ? 1 : 0;
      }
      if (accept) {
        ++n_acc;
        if (fin) {
          // dense output at the end of the interval (interp.py), then adjoint.py:579-580: y <- stored forward value,
          // a += upstream gradient of that output time
          const float x = (float)((s_end - t0) / (t1 - t0));
          float gy[S::DL];
#pragma unroll
          for (int c = 0; c < S::DL; ++c) {
            float m = ka[0][c] * (dt32 * kCMid[0]);
#pragma unroll
            for (int j = 2; j < 7; ++j) m = fmaf(ka[j][c], dt32 * kCMid[j], m);
            a[c] = interp_at(a[c], ua[c], a[c] + m, ka[0][c], ka[6][c], dt32, x);
            gy[c] = 0.f;
            y[c] = 0.f;
          }
          if (valid) {
            load_frag<S::DL>(p.traj + adj_toff(p.layout, i - 1, b, p.B, p.T, D) + l * S::DL, y);
            load_frag<S::DL>(p.grad_traj + adj_toff(p.layout, i - 1, b, p.B, p.T, D) + l * S::DL, gy);
          }
#pragma unroll
          for (int c = 0; c < S::DL; ++c) a[c] += gy[c];
          for (int n = tid; n < P; n += NT) {
            const float th0 = s_th[n], k1 = s_k1[n];
            const float th1 = th0 + fmaf(k1, kBeta[5][0] * dt32, tSol[n]);
            const float thm = th0 + fmaf(k1, dt32 * kCMid[0], tMid[n]);
            s_th[n] = interp_at(th0, th1, thm, k1, tK7[n], dt32, x);
          }
          done = true;
        } else {
#pragma unroll
          for (int c = 0; c < S::DL; ++c) { y[c] = u[c]; a[c] = ua[c]; ky[0][c] = ky[6][c]; ka[0][c] = ka[6][c]; }
          for (int n = tid; n < P; n += NT) {
            s_th[n] = s_th[n] + fmaf(s_k1[n], kBeta[5][0] * dt32, tSol[n]);
            s_k1[n] = tK7[n];
          }
          t0 = t1;
        }
      }
      dt = optimal_step(dt, er, p.o);
      dt = fmin(fmax(dt, p.o.min_step), p.o.max_step);
      ++n_att;
      ++n_steps;
    }
  }

  const float poison = status != 0 ? __int_as_float(0x7fc00000) : 0.f;
  if (valid) {
#pragma unroll
    for (int c = 0; c < S::DL; ++c) a[c] += poison;
    store_frag<S::DL>(p.grad_y0 + (size_t)b * D + l * S::DL, a);
  }
  if (blockIdx.x == 0) {
    for (int n = tid; n < P; n += NT) {
      int ts;
      p.grad_params[A::canon(n, ts)] = s_th[n] + poison;
    }
  }
#ifdef GODE_ADJ_TIMING
  if (logger) printf("adj timing: attempts %d  stages(other) %lld  reduce %lld  rest %lld | fwd %lld vjp %lld theta %lld cycles/attempt\n", n_att, tc_stage / max(n_att, 1), tc_red / max(n_att, 1), tc_rest / max(n_att, 1), tc_f / max(n_att, 1), tc_v / max(n_att, 1), tc_t / max(n_att, 1));
#endif
  if (logger && status != 0 && p.mailbox) *reinterpret_cast<volatile int32_t*>(p.mailbox) = status;
  if (logger && p.log) {
    p.log->status = status;
    p.log->n_attempts = n_att;
    p.log->n_accepted = n_acc;
    p.log->nfe = nfe;
    p.log->dt0 = dt0;
    p.log->t_final = t0;
  }
  if (logger) ss.finish(p.gs);   // unconditionally (the ODE-RNN frames pass no log): the next launch continues from these tags
}

// ------------------------------------------------------------------------------------------------------------
template <int D, int H, int L, int WARPS>
static int adj_grid(int B) {
  const int per_cta = WARPS * Shape<D, H, L>::G;
  return (B + per_cta - 1) / per_cta;
}

size_t dopri5_small_adjoint_workspace_bytes(int B, int D, int H) {
  (void)D; (void)H;
  using A = AdjLayout<16, 16, 8, kAdjWarps>;
  const int grid = adj_grid<16, 16, 8, kAdjWarps>(B);
  (void)grid;
  return (size_t)GODE_SYNC_REGION_BYTES;   // rows and totals of the per-attempt reduction live in the persistent region
}

template <int L, int MINB>
static int launch_adj(Dp5AdjArgs& a, void* workspace, size_t ws_bytes, cudaStream_t st) {
  constexpr int WARPS = kAdjWarps;
  using A = AdjLayout<16, 16, L, WARPS>;
  auto kern = dopri5_adjoint_bwd_kernel<16, 16, L, WARPS, MINB>;
  const size_t smem = sizeof(float) * A::kSmemFloats;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(1000 + (int)e);
  const int grid = adj_grid<16, 16, L, WARPS>(a.B);
  static int limit_cache = 0;
  const int cap = coop_limit(kern, WARPS * 32, smem, limit_cache);
  if (cap <= 0 || grid > cap || grid > kSyncMaxGrid) return GODE_ERR_COOP;
  if ((size_t)(grid + 1) * A::VT * sizeof(unsigned long long) > kSyncRowBytes) return GODE_ERR_COOP;
  char* base = reinterpret_cast<char*>(workspace);
  grid_sync_bind(a.gs, base);
  void* args[] = {(void*)&a};
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(WARPS * 32), args, smem, st);
  if (e != cudaSuccess) return -(1000 + (int)e);
  return launch_status();
}

int dopri5_small_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                             const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                             const GodeAdaptiveOpts* opts, int param_mask, float* grad_y0, float* grad_params,
                             GodeStepLog* log, double* att_dt, float* att_er, uint8_t* att_acc, void* workspace, size_t ws_bytes,
                             cudaStream_t st) {
  if (T > kAdjMaxT) return GODE_ERR_T_TOO_LONG;
  if (!(D == 16 && H == 16)) return GODE_ERR_SHAPE;
  if (ws_bytes < dopri5_small_adjoint_workspace_bytes(B, D, H)) return GODE_ERR_WORKSPACE;
  Dp5AdjArgs a{};
  a.traj = traj; a.grad_traj = grad_traj; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
  a.grad_y0 = grad_y0; a.grad_params = grad_params; a.log = log; a.mailbox = status_mailbox(); a.att_dt = att_dt; a.att_er = att_er; a.att_acc = att_acc;
  a.o = *opts; a.B = B; a.T = T; a.layout = layout; a.param_mask = param_mask;
  for (int i = 0; i < T; ++i) a.t[i] = t_host[i];
  // 8 lanes per trajectory (32 per CTA) has the shortest stage chain and wins while its CTAs get an SM each; beyond that
  // (B > 4736) 4 lanes (64 per CTA, one CTA per SM up to 9472) is faster: 0.64 vs 0.78 ms at B = 8192, 0.61 vs 0.53 ms at
  // B = 4096 (scripts/dp5_adj_lanes.py).  Developer switch GODE_ADJ_LANES = '8' | '4' forces the first choice.
  const char* force = getenv("GODE_ADJ_LANES");
  const bool first8 = force ? force[0] == '8' : B <= 32 * sm_count();
  // (one CTA per SM is all the 8-lane mapping is asked for — B <= 32 * SMs — so it is compiled for 255 registers; the
  // two-per-SM, 128-register build it used to be spilled around the reduction)
  int rc = first8 ? launch_adj<8, 1>(a, workspace, ws_bytes, st) : GODE_ERR_COOP;
  if (rc != GODE_ERR_COOP) return rc;
  rc = launch_adj<4, 1>(a, workspace, ws_bytes, st);
  if (rc != GODE_ERR_COOP) return rc;
  return launch_adj<2, 1>(a, workspace, ws_bytes, st);  // 128 trajectories per CTA: holds 18 944
}

}  // namespace gode
