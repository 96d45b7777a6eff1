// tc_common.cuh — thin inline-PTX layer over the 5th-generation tensor cores (tcgen05), TMEM, mbarrier and the
// bulk-copy (TMA) engine, for sm_100a.  No CUTLASS: only what the fused Runge–Kutta tile kernels need.
//
// Operand convention used by every kernel here: K-major operands in the canonical NO-SWIZZLE ("interleave")
// shared-memory layout of the UMMA descriptor: the tile is a grid of 8-row x 16-byte core matrices, each 128
// contiguous bytes; along M/N consecutive 8-row groups are SBO bytes apart, along K consecutive 16-byte chunks are LBO
// bytes apart.  We store a tile of R rows as [k_chunk][row][16 B]:  LBO = R*16, SBO = 128.  A thread that owns a row
// writes its 16-byte chunks with one STS.128 each; consecutive rows are consecutive 16-byte words => conflict-free.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gode {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (64-bit), SWIZZLE_NONE, K-major -----------------------------------------------
// bits [0,14) start address >> 4 | [16,30) leading-dim byte offset >> 4 | [32,46) stride-dim byte offset >> 4 |
// [46,48) descriptor version = 1 on sm_100 | [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---- instruction descriptor (32-bit) for kind::tf32 / kind::f16, fp32 accumulate, both operands K-major -------------
// [4,6) D format (1 = f32) | [7,10) A format | [10,13) B format (0 f16, 1 bf16, 2 tf32) | bit 15/16 A/B major (0 = K) |
// [17,23) N >> 3 | [24,29) M >> 4
enum : uint32_t { kFmtBF16 = 1, kFmtTF32 = 2 };
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- TMEM ---------------------------------------------------------------------------------------------------------
// one full warp allocates; the base address (lane 0, column c) lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch, bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 bit, 16 consecutive columns: thread i of warp w gets row 32*(w%4)+i, columns [c, c+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// same without the wait: issue several, then tmem_ld_wait() once (keeps more TMEM reads in flight)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- MMA: D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread ------------------------------------------------------
template <bool TF32>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// make the mbarrier observe completion of all MMAs issued so far by this thread (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
               : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}"
      ::"r"(smem_u32(mbar)), "r"(parity)
      : "memory");
}

// ---- TMA bulk copy global -> shared (1-D), completion on an mbarrier ------------------------------------------------
// size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

// ---- element packing ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace tc
}  // namespace gode

namespace gode {
namespace tc {
// two tanh per MUFU op on packed bf16 (sm_90+); the packed result is directly a BF16 MMA operand pair
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
}  // namespace tc
}  // namespace gode

namespace gode {
namespace tc {
// ---- TMEM stores: thread i of warp w writes lane 32*(w%4)+i, 8 consecutive 32-bit columns --------------------------------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MMA with the A operand in TMEM (16-bit elements packed two per 32-bit column, lane = row), B in shared memory ------
__device__ __forceinline__ void mma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- named barrier over a subset of the CTA's warps (ids 1..15; 0 is __syncthreads) -------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// one lane of a fully converged warp (warp-uniform control flow keeps MMA descriptors in uniform registers; a
// `threadIdx.x == 0` branch instead makes ptxas wrap every tcgen05.mma in an ELECT/R2UR waterfall loop, ~60 cycles each)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
}  // namespace tc
}  // namespace gode

namespace gode {
namespace tc {
// tanh of TWO values on the FMA pipe plus ONE MUFU op (instead of two MUFU.TANH): tanh(x) = 1 - 2 / (1 + e^(2x)) with
//   e^(2x) = 2^t, t = 2 log2(e) x, split by the magic-number trick into n = round(t) (left in the low mantissa bits of
//   t + 1.5*2^23) and f = t - n in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (7.5e-5 relative), 2^n by adding n
//   into the exponent field with one integer multiply-add;
//   the two reciprocals share one MUFU.RCP: r = 1 / (da db), 1/da = r db, 1/db = r da.
// Max abs error 3.7e-5 over [-30, 30] (checked against np.tanh), saturates cleanly (|t| clamped to 126; da db = inf -> r = 0).
// Arithmetic is packed (FFMA2/FMUL2/FADD2: two lanes per issue slot).  Used to take part of the tanh load of the wide
// forward off the 16-lane MUFU pipe, which bounds that kernel.
__device__ __forceinline__ float2 tanh_pair_fma(float a, float b) {
  const float kScale = 2.8853900817779268f, kMagic = 12582912.f;
  float2 t = __fmul2_rn(make_float2(a, b), make_float2(kScale, kScale));
  t.x = fminf(fmaxf(t.x, -126.f), 126.f);
  t.y = fminf(fmaxf(t.y, -126.f), 126.f);
  const float2 s = __fadd2_rn(t, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(s, make_float2(-kMagic, -kMagic));
  const float2 f = __fadd2_rn(t, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(f, make_float2(0.055171899f, 0.055171899f), make_float2(0.24261115f, 0.24261115f));
  p = __ffma2_rn(p, f, make_float2(0.69326097f, 0.69326097f));
  p = __ffma2_rn(p, f, make_float2(0.99992806f, 0.99992806f));
  const float ea = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(s.x) << 23));
  const float eb = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(s.y) << 23));
  const float2 d = __fadd2_rn(make_float2(ea, eb), make_float2(1.f, 1.f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d.x * d.y));
  const float2 q = __fmul2_rn(make_float2(d.y, d.x), make_float2(r, r));  // (1/da, 1/db)
  return __ffma2_rn(q, make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
}
// tanh of TWO values on the FMA pipe alone (no MUFU op): clamp to |x| <= 3.2, then the odd polynomial x P(x^2) with seven
// coefficients, minimax in RELATIVE error (linear program over a dense grid, saturation x > 3.2 included): 1.8e-3 relative at
// most — below the bf16 rounding (3.9e-3) the result goes through.  Twelve issue slots per pair (4 FMNMX, 2 FMUL2, 6 FFMA2).
// Lets the FMA pipe take a share of the tanh load that bounds the wide tensor-core forward (16 MUFU lanes per SM).
__device__ __forceinline__ float2 tanh_pair_poly(float a, float b) {
  const float c = 3.2f;
  const float2 x = make_float2(fminf(fmaxf(a, -c), c), fminf(fmaxf(b, -c), c));
  const float2 u = __fmul2_rn(x, x);
  float2 p = __ffma2_rn(u, make_float2(4.05731589e-06f, 4.05731589e-06f), make_float2(-1.53008630e-04f, -1.53008630e-04f));
  p = __ffma2_rn(p, u, make_float2(2.35090358e-03f, 2.35090358e-03f));
  p = __ffma2_rn(p, u, make_float2(-1.92475617e-02f, -1.92475617e-02f));
  p = __ffma2_rn(p, u, make_float2(9.43633094e-02f, 9.43633094e-02f));
  p = __ffma2_rn(p, u, make_float2(-3.13762426e-01f, -3.13762426e-01f));
  p = __ffma2_rn(p, u, make_float2(9.98174846e-01f, 9.98174846e-01f));
  return __fmul2_rn(p, x);
}
}  // namespace tc
}  // namespace gode
