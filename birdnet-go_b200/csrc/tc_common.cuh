// tc_common.cuh — tcgen05 / TMA / mbarrier primitives shared by the tensor-core kernels (pw_tc.cu, mbconv_tc.cu).
// Raw PTX for sm_100a; SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace bnb {
namespace tc {

constexpr int kBM = 128;            // rows per tile = TMEM lanes = UMMA M
constexpr int kBK = 64;             // K per stage: 64 fp16 = one 128-byte swizzle row
constexpr uint32_t kSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a broken pipeline traps (cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > kSpinLimit) __trap();
  }
}
// Same, for roles that wait long and are not latency critical (producers waiting for a free slot, consumers waiting
// for data far ahead): back off between polls so the spinning warp does not steal issue slots from the math warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __nanosleep(100);
    if (spin > (kSpinLimit >> 4)) __trap();
  }
}
// One lane of a CONVERGED warp.  Guarding the tcgen05 / bulk-copy issue with this (and not with `lane == 0`) matters: ptxas
// knows an elect.sync branch holds exactly one thread and emits the UTCHMMA / UBLKCP with its uniform-register operands
// directly; behind `lane == 0` every such instruction was wrapped in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop
// (~10 dependent instructions per MMA on an SMSP shared with the epilogue warps: 200 cycles per MMA, profiles/r02 SASS notes).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Per-thread asynchronous 16-byte copies (LDGSTS, L2-only) and their completion hooked to an mbarrier: the arrival fires when all
// earlier cp.async of THIS thread have landed (.noinc: the barrier's expected count must already include this thread).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// 256-bit global accesses (sm_100: LDG / STG .ENL2.256): two adjacent 16-byte pieces of a plane row as ONE full 32-byte sector
// instead of two half-sector accesses (the epilogue of pw2.cu scatters / gathers such piece pairs)
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ldg256_cg(const void* p, uint4& a, uint4& b) {      // L2-coherent (see PDL rules in common.cuh)
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p) : "memory");
}

// one 2-D TMA tile load (tensor map in kernel-parameter space): coordinates (c0 = innermost = k, c1 = row)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 format, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 4-D TMA tile load: coordinates innermost first (c, w, h, b); out-of-range (also negative) coordinates are zero-filled
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// SiLU of two values with ONE reciprocal: 1/a = b * rcp(a*b).  x is clamped at -80 so that a, b = 1 + exp(-x) stay
// finite (a*b may overflow to +inf -> rcp = 0 -> both results -0, the correct limit).  MUFU-bound epilogues: 1.5 instead of 2 per value.
__device__ __forceinline__ void silu2(float& x, float& y) {
  float ex, ey, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * fmaxf(x, -80.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(-1.4426950408889634f * fmaxf(y, -80.f)));
  const float a = 1.0f + ex, b = 1.0f + ey;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
  x = x * (r * b); y = y * (r * a);
}

// fp32 pair -> packed fp16 hi and lo words (Veltkamp split: x = hi + lo exactly, hi has 11 significant bits)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float t0 = __fmul_rn(x0, 8193.0f), t1 = __fmul_rn(x1, 8193.0f);
  const float h0 = __fsub_rn(t0, __fsub_rn(t0, x0)), h1 = __fsub_rn(t1, __fsub_rn(t1, x1));
  const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn(__fsub_rn(x0, h0), __fsub_rn(x1, h1));
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}


// ---------------------------------------------------------------------------------------------------------------
// round-2 additions: shared-memory descriptors for the three K-major swizzle modes, narrow TMEM loads at arbitrary
// column offsets, fp16 hi/lo "plane" helpers, four-way shared-reciprocal SiLU
// ---------------------------------------------------------------------------------------------------------------
// K-major swizzled operand tile with `row_bytes` per row (128 / 64 / 32 = SWIZZLE_128B / 64B / 32B); 8-row atoms are
// contiguous (stride byte offset = 8 * row_bytes).  layout_type: 2 / 4 / 6 (sm_100 encoding).
__device__ __forceinline__ uint64_t make_desc_rb(uint32_t saddr, uint32_t row_bytes) {
  const uint64_t lt = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= lt << 61;
  return d;
}
// byte offset of 16-byte chunk `chunk` of row `r` inside a swizzled plane whose base is aligned to 8 * row_bytes
__host__ __device__ __forceinline__ uint32_t swz_off(uint32_t r, uint32_t chunk, uint32_t row_bytes) {
  const uint32_t a = r * row_bytes + chunk * 16u;
  const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  return a ^ (((a >> 7) & mask) << 4);
}

#define BNB_TMEM_LD_FN(NAME, XN, OUTS, ...)                                                                     \
  __device__ __forceinline__ void NAME(uint32_t taddr, uint32_t* r) {                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b." XN ".b32 {" __VA_ARGS__ "}, [%" #OUTS "];" : BNB_TMEM_OUT_##OUTS : "r"(taddr)); \
  }
#define BNB_TMEM_OUT_1 "=r"(r[0])
#define BNB_TMEM_OUT_2 "=r"(r[0]), "=r"(r[1])
#define BNB_TMEM_OUT_4 "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
#define BNB_TMEM_OUT_8 "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
BNB_TMEM_LD_FN(tmem_ld1, "x1", 1, "%0")
BNB_TMEM_LD_FN(tmem_ld2, "x2", 2, "%0, %1")
BNB_TMEM_LD_FN(tmem_ld4, "x4", 4, "%0, %1, %2, %3")
BNB_TMEM_LD_FN(tmem_ld8, "x8", 8, "%0, %1, %2, %3, %4, %5, %6, %7")
// N consecutive 32-bit columns starting at ANY column (pieces of 16 / 8 / 4 / 2 / 1)
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N >= 16) { tmem_ld16(taddr, r); tmem_ld_n<N - 16>(taddr + 16, r + 16); }
  else if constexpr (N >= 8) { tmem_ld8(taddr, r); tmem_ld_n<N - 8>(taddr + 8, r + 8); }
  else if constexpr (N >= 4) { tmem_ld4(taddr, r); tmem_ld_n<N - 4>(taddr + 4, r + 4); }
  else if constexpr (N >= 2) { tmem_ld2(taddr, r); tmem_ld_n<N - 2>(taddr + 2, r + 2); }
  else if constexpr (N == 1) { tmem_ld1(taddr, r); }
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 2-D / 4-D TMA loads are above; instruction descriptor with explicit M (128) and N
__device__ __forceinline__ uint32_t make_idesc_mn(uint32_t m, uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

// ---- fp16 hi/lo planes: x ~= hi + lo (22 significant bits), each plane a plain fp16 tensor ----------------------------
// two fp32 values -> packed hi pair and packed lo pair (one packed cvt each; the hi halves are widened back exactly)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ float2 join2(uint32_t hi, uint32_t lo) {
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return make_float2(h.x + l.x, h.y + l.y);
}

// SiLU of four values with ONE reciprocal.  The exponent argument is capped at 30 (x >= -20.8: below that
// silu(x) is within 2e-8 of the capped value), so every 1 + exp(-x) <= 2^30 + 1 and the product of four stays finite.
__device__ __forceinline__ float silu_e(float t) {       // 1 + exp(-x) from t = -x * log2(e)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(t, 30.f)));
  return 1.0f + e;
}
__device__ __forceinline__ float rcp_f(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
constexpr float kNegLog2e = -1.4426950408889634f;
__device__ __forceinline__ float ex2_f(float t) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t)); return e; }
// 2^t for t in [-30, 30] WITHOUT the MUFU pipe: round-to-nearest split t = n + f with the 1.5 * 2^23 constant (two FADDs), degree-6
// interpolating polynomial for 2^f on [-0.5, 0.5] (6 FFMAs, max relative error 1.0e-7 in fp32 Horner form: better than
// ex2.approx's 2 ulp), n added to the exponent field (shift + integer add).  The FA4 trick; measured NOT to pay here (mbconv2.cu,
// BNB_MB2_POLY): the epilogue is issue-bound, so 11 extra instructions cost more than the MUFU slot they free.
__device__ __forceinline__ float ex2_poly(float t) {
  const float magic = 12582912.0f;                       // 1.5 * 2^23
  const float r = t + magic;
  const float f = t - (r - magic);
  float p = 1.5461444697e-04f;
  p = fmaf(p, f, 1.3400428177e-03f); p = fmaf(p, f, 9.6180566785e-03f); p = fmaf(p, f, 5.5503272267e-02f);
  p = fmaf(p, f, 2.4022650922e-01f); p = fmaf(p, f, 6.9314720670e-01f); p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// Four SiLUs with ONE reciprocal (of the product of the four denominators) and packed fp32 multiplies / adds (sm_100 FMUL2 /
// FADD2 work on register pairs): 22 instructions instead of 30.  POLY of the four exponentials (which ones is fixed by the
// position in the group, never by data: results stay independent of batch composition) use ex2_poly instead of MUFU.EX2.
template <int POLY = 0>
__device__ __forceinline__ void silu4(float& x0, float& x1, float& x2, float& x3) {
  const float2 k = make_float2(kNegLog2e, kNegLog2e), one = make_float2(1.0f, 1.0f);
  const float2 t01 = __fmul2_rn(make_float2(x0, x1), k), t23 = __fmul2_rn(make_float2(x2, x3), k);
  float2 e01, e23;
  e01.x = ex2_f(fminf(t01.x, 30.f));
  e01.y = POLY >= 2 ? ex2_poly(fmaxf(fminf(t01.y, 30.f), -30.f)) : ex2_f(fminf(t01.y, 30.f));
  e23.x = ex2_f(fminf(t23.x, 30.f));
  e23.y = POLY >= 1 ? ex2_poly(fmaxf(fminf(t23.y, 30.f), -30.f)) : ex2_f(fminf(t23.y, 30.f));
  const float2 ab2 = __fadd2_rn(e01, one), cd2 = __fadd2_rn(e23, one);          // (a, b), (c, d) = 1 + exp(-x)
  const float ab = ab2.x * ab2.y, cd = cd2.x * cd2.y;
  const float r = rcp_f(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  const float2 s01 = __fmul2_rn(make_float2(ab2.y, ab2.x), make_float2(rab, rab));   // 1 / a = rab * b, 1 / b = rab * a
  const float2 s23 = __fmul2_rn(make_float2(cd2.y, cd2.x), make_float2(rcd, rcd));
  const float2 y01 = __fmul2_rn(make_float2(x0, x1), s01), y23 = __fmul2_rn(make_float2(x2, x3), s23);
  x0 = y01.x; x1 = y01.y; x2 = y23.x; x3 = y23.y;
}
__device__ __forceinline__ void silu2b(float& x0, float& x1) {
  const float a = silu_e(x0 * kNegLog2e), b = silu_e(x1 * kNegLog2e);
  const float r = rcp_f(a * b);
  x0 *= r * b; x1 *= r * a;
}
__device__ __forceinline__ void silu1b(float& x0) { x0 *= rcp_f(silu_e(x0 * kNegLog2e)); }
template <int N, int POLY = 0>
__device__ __forceinline__ void silu_n(float* v) {
#pragma unroll
  for (int i = 0; i + 4 <= N; i += 4) silu4<POLY>(v[i], v[i + 1], v[i + 2], v[i + 3]);
  if constexpr (N % 4 == 3) { silu2b(v[N - 3], v[N - 2]); silu1b(v[N - 1]); }
  if constexpr (N % 4 == 2) silu2b(v[N - 2], v[N - 1]);
  if constexpr (N % 4 == 1) silu1b(v[N - 1]);
}

}  // namespace tc
}  // namespace bnb
