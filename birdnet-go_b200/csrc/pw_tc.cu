// pw_tc.cu — the dense 1x1 ("pointwise") convolutions of the BirdNET v2.4 conv stack on the 5th-gen
// tensor cores: tcgen05.mma (kind::f16) with fp32 accumulators in TMEM, sm_100a only.
//
// Replaces the CONV_2D 1x1 / FULLY_CONNECTED ops the reference executes inside TFLite-XNNPACK
// (/root/reference/internal/inference/tflite/classifier.go:107); 85 % of the model's MACs.
//
// Numerics: plain fp16/bf16 operands miss the 1e-3 sigmoid parity bar by >3x, so
// both operands are split x = hi + lo (two fp16 values, ~22 significant bits) and every K-step issues
// three MMAs  D += Ahi*Bhi + Alo*Bhi + Ahi*Blo  into the same fp32 TMEM accumulator.
//
// Structure (one persistent CTA per SM, 10 warps, warp-specialised, mbarrier pipelines):
//   warps 0-3  epilogue   : tcgen05.ld 128 lanes x 16 columns -> bias / SiLU / ReLU / residual -> st.global
//   warp  4    MMA issuer : one lane issues tcgen05.mma; tcgen05.commit frees smem stages / publishes TMEM
//   warp  5    B loader   : weights were pre-split and pre-swizzled at load time into the exact shared-memory
//                           image (128B-swizzle, K-major); one cp.async.bulk per stage brings them in
//   warps 6-9  A producer : activations are fp32 in HBM/L2 -> coalesced 32 B loads -> optional SE gate or
//                           affine+ReLU -> hi/lo fp16 split -> 16 B st.shared into the swizzled K-major tile
//   accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the
//   main loop of tile t+1.
#include <cuda_fp16.h>

#include "kernels.h"
#include "pw_tc.h"

namespace bnb {

namespace {

constexpr int kBM = 128;            // rows per tile = TMEM lanes = UMMA M
constexpr int kBK = 64;             // K per stage: 64 fp16 = one 128-byte swizzle row
constexpr int kThreads = 320;
constexpr int kEpiWarps = 4, kMmaWarp = 4, kLoadWarp = 5, kProdWarp0 = 6, kProdThreads = 128;
constexpr int kAccCols = 256;       // TMEM columns per accumulator buffer
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct SmemLayout {
  uint32_t a_hi, a_lo, b_hi, b_lo;   // byte offsets of stage 0 from the 1024-aligned base
  uint32_t stage_bytes;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_SILU) return __fdividef(v, 1.0f + __expf(-v));
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__global__ void __launch_bounds__(kThreads, 1)
pw_tc_kernel(const PwTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_bytes = kBM * 128u;                     // one fp16 [128][64] slab
  const uint32_t b_bytes = (uint32_t)a.bn_max * 128u;      // one fp16 [bn_max][64] slab
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const uint32_t bars = base + (uint32_t)a.stages * stage_bytes;   // full[S], empty[S], tfull[2], tempty[2], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * a.stages + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * a.stages + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (size_t)a.stages * stage_bytes + 8u * (2 * a.stages + 4));

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), kProdThreads + 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * kAccCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (a.M + kBM - 1) / kBM;
  const int total_tiles = m_tiles * a.n_tiles;
  const int k_stages = (a.K + kBK - 1) / kBK;

  if (warp >= kProdWarp0) {
    // ============================== A producers ==============================================
    const int pt = threadIdx.x - kProdWarp0 * 32;           // 0..127
    uint32_t it = 0;                                        // global stage counter
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / a.n_tiles) * kBM;
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        uint8_t* hi = base_ptr + (size_t)s * stage_bytes;
        uint8_t* lo = hi + a_bytes;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          const int q = q8 * kProdThreads + pt;             // chunk id: row = q / 8, 16-byte chunk c = q % 8
          const int r = q >> 3, c = q & 7;
          const int m = m0 + r, k = ks * kBK + c * 8;
          if (k >= a.k_pad) continue;                        // columns no MMA of this stage reads
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
          if (m < a.M && k < a.K) {
            const float* src;
            if (a.a_mode == A_PLAIN) src = a.A + (size_t)m * a.K + k;
            else {
              const int bidx = m / a.out_w, wo = m - bidx * a.out_w;
              const int a_seg = a.kw * a.cin, seg = k / a_seg, j = k - seg * a_seg;
              src = a.A + ((size_t)bidx * (a.K / a_seg) * a.in_w + wo) * a.cin + (size_t)seg * a.in_w * a.cin + j;
            }
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(src));
            v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
            if (k + 4 < a.K) {                               // K is a multiple of 4, not necessarily of 8
              const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + 4));
              v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
            }
            if (a.gate) {
              const float* g = a.gate + (size_t)(m / a.rows_per_chunk) * a.K + k;
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(g));
              v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
              if (k + 4 < a.K) {
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
                v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
              }
            }
            if (a.a_mul) {
              const int ch = k % a.a_ch;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (k + i < a.K) v[i] = fmaxf(fmaf(v[i], __ldg(a.a_mul + ch + i), __ldg(a.a_add + ch + i)), 0.f);
            }
          }
          uint32_t ph_[4], pl_[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
            const __half l0 = __float2half_rn(v[2 * i] - __half2float(h0)), l1 = __float2half_rn(v[2 * i + 1] - __half2float(h1));
            ph_[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            pl_[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
          const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(hi + off) = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
          *reinterpret_cast<uint4*>(lo + off) = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
        }
        fence_proxy_async();
        mbar_arrive(full_bar(s));
      }
    }
  } else if (warp == kLoadWarp) {
    // ============================== B loader ===================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_tiles;
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t dst = base + (uint32_t)s * stage_bytes + 2 * a_bytes;
          const uint8_t* src = a.Wimg + ((size_t)nt * k_stages + ks) * (size_t)(2 * b_bytes);
          mbar_arrive_expect_tx(full_bar(s), 2 * b_bytes);
          bulk_g2s(dst, src, 2 * b_bytes, full_bar(s));
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ================================================
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int nt = tile % a.n_tiles;
        const int n0 = nt * a.bn_max;
        const int bn = min(a.bn_max, a.n_pad - n0);
        const uint32_t idesc = make_idesc((uint32_t)bn);
        const int buf = tcount & 1;
        mbar_wait(tempty_bar(buf), ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint64_t d_ahi = make_desc(sa), d_alo = make_desc(sa + a_bytes);
          const uint64_t d_bhi = make_desc(sa + 2 * a_bytes), d_blo = make_desc(sa + 2 * a_bytes + b_bytes);
          const int kk_n = min(kBK, a.k_pad - ks * kBK) / 16;
          for (int kk = 0; kk < kk_n; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);        // +32 bytes (>>4) inside the swizzle row
            umma(d_tmem, d_ahi + adv, d_bhi + adv, idesc, (ks | kk) != 0);
            umma(d_tmem, d_alo + adv, d_bhi + adv, idesc, 1);
            umma(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1);
          }
          umma_commit(empty_bar(s));                        // smem stage reusable once these MMAs retire
        }
        umma_commit(tfull_bar(buf));                        // accumulator complete
      }
    }
  } else {
    // ============================== epilogue (warps 0-3) =======================================
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / a.n_tiles, nt = tile % a.n_tiles;
      const int n0 = nt * a.bn_max;
      const int bn = min(a.bn_max, a.n_pad - n0);
      const int buf = tcount & 1;
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      tc_fence_after();
      const int m = mt * kBM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * kAccCols;
      float* crow = a.C + (size_t)m * a.N;
      const float* rrow = a.residual ? a.residual + (size_t)m * a.N : nullptr;
      for (int c0 = 0; c0 < bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (m < a.M) {
          const int n = n0 + c0;
          if (n + 16 <= a.N) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bz = __ldg(reinterpret_cast<const float4*>(a.bias + n) + j4);
              float4 o;
              o.x = act_apply(__uint_as_float(r[4 * j4 + 0]) + bz.x, a.act);
              o.y = act_apply(__uint_as_float(r[4 * j4 + 1]) + bz.y, a.act);
              o.z = act_apply(__uint_as_float(r[4 * j4 + 2]) + bz.z, a.act);
              o.w = act_apply(__uint_as_float(r[4 * j4 + 3]) + bz.w, a.act);
              if (rrow) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(rrow + n) + j4);
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
              }
              if (a.c_vec4) *(reinterpret_cast<float4*>(crow + n) + j4) = o;
              else { crow[n + 4 * j4] = o.x; crow[n + 4 * j4 + 1] = o.y; crow[n + 4 * j4 + 2] = o.z; crow[n + 4 * j4 + 3] = o.w; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n + j < a.N) {
                float o = act_apply(__uint_as_float(r[j]) + __ldg(a.bias + n + j), a.act);
                if (rrow) o += __ldg(rrow + n + j);
                crow[n + j] = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kAccCols));
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// Host side: layer preparation (weight split + swizzled image) and launch
// ---------------------------------------------------------------------------------------------------
PwTcLayer pw_tc_prepare(const float* w, int N, int K, std::vector<uint8_t>* image) {
  PwTcLayer L;
  L.N = N; L.K = K;
  L.n_pad = (N + 15) / 16 * 16;
  L.k_pad = (K + 15) / 16 * 16;
  L.n_tiles = (L.n_pad + 255) / 256;
  L.bn_max = ((L.n_pad + L.n_tiles - 1) / L.n_tiles + 15) / 16 * 16;
  L.n_tiles = (L.n_pad + L.bn_max - 1) / L.bn_max;
  L.k_stages = (K + kBK - 1) / kBK;
  const size_t slab = (size_t)L.bn_max * 128;              // one fp16 [bn_max][64] swizzled slab
  image->assign((size_t)L.n_tiles * L.k_stages * 2 * slab, 0);
  for (int nt = 0; nt < L.n_tiles; ++nt)
    for (int ks = 0; ks < L.k_stages; ++ks) {
      uint8_t* hi = image->data() + ((size_t)nt * L.k_stages + ks) * 2 * slab;
      uint8_t* lo = hi + slab;
      for (int r = 0; r < L.bn_max; ++r) {
        const int n = nt * L.bn_max + r;
        if (n >= N) continue;
        for (int kc = 0; kc < kBK; ++kc) {
          const int k = ks * kBK + kc;
          if (k >= K) break;
          const float x = w[(size_t)n * K + k];
          const __half h = __float2half_rn(x);
          const __half l = __float2half_rn(x - __half2float(h));
          const int c = kc >> 3, e = kc & 7;
          const size_t off = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((c ^ (r & 7)) << 4) + (size_t)e * 2;
          const unsigned short hb = __half_as_ushort(h), lb = __half_as_ushort(l);
          memcpy(hi + off, &hb, 2); memcpy(lo + off, &lb, 2);
        }
      }
    }
  // pipeline depth from the shared-memory budget
  const size_t stage = 2 * (size_t)kBM * 128 + 2 * slab;
  int stages = (int)((220 * 1024) / stage);
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  L.stages = stages;
  L.smem_bytes = (size_t)stages * stage + 1024 /*alignment*/ + 8 * (2 * stages + 4) + 16;
  return L;
}

void launch_pw_tc(const PwTcLayer& L, const PwArgs& p, const uint8_t* d_image, cudaStream_t s, LaunchCounter& lc) {
  static size_t max_set = 0;
  if (L.smem_bytes > 227 * 1024) throw std::runtime_error("pw_tc: shared memory budget exceeded");
  if (L.smem_bytes > max_set) {
    BNB_CUDA(cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_bytes));
    max_set = L.smem_bytes;
  }
  PwTcArgs a{};
  a.A = p.A; a.Wimg = d_image; a.bias = p.bias; a.C = p.C; a.residual = p.residual; a.gate = p.gate;
  a.a_mul = p.a_mul; a.a_add = p.a_add; a.a_ch = p.a_ch;
  a.M = p.M; a.N = p.N; a.K = p.K; a.rows_per_chunk = p.rows_per_chunk; a.act = p.act; a.a_mode = p.a_mode;
  a.in_w = p.in_w; a.out_w = p.out_w; a.cin = p.cin; a.kw = p.kw;
  a.n_pad = L.n_pad; a.k_pad = L.k_pad; a.n_tiles = L.n_tiles; a.bn_max = L.bn_max; a.stages = L.stages;
  a.c_vec4 = (p.N % 4 == 0) ? 1 : 0;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int tiles = m_tiles * L.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  pw_tc_kernel<<<grid, kThreads, L.smem_bytes, s>>>(a);
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
