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
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.h"
#include "pw_tc.h"
#include "tc_common.cuh"

namespace bnb {

namespace {

using namespace tc;

constexpr int kThreads = 576;         // 18 warps: 8 epilogue, 1 MMA issuer, 1 loader, 8 converters
constexpr int kEpiWarps = 8, kMmaWarp = 8, kLoadWarp = 9, kProdWarp0 = 10, kProdThreads = 256;
constexpr int kRowsPerPass = kProdThreads / 8, kPasses = 128 / kRowsPerPass;   // converter: 32 rows per pass, 4 passes
constexpr int kAccCols = 256;       // TMEM columns per accumulator buffer

#define BNB_TRACE(ev, iter) do { if (a.trace && blockIdx.x == 0 && (iter) < 64) a.trace[(ev) * 64 + (iter)] = clock64(); } while (0)

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_SILU) return __fdividef(v, 1.0f + __expf(-v));
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Shared-memory stage: [A: 32 KB][B hi: bn*128][B lo: bn*128].  The A buffer first receives the RAW fp32 rows
// (128 rows x 256 B, one cp.async.bulk per row, fully asynchronous => many stages of global latency in flight), is
// read into registers by the 4 producer warps, and is then overwritten IN PLACE by the fp16 hi (16 KB) and lo (16 KB)
// 128B-swizzled K-major tiles the MMA consumes.
__global__ void __launch_bounds__(kThreads, 1)
pw_tc_kernel(const PwTcArgs a, const __grid_constant__ CUtensorMap a_map) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_bytes = kBM * 128u;                     // one fp16 [128][64] slab (16 KB); raw fp32 uses 2 of them
  const uint32_t b_bytes = (uint32_t)a.bn * 128u;          // one fp16 [bn][64] slab
  const uint32_t stage_bytes = 2 * a_bytes + (a.b_res ? 0u : 2 * b_bytes);
  const uint32_t bres_bytes = a.b_res ? 2 * b_bytes : 0u;          // resident weight tile (hi | lo) after the stage ring
  const uint32_t bres = base + (uint32_t)a.stages * stage_bytes;
  const uint32_t bars = bres + bres_bytes + 8u * 4096u;   // [8 x 4 KB epilogue staging] rawfull[S], full[S], empty[S], tfull[2], tempty[2], tmem ptr
  auto raw_bar = [&](int s) { return bars + 8u * s; };
  auto full_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * a.stages + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (3 * a.stages + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (3 * a.stages + 2 + b); };
  const uint32_t bres_bar = bars + 8u * (3 * a.stages + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (size_t)a.stages * stage_bytes + bres_bytes + 8u * 4096u + 8u * (3 * a.stages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(raw_bar(s), 1); mbar_init(full_bar(s), kProdThreads + 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiWarps * 32); }
    mbar_init(bres_bar, 1);
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

  // N-stationary CTAs: CTA b keeps n-tile (b % n_tiles) for its whole life (bias strip and, for one-stage layers, the
  // weight tile are loaded once) and walks over m-tiles b / n_tiles, + gridDim.x / n_tiles, ...  (gridDim.x % n_tiles == 0)
  const int m_tiles = (a.M + kBM - 1) / kBM;
  const int nt_fix = blockIdx.x % a.n_tiles, mt_first = blockIdx.x / a.n_tiles, mt_step = gridDim.x / a.n_tiles;
  const int k_stages = (a.K + kBK - 1) / kBK;
  const int pitch = a.box_k * 4;                               // bytes per raw fp32 row in the landing buffer

  if (warp >= kProdWarp0) {
    // ============================== A converters (4 warps) ====================================
    const int pt = threadIdx.x - kProdWarp0 * 32;           // 0..255
    const int c = pt & 7, r0 = pt >> 3;                     // this thread: 16-byte chunk c of rows r0 + 32*q
    uint32_t it = 0;
    for (int mt = mt_first; mt < m_tiles; mt += mt_step) {
      const int m0 = mt * kBM;
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
        uint8_t* buf = base_ptr + (size_t)s * stage_bytes;
        const int k = ks * kBK + c * 8;
        const bool k_live = k < a.k_pad;                     // chunk read by some MMA of this stage
        float v[kPasses][8];
        // SE gate of each row's chunk: independent of the A tile, so the loads are issued BEFORE the wait on the
        // TMA data and their L2 latency hides behind it (they were the converter's critical path).
        float4 g0[kPasses], g1[kPasses];
        if (a.gate && k_live) {
#pragma unroll
          for (int q = 0; q < kPasses; ++q) {
            const int m = m0 + r0 + kRowsPerPass * q;
            g0[q] = make_float4(1.f, 1.f, 1.f, 1.f); g1[q] = g0[q];
            if (m < a.M && k < a.K) {
              const float* gp = a.gate + (size_t)(m / a.rows_per_chunk) * a.K + k;
              g0[q] = __ldg(reinterpret_cast<const float4*>(gp));
              if (k + 4 < a.K) g1[q] = __ldg(reinterpret_cast<const float4*>(gp + 4));
            }
          }
        }
        mbar_wait_relaxed(raw_bar(s), ph);
        if (pt == 0) BNB_TRACE(1, it);
        if (k_live) {
#pragma unroll
          for (int q = 0; q < kPasses; ++q) {
            const int r = r0 + kRowsPerPass * q, m = m0 + r;
            if (m < a.M && k < a.K) {
              const float4 x0 = *reinterpret_cast<const float4*>(buf + r * pitch + c * 32);
              float4 x1 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k + 4 < a.K) x1 = *reinterpret_cast<const float4*>(buf + r * pitch + c * 32 + 16);
              v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x0.z; v[q][3] = x0.w;
              v[q][4] = x1.x; v[q][5] = x1.y; v[q][6] = x1.z; v[q][7] = x1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[q][i] = 0.f;
            }
          }
          if (a.gate) {
#pragma unroll
            for (int q = 0; q < kPasses; ++q) {
              v[q][0] *= g0[q].x; v[q][1] *= g0[q].y; v[q][2] *= g0[q].z; v[q][3] *= g0[q].w;
              v[q][4] *= g1[q].x; v[q][5] *= g1[q].y; v[q][6] *= g1[q].z; v[q][7] *= g1[q].w;
            }
          }
        }
        // every converter has finished READING the raw rows before anyone overwrites them with the fp16 tiles
        asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");
        if (k_live) {
          uint8_t* hi = buf;
          uint8_t* lo = buf + a_bytes;
#pragma unroll
          for (int q = 0; q < kPasses; ++q) {
            const int r = r0 + kRowsPerPass * q;
            uint32_t ph_[4], pl_[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // Veltkamp split in fp32 (x = hi + lo exactly, hi has 11 significant bits = an exact fp16), then ONE packed
              // cvt.rn.f16x2.f32 per pair: the scalar F2F conversions run at 16/clk/SM and were the converter's bottleneck.
              const float x0 = v[q][2 * i], x1 = v[q][2 * i + 1];
              const float t0 = __fmul_rn(x0, 8193.0f), t1 = __fmul_rn(x1, 8193.0f);
              const float h0 = __fsub_rn(t0, __fsub_rn(t0, x0)), h1 = __fsub_rn(t1, __fsub_rn(t1, x1));
              const float l0 = __fsub_rn(x0, h0), l1 = __fsub_rn(x1, h1);
              const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn(l0, l1);
              ph_[i] = *reinterpret_cast<const uint32_t*>(&hh);
              pl_[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(hi + off) = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
            *reinterpret_cast<uint4*>(lo + off) = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(full_bar(s));
        if (pt == 0) BNB_TRACE(2, it);
      }
    }
  } else if (warp == kLoadWarp) {
    // ============================== loader warp: async bulk copies of A tiles (fp32, TMA) and B slabs ==
    uint32_t it = 0;
    const int n0 = nt_fix * a.bn;
    const uint32_t bn_bytes = (uint32_t)min(a.bn, a.n_pad - n0) * 128u;
    if (a.b_res && lane == 0) {                             // one-stage layer: this CTA's weight tile is loaded once
      const uint8_t* wsrc = a.Wimg + (size_t)n0 * 128;
      mbar_arrive_expect_tx(bres_bar, 2 * bn_bytes);
      bulk_g2s(bres, wsrc, bn_bytes, bres_bar);
      bulk_g2s(bres + b_bytes, wsrc + (size_t)a.n_pad * 128, bn_bytes, bres_bar);
    }
    for (int mt = mt_first; mt < m_tiles; mt += mt_step) {
      const int m0 = mt * kBM;
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
        mbar_wait_relaxed(empty_bar(s), ph ^ 1);
        const uint32_t dst = base + (uint32_t)s * stage_bytes;
        if (lane == 0) {
          BNB_TRACE(0, it);
          // A: ONE 2-D TMA request per stage (box = box_k x 128 rows of fp32; out-of-range rows / columns arrive as zeros)
          mbar_arrive_expect_tx(raw_bar(s), (uint32_t)a.box_k * 4u * kBM);
          tma_load_2d(dst, &a_map, ks * kBK, m0, raw_bar(s));
          if (a.b_res) mbar_arrive(full_bar(s));
          else {
            mbar_arrive_expect_tx(full_bar(s), 2 * bn_bytes);
            const uint8_t* wsrc = a.Wimg + ((size_t)ks * 2) * (size_t)a.n_pad * 128 + (size_t)n0 * 128;
            bulk_g2s(dst + 2 * a_bytes, wsrc, bn_bytes, full_bar(s));
            bulk_g2s(dst + 2 * a_bytes + b_bytes, wsrc + (size_t)a.n_pad * 128, bn_bytes, full_bar(s));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ================================================
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      if (a.b_res) mbar_wait(bres_bar, 0);
      for (int mt = mt_first; mt < m_tiles; mt += mt_step, ++tcount) {
        const int nt = nt_fix;
        const int n0 = nt * a.bn;
        const int bn = min(a.bn, a.n_pad - n0);
        const uint32_t idesc = make_idesc((uint32_t)bn);
        const int buf = tcount & 1;
        mbar_wait(tempty_bar(buf), ((tcount >> 1) & 1) ^ 1);
        BNB_TRACE(7, tcount);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(full_bar(s), ph);
          BNB_TRACE(3, it);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint64_t d_ahi = make_desc(sa), d_alo = make_desc(sa + a_bytes);
          const uint32_t sb = a.b_res ? bres : sa + 2 * a_bytes;
          const uint64_t d_bhi = make_desc(sb), d_blo = make_desc(sb + b_bytes);
          const int kk_n = min(kBK, a.k_pad - ks * kBK) / 16;
          for (int kk = 0; kk < kk_n; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);        // +32 bytes (>>4) inside the swizzle row
            umma(d_tmem, d_ahi + adv, d_bhi + adv, idesc, (ks | kk) != 0);
            umma(d_tmem, d_alo + adv, d_bhi + adv, idesc, 1);
            umma(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1);
          }
          umma_commit(empty_bar(s));                        // smem stage reusable once these MMAs retire
          BNB_TRACE(4, it);
        }
        umma_commit(tfull_bar(buf));                        // accumulator complete
      }
    }
  } else {
    // ============================== epilogue (warps 0-7) =======================================
    // warp w owns TMEM lanes 32*(w%4).. (rows) and one half of the tile's columns, 32 columns per tcgen05.ld.
    // A thread holds one ROW of the chunk, so direct stores would scatter 16-byte pieces over 32 different rows
    // (measured: 7.5-19 k cycles per tile).  Instead the 32x32 sub-tile goes through a warp-private, XOR-swizzled
    // shared-memory buffer (conflict-free 16 B stores) and is read back 4 rows x 128 B per instruction, so every
    // global store instruction writes four full 128-byte row segments (a TMA tensor store of the same box was tried
    // and is request-bound: 32 row requests per 4 KB).  The residual is added with the same coalesced pattern.
    // Layers whose row pitch is not 16-byte aligned (the 6522-wide head) keep direct stores.
    const int quarter = warp & 3, half = warp >> 2;
    uint8_t* extra = base_ptr + (size_t)a.stages * stage_bytes + bres_bytes;
    float* s_bias = reinterpret_cast<float*>(extra + 8 * 4096 + ((8u * (3 * a.stages + 5) + 16 + 15) & ~15u)) + warp * 128;
    uint8_t* s_out = extra + warp * 4096;                  // 32 rows x 128 B, 1024-aligned
    // this CTA's n-tile never changes: column split and bias strip are set up once
    const int n0 = nt_fix * a.bn;
    const int bn = min(a.bn, a.n_pad - n0);
    const int c_split = min(bn, ((bn / 2 + 31) / 32) * 32);
    const int c_begin = half ? c_split : 0, c_end = half ? bn : c_split;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = c_begin + lane + 32 * j;
      s_bias[lane + 32 * j] = (cc < c_end + 16 && n0 + cc < a.n_pad + 16) ? __ldg(a.bias + n0 + cc) : 0.f;   // bias is padded past n_pad at upload
    }
    __syncwarp();
    uint32_t tcount = 0;
    for (int mt = mt_first; mt < m_tiles; mt += mt_step, ++tcount) {
      const int buf = tcount & 1;
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      if (threadIdx.x == 0) BNB_TRACE(5, tcount);
      tc_fence_after();
      const int m_w = mt * kBM + quarter * 32;             // first row of this warp
      const int m = m_w + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kAccCols;
      float* crow = a.C + (size_t)m * a.N;
      const float* rrow = a.residual ? a.residual + (size_t)m * a.N : nullptr;
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        // residual rows of this 32-column chunk, in the coalesced pattern of the store loop below: issued before the
        // TMEM load and the staging pass so the global-memory latency is off the store path (ncu: 16 % of this kernel's
        // stall samples sat on these loads when they were issued next to the stores)
        float4 rv[8];
        if (a.residual != nullptr && a.c_vec4) {
          const int chunk = lane & 7, ncol = n0 + c0 + 4 * chunk;
          const bool col_ok = (c0 + 4 * chunk < bn) && (ncol + 4 <= a.N);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int mrow = m_w + 4 * j + (lane >> 3);
            rv[j] = (col_ok && mrow < a.M) ? __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)mrow * a.N + ncol))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        if (!(a.dbg & 4)) tmem_ld32(taddr + (uint32_t)c0, r);               // columns beyond bn are never stored (clipped / masked)
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const int n = n0 + c0;
        __syncwarp();                                      // previous chunk's read-back of s_out is complete
        if (!(a.dbg & 4)) tmem_ld_wait();
        const int act = (a.dbg & 1) ? ACT_NONE : a.act;
        if (a.c_vec4) {
          // (1) row-per-lane -> swizzled staging, (2) read back 4 rows x 128 B per instruction, (3) coalesced global store
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bz = *reinterpret_cast<const float4*>(s_bias + (c0 - c_begin) + 4 * j4);
            float4 o;
            o.x = __uint_as_float(r[4 * j4 + 0]) + bz.x; o.y = __uint_as_float(r[4 * j4 + 1]) + bz.y;
            o.z = __uint_as_float(r[4 * j4 + 2]) + bz.z; o.w = __uint_as_float(r[4 * j4 + 3]) + bz.w;
            if (act == ACT_SILU) { silu2(o.x, o.y); silu2(o.z, o.w); }
            else if (act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(s_out + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = o;
          }
          __syncwarp();
          const int chunk = lane & 7, ncol = n + 4 * chunk;
          const bool col_ok = (c0 + 4 * chunk < bn) && (ncol + 4 <= a.N);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 4 * j + (lane >> 3), mrow = m_w + row;
            if (col_ok && mrow < a.M && !(a.dbg & 2)) {
              float4 o = *reinterpret_cast<const float4*>(s_out + row * 128 + ((chunk ^ (row & 7)) << 4));
              if (a.residual) { o.x += rv[j].x; o.y += rv[j].y; o.z += rv[j].z; o.w += rv[j].w; }
              *reinterpret_cast<float4*>(a.C + (size_t)mrow * a.N + ncol) = o;
            }
          }
        } else if (m < a.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < bn && n + j < a.N) {
              float o = act_apply(__uint_as_float(r[j]) + s_bias[(c0 - c_begin) + j], a.act);
              if (rrow) o += __ldg(rrow + n + j);
              crow[n + j] = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (threadIdx.x == 0) BNB_TRACE(6, tcount);
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
  L.k_stages = (K + kBK - 1) / kBK;
  // image layout: [k_stage][hi | lo][n_pad rows][128 B], rows 128B-swizzled in 8-row groups, so ANY row range
  // [n0, n0 + bn) (multiples of 8) of one k-stage is one contiguous byte range => the N tiling is a launch-time choice.
  const size_t plane = (size_t)L.n_pad * 128;
  image->assign((size_t)L.k_stages * 2 * plane, 0);
  for (int ks = 0; ks < L.k_stages; ++ks) {
    uint8_t* hi = image->data() + (size_t)ks * 2 * plane;
    uint8_t* lo = hi + plane;
    for (int n = 0; n < N; ++n)
      for (int kc = 0; kc < kBK; ++kc) {
        const int k = ks * kBK + kc;
        if (k >= K) break;
        const float x = w[(size_t)n * K + k];
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        const int c = kc >> 3, e = kc & 7;
        const size_t off = (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((c ^ (n & 7)) << 4) + (size_t)e * 2;
        const unsigned short hb = __half_as_ushort(h), lb = __half_as_ushort(l);
        memcpy(hi + off, &hb, 2); memcpy(lo + off, &lb, 2);
      }
  }
  return L;
}

// Shared-memory bytes of one launch configuration (must mirror the carve-up inside the kernel).
static size_t smem_for(int bn, int stages, bool b_res) {
  const size_t stage = 2 * (size_t)kBM * 128 + (b_res ? 0 : 2 * (size_t)bn * 128);
  return (size_t)stages * stage + (b_res ? 2 * (size_t)bn * 128 : 0) + 1024 /*alignment*/ + kEpiWarps * 4096 +
         ((8 * (3 * stages + 5) + 16 + 15) & ~15) + kEpiWarps * 128 * sizeof(float);
}

// N tiling for a given M: fill the 148 SMs, keep >= 3 pipeline stages when K is long, always fit >= 2 stages.
void pw_tc_tiling(const PwTcLayer& L, int M, int* bn_out, int* stages_out, size_t* smem_out, int* b_res_out) {
  const int m_tiles = (M + kBM - 1) / kBM;
  const size_t budget = 227 * 1024;
  const bool b_res = L.k_stages == 1;                                                    // weight tile resident per CTA
  int n_tiles = (L.n_pad + 255) / 256;
  auto bn_of = [&](int nt) { return nt == 1 ? L.n_pad : ((L.n_pad + nt - 1) / nt + 31) / 32 * 32; };   // 32-column store chunks must not spill into the next n-tile
  int bn = bn_of(n_tiles);
  while (smem_for(bn, 2, b_res) > budget) { ++n_tiles; bn = bn_of(n_tiles); }            // two stages must fit
  if (L.k_stages >= 3) while (bn > 128) { ++n_tiles; bn = bn_of(n_tiles); }            // deeper ring for long K
  while (m_tiles * ((L.n_pad + bn - 1) / bn) < kNumSMs && bn > 64) { ++n_tiles; bn = bn_of(n_tiles); }
  int stages = 6;
  while (stages > 2 && smem_for(bn, stages, b_res) > budget) --stages;
  *bn_out = bn; *stages_out = stages; *smem_out = smem_for(bn, stages, b_res);
  if (b_res_out) *b_res_out = b_res ? 1 : 0;
}

// 2-D tensor map over the fp32 activation matrix A[M][K] (K contiguous): box = box_k x 128 rows, no swizzle.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
static CUtensorMap encode_map(const float* A, int M, int K, int box_k, int box_rows, bool swizzle128) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::mutex mu;
  static std::map<std::tuple<const float*, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    BNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("pw_tc: cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeFn>(p);
  }
  auto key = std::make_tuple(A, M, K, box_rows * 2 + (swizzle128 ? 1 : 0));
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("pw_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

void pw_tc_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void launch_pw_tc(const PwTcLayer& L, const PwArgs& p, const uint8_t* d_image, cudaStream_t s, LaunchCounter& lc) {
  if (p.a_mode != A_PLAIN || p.a_mul) throw std::runtime_error("pw_tc: A must be a plain [M][K] matrix (run the prep kernel first)");
  int bn = 0, stages = 0;
  size_t smem_bytes = 0;
  int b_res = 0;
  pw_tc_tiling(L, p.M, &bn, &stages, &smem_bytes, &b_res);
  if (smem_bytes > 227 * 1024) throw std::runtime_error("pw_tc: shared memory budget exceeded");
  PwTcArgs a{};
  a.A = p.A; a.Wimg = d_image; a.bias = p.bias; a.C = p.C; a.residual = p.residual; a.gate = p.gate;
  a.M = p.M; a.N = p.N; a.K = p.K; a.rows_per_chunk = p.rows_per_chunk; a.act = p.act;
  a.box_k = p.K < kBK ? p.K : kBK;
  a.b_res = b_res;
  a.n_pad = L.n_pad; a.k_pad = L.k_pad; a.bn = bn; a.n_tiles = (L.n_pad + bn - 1) / bn; a.stages = stages;
  a.c_vec4 = (p.N % 4 == 0) ? 1 : 0;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int tiles = m_tiles * a.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : (kNumSMs / a.n_tiles) * a.n_tiles;   // multiple of n_tiles (N-stationary CTAs)
  const CUtensorMap amap = encode_map(p.A, p.M, p.K, a.box_k, kBM, false);
  // debug timeline: BNB_PWTC_TRACE=<file> BNB_PWTC_TRACE_IDX=<n-th pw_tc launch of the process>
  static const int dbg_flags = getenv("BNB_PWTC_DBG") ? atoi(getenv("BNB_PWTC_DBG")) : 0;
  a.dbg = dbg_flags;
  static std::atomic<long long> launch_idx{0};
  static const char* trace_path = getenv("BNB_PWTC_TRACE");
  static const long long trace_idx = getenv("BNB_PWTC_TRACE_IDX") ? atoll(getenv("BNB_PWTC_TRACE_IDX")) : 0;
  long long* trace = nullptr;
  const long long my_idx = launch_idx.fetch_add(1);
  if (trace_path && my_idx == trace_idx) { BNB_CUDA(cudaMallocManaged(&trace, 8 * 64 * sizeof(long long))); memset(trace, 0, 8 * 64 * sizeof(long long)); a.trace = trace; }
  pw_tc_kernel<<<grid, kThreads, smem_bytes, s>>>(a, amap);
  if (trace) {
    BNB_CUDA(cudaStreamSynchronize(s));
    FILE* f = fopen(trace_path, "w");
    if (f) {
      fprintf(f, "# M=%d N=%d K=%d bn=%d stages=%d grid=%d tiles=%d\n# iter loader_issue conv_raw_ready conv_done mma_full_ready mma_committed epi_acc_ready epi_done mma_tmem_free\n", p.M, p.N, p.K, bn, stages, grid, tiles);
      const long long t0 = trace[0];
      for (int i = 0; i < 64; ++i) { fprintf(f, "%d", i); for (int e = 0; e < 8; ++e) fprintf(f, " %lld", trace[e * 64 + i] ? trace[e * 64 + i] - t0 : -1); fprintf(f, "\n"); }
      fclose(f);
    }
    cudaFree(trace);
  }
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
