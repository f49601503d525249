// pw2.cu — the dense GEMM layers of the conv stack that are NOT fused into mbconv2.cu: every 1x1 project conv
// (+ optional SE gate on the input, + residual), the final 3x3 VALID conv (after a tiny im2col prep) and the FC head,
// on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM), sm_100a only.
//
// Replaces the CONV_2D 1x1 / FULLY_CONNECTED ops the reference executes inside TFLite-XNNPACK
// (/root/reference/internal/inference/tflite/classifier.go:107).
//
// Round-2 design (against pw_tc.cu, VERDICT r1 "weak" #4; measurements in profiles/r02_*): the A operand arrives as a RowTiles
// image (layouts.h) — fp16 hi and lo planes already in this kernel's swizzled shared-memory order, written by the producing
// kernel — so an (m-tile, K stage) block is ONE 32 KB cp.async.bulk and goes straight into tcgen05.mma: no fp32 -> fp16
// converter, no bar.sync between the load and the MMA.  Three MMAs per K-step keep the fp32-level accuracy:
// D += Ahi*Bhi + Alo*Bhi + Ahi*Blo.  Layers with a squeeze-excite gate on A pass through 8 converter warps that rewrite each
// 16-byte piece in place thread-privately (hi+lo -> *gate -> hi/lo), with no barrier among them.  Weights stay resident in
// shared memory whenever a 3-deep A ring still fits (all front-phase layers).  One elected lane issues MMAs / copies
// (tc_common.cuh: elect_one).  The 16 epilogue warps add bias (+ residual, read from the block input's PatchTiles image) and
// write 16-byte pieces STRAIGHT from registers: into the next block's PatchTiles image (a pixel goes to every tile whose halo
// holds it: per-pixel lookup table), into plain planes, or as fp32 (post conv, logits).
#include "pw2.h"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

#include "tc_common.cuh"

namespace bnb {

namespace {

using namespace tc;

constexpr int kEpiWarps = 16, kMmaWarp = 16, kLoadWarp = 17, kConvWarp0 = 18, kConvThreads = 256;
constexpr int kThreads = (kEpiWarps + 2) * 32 + kConvThreads;       // 26 warps: 16 epilogue, MMA issuer, loader, 8 converters (gate layers only)
constexpr int kAccCols = 256;
constexpr uint32_t kABytes = kBM * 128u;      // one fp16 [128][64] plane tile

struct Pw2Args {
  const uint8_t* a_img; const uint8_t* Wimg; const float* bias; const float* gate;
  const uint8_t* r_img; __half* oh; __half* ol; float* out32; uint8_t* o_img;
  int M, N, K, rows_per_chunk, act;
  int o_pitch, out_mode;                  // out_mode: 0 fp32, 1 plain planes, 2 PatchTiles image
  uint32_t a_tile_bytes;
  int n_pad, k_pad, n_tiles, bn, stages, b_res, conv, out_vec;
  int epi_teams;                          // 2: the 16 epilogue warps work as two teams of 8 on alternate m-tiles (narrow layers), 1: all on every tile
  int a_ldgsts;                           // A stage blocks fetched by the loader warp's 32 lanes with cp.async (16 B each) instead of one bulk copy
  int n_acc;                              // independent accumulators per tile (column ranges of bn): MMA i goes to accumulator i % n_acc
  PatchTiles rp, op;                      // residual / output patch layouts
  long long* trace;                       // debug timeline (BNB_PW2_TRACE): [2 CTAs][8 events][64 slots] clock64 stamps, else null
};

// events: 0 kernel start, 1 stage load issued, 2 stage data landed (converter / MMA saw it), 3 stage converted, 4 stage MMAs
// committed, 5 accumulator seen by the epilogue, 6 m-tile stored, 7 kernel end.  Index = stage iteration (1..4) or m-tile count (5, 6).
#define PW2_TRACE(ev, i) do { if (a.trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (i) < 64) \
    a.trace[((blockIdx.x == 0 ? 0 : 1) * 8 + (ev)) * 64 + (i)] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
pw2_kernel(const __grid_constant__ Pw2Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t b_bytes = (uint32_t)a.bn * 128u;
  const uint32_t stage_bytes = 2 * kABytes + (a.b_res ? 0u : 2 * b_bytes);
  const int k_stages_all = (a.K + kBK - 1) / kBK;
  const uint32_t bres_bytes = a.b_res ? (uint32_t)k_stages_all * 2 * b_bytes : 0u;    // resident weights: every K stage of this CTA's n-tile
  const uint32_t bres = base + (uint32_t)a.stages * stage_bytes;
  const uint32_t bars = bres + bres_bytes;
  auto tma_bar = [&](int s) { return bars + 8u * s; };
  auto full_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * a.stages + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (3 * a.stages + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (3 * a.stages + 2 + b); };
  const uint32_t bres_bar = bars + 8u * (3 * a.stages + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (3 * a.stages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(tma_bar(s), a.a_ldgsts ? 33 : 1); mbar_init(full_bar(s), kConvThreads); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiWarps * 32 / a.epi_teams); }
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

  // N-stationary CTAs (as pw_tc.cu): CTA b keeps n-tile (b % n_tiles) and walks over m-tiles b / n_tiles + i * gridDim.x / n_tiles
  const int m_tiles = (a.M + kBM - 1) / kBM;
  const int nt_fix = blockIdx.x % a.n_tiles, mt_first = blockIdx.x / a.n_tiles, mt_step = gridDim.x / a.n_tiles;
  const int k_stages = (a.K + kBK - 1) / kBK;
  if (threadIdx.x == 0) PW2_TRACE(0, 0);
  // PDL (common.cuh): the loader first issues the resident weight tile (a constant), then joins the wait
  pdl_trigger();
  if (warp != kLoadWarp) pdl_wait();

  if (warp >= kConvWarp0) {
    // ============================== converters (gate layers): A <- split((hi + lo) * gate), in place, thread-private ===
    // (Fetching the pieces with LDG straight from the image instead of bulk copy + in-place rewrite was tried in r02: the stage
    // period stayed ~2.4 k cycles — one SM pulls ~13 B/clk from L2 whichever path asks — and the layers whose m-tiles span
    // several chunks got slower: profiles/r02 notes.)
    if (a.conv) {
      const int pt = threadIdx.x - kConvWarp0 * 32;
      const int c = pt & 7, r0 = pt >> 3;
      uint32_t it = 0;
      for (int mt = mt_first; mt < m_tiles; mt += mt_step) {
        const int m0 = mt * kBM;
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          const int s = it % a.stages; const uint32_t ph = (it / a.stages) & 1;
          uint8_t* hi_p = base_ptr + (size_t)s * stage_bytes;
          uint8_t* lo_p = hi_p + kABytes;
          const int k = ks * kBK + c * 8;
          const bool live = k < a.K;                                  // K % 8 == 0 for every gated layer (checked at launch)
          float4 g0[4], g1[4];
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int m = m0 + r0 + 32 * q;
              g0[q] = make_float4(1.f, 1.f, 1.f, 1.f); g1[q] = g0[q];
              if (m < a.M) {
                const float* gp = a.gate + (size_t)(m / a.rows_per_chunk) * a.K + k;
                g0[q] = __ldcg(reinterpret_cast<const float4*>(gp));          // written by the previous kernel: coherent load (PDL, common.cuh)
                g1[q] = __ldcg(reinterpret_cast<const float4*>(gp + 4));
              }
            }
          }
          mbar_wait_relaxed(tma_bar(s), ph);
          if (pt == 0) PW2_TRACE(2, it);
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int r = r0 + 32 * q;
              if (m0 + r >= a.M) continue;                            // rows past M arrived as zeros
              const uint32_t off = swz_off((uint32_t)r, (uint32_t)c, 128u);
              const uint4 h = *reinterpret_cast<const uint4*>(hi_p + off);
              const uint4 l = *reinterpret_cast<const uint4*>(lo_p + off);
              float2 v0 = join2(h.x, l.x), v1 = join2(h.y, l.y), v2 = join2(h.z, l.z), v3 = join2(h.w, l.w);
              v0.x *= g0[q].x; v0.y *= g0[q].y; v1.x *= g0[q].z; v1.y *= g0[q].w;
              v2.x *= g1[q].x; v2.y *= g1[q].y; v3.x *= g1[q].z; v3.y *= g1[q].w;
              uint4 ho, lo;
              split2(v0.x, v0.y, ho.x, lo.x); split2(v1.x, v1.y, ho.y, lo.y);
              split2(v2.x, v2.y, ho.z, lo.z); split2(v3.x, v3.y, ho.w, lo.w);
              *reinterpret_cast<uint4*>(hi_p + off) = ho;
              *reinterpret_cast<uint4*>(lo_p + off) = lo;
            }
          }
          fence_proxy_async();
          mbar_arrive(full_bar(s));
          if (pt == 0) PW2_TRACE(3, it);
        }
      }
    }
  } else if (warp == kLoadWarp) {
    // ============================== loader: one bulk copy per (m-tile, K stage) A block, weight slabs by bulk copies ========
    // (whole warp walks the loop, one elected lane issues: see elect_one())
    uint32_t it = 0, ph = 0; int s = 0;
    const int n0 = nt_fix * a.bn;
    const uint32_t bn_bytes = (uint32_t)min(a.bn, a.n_pad - n0) * 128u;
    if (a.b_res) {                              // this CTA's weight tile, every K stage, loaded once
      if (elect_one()) {
        mbar_arrive_expect_tx(bres_bar, (uint32_t)k_stages * 2 * bn_bytes);
        for (int ks = 0; ks < k_stages; ++ks) {
          const uint8_t* wsrc = a.Wimg + ((size_t)ks * 2) * (size_t)a.n_pad * 128 + (size_t)n0 * 128;
          bulk_g2s(bres + (uint32_t)ks * 2 * b_bytes, wsrc, bn_bytes, bres_bar);
          bulk_g2s(bres + (uint32_t)ks * 2 * b_bytes + b_bytes, wsrc + (size_t)a.n_pad * 128, bn_bytes, bres_bar);
        }
      }
      __syncwarp();
    }
    pdl_wait();
    for (int mt = mt_first; mt < m_tiles; mt += mt_step) {
      for (int ks = 0; ks < k_stages; ++ks, ++it) {
        mbar_wait_relaxed(empty_bar(s), ph ^ 1);
        const uint32_t dst = base + (uint32_t)s * stage_bytes;
        // the (m-tile, stage) A operand, hi | lo, is one contiguous 32 KB block of the RowTiles image
        const uint8_t* asrc = a.a_img + (size_t)mt * a.a_tile_bytes + (size_t)ks * (2 * kABytes);
        if (a.a_ldgsts) {
          // BNB_PW2_LDGSTS=1 (experiment, see launch_pw2): 2048 x 16 B through the LSU path, 64 per lane
#pragma unroll 8
          for (int i = lane; i < 2 * kABytes / 16; i += 32) cp_async16(dst + 16u * (uint32_t)i, asrc + 16 * i);
          cp_async_mbar_arrive(tma_bar(s));
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(tma_bar(s), (a.a_ldgsts ? 0u : 2 * kABytes) + (a.b_res ? 0u : 2 * bn_bytes));
          if (!a.a_ldgsts) bulk_g2s(dst, asrc, 2 * kABytes, tma_bar(s));
          PW2_TRACE(1, it);
          if (!a.b_res) {
            const uint8_t* wsrc = a.Wimg + ((size_t)ks * 2) * (size_t)a.n_pad * 128 + (size_t)n0 * 128;
            bulk_g2s(dst + 2 * kABytes, wsrc, bn_bytes, tma_bar(s));
            bulk_g2s(dst + 2 * kABytes + b_bytes, wsrc + (size_t)a.n_pad * 128, bn_bytes, tma_bar(s));
          }
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ===========================================================================
    // Whole warp, warp-uniform values (counters, no runtime divisions), the tcgen05 instructions issued by one ELECTED lane: a
    // stage's 12 MMAs are consecutive UTCHMMA instructions (see elect_one() and mbconv2.cu).  The products of a tile are dealt
    // round-robin to n_acc column ranges of the accumulator buffer (their sum is the result; the epilogue adds the ranges in a
    // fixed order).
    {
      uint32_t it = 0, tcount = 0, ph = 0; int s = 0;
      if (a.b_res) mbar_wait(bres_bar, 0);
      const int n0 = nt_fix * a.bn;
      const int bn = min(a.bn, a.n_pad - n0);
      const uint32_t idesc = make_idesc((uint32_t)bn);
      for (int mt = mt_first; mt < m_tiles; mt += mt_step, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(tempty_bar(buf), ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kAccCols;
        uint32_t acc = 0, started = 0;                     // next accumulator; how many accumulators have received their first product
        for (int ks = 0; ks < k_stages; ++ks, ++it) {
          mbar_wait(a.conv ? full_bar(s) : tma_bar(s), ph);
          if (a.a_ldgsts && !a.conv) fence_proxy_async();     // cp.async wrote through the generic proxy; the MMA reads through the async one
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint64_t d_ahi = make_desc(sa), d_alo = make_desc(sa + kABytes);
          const uint32_t sb = a.b_res ? bres + (uint32_t)ks * 2 * b_bytes : sa + 2 * kABytes;
          const uint64_t d_bhi = make_desc(sb), d_blo = make_desc(sb + b_bytes);
          const int kk_n = min(kBK, a.k_pad - ks * kBK) >> 4;
          if (elect_one()) {
            if (!a.conv) PW2_TRACE(2, it);
            if (a.n_acc == 1 || a.n_acc == 3) {            // term t -> accumulator t (n_acc == 3) or the single accumulator
              const uint32_t step = a.n_acc == 3 ? (uint32_t)bn : 0u;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < kk_n) {
                  const uint64_t adv = (uint64_t)(kk * 2);
                  const uint32_t first = (uint32_t)((ks | kk) != 0);
                  umma(d_tmem, d_ahi + adv, d_bhi + adv, idesc, first);
                  umma(d_tmem + step, d_alo + adv, d_bhi + adv, idesc, a.n_acc == 3 ? first : 1u);
                  umma(d_tmem + 2 * step, d_ahi + adv, d_blo + adv, idesc, a.n_acc == 3 ? first : 1u);
                }
              }
            } else {
              uint32_t acc_l = acc, started_l = started;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < kk_n) {
                  const uint64_t adv = (uint64_t)(kk * 2);
#pragma unroll
                  for (int t = 0; t < 3; ++t) {                // hi*hi, lo*hi, hi*lo
                    const uint32_t d = d_tmem + acc_l * (uint32_t)bn;
                    umma(d, (t == 1 ? d_alo : d_ahi) + adv, (t == 2 ? d_blo : d_bhi) + adv, idesc, started_l >= (uint32_t)a.n_acc);
                    acc_l = (acc_l + 1 == (uint32_t)a.n_acc) ? 0u : acc_l + 1;
                    ++started_l;
                  }
                }
              }
            }
            umma_commit(empty_bar(s));
            PW2_TRACE(4, it);
          }
          __syncwarp();
          // every lane tracks the accumulator rotation the elected lane just walked through
          { const uint32_t n = 3u * (uint32_t)kk_n; started += n; if (a.n_acc > 1) { acc += n; while (acc >= (uint32_t)a.n_acc) acc -= (uint32_t)a.n_acc; } }
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar(buf));
        __syncwarp();
      }
    }
  } else {
    // ============================== epilogue (warps 0-15) ===============================================================
    // warp w owns TMEM lanes 32*(w%4).. (rows) and the 16-column chunks c with c % 4 == w / 4.  A thread holds one ROW of the
    // chunk = two 8-channel pieces, i.e. two adjacent 16-byte pieces of each fp16 plane: they go straight from registers to
    // memory (no shared-memory transpose: the staging buffers cost 32 KB that now hold resident weights, and their traffic
    // competed with the MMA operand reads).  Sixteen warps because one warp per 32x32 sub-tile was latency-bound
    // (profiles/r02 timeline: 4.5 k cycles per m-tile).
    // Narrow layers (bn <= 64: a warp would own one chunk per m-tile) run the warps as TWO TEAMS on alternate m-tiles = alternate
    // TMEM buffers: the epilogue of a tile is a chain of dependent global accesses (lookup table -> residual pieces -> stores:
    // 2-4 k cycles each while the copy engine keeps the memory system busy), and one team handled 8.3 k cycles per tile
    // against 5-6 k of main loop (r02 timeline of block 1); two tiles in flight hide it.
    const int quarter = warp & 3, sub = warp >> 2;
    const int team = a.epi_teams == 2 ? (sub >> 1) : 0, csub = a.epi_teams == 2 ? (sub & 1) : sub, cstep = a.epi_teams == 2 ? 32 : 64;
    const int n0 = nt_fix * a.bn;
    const int bn = min(a.bn, a.n_pad - n0);
    const bool plane_mode = a.out_mode != 0;
    uint32_t tcount = 0;
    for (int mt = mt_first; mt < m_tiles; mt += mt_step, ++tcount) {
      const int buf = tcount & 1;
      if (a.epi_teams == 2 && buf != team) continue;
      const int mrow = mt * kBM + quarter * 32 + lane;
      const bool row_ok = mrow < a.M;
      // where this row lives in the residual / output patch images (one table lookup per row and m-tile)
      int pb = 0; uint32_t res_ent = 0xffffffffu; uint4 dst_ent = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (plane_mode && row_ok) {
        uint32_t pix;
        if (a.r_img != nullptr) { a.rp.split_chunk((uint32_t)mrow, &pb, &pix); res_ent = __ldg(a.rp.res_tbl + pix); }
        if (a.out_mode == 2) { a.op.split_chunk((uint32_t)mrow, &pb, &pix); dst_ent = __ldg(a.op.dst_tbl + pix); }
      }
      // residual pieces (the block input: pixel m is interior to exactly one patch tile of its image): independent of the
      // accumulator, so the first chunk's are requested BEFORE the wait for it — on the 36-column front layers a warp has one
      // chunk per m-tile and their L2 latency (1-2 k cycles) was the epilogue's critical path (r02 timeline: 8.5 k cycles per tile)
      uint4 rvh[2], rvl[2];
      auto load_res = [&](int n) {
        rvh[0] = make_uint4(0, 0, 0, 0); rvl[0] = rvh[0]; rvh[1] = rvh[0]; rvl[1] = rvh[0];
        if (plane_mode && a.r_img != nullptr && row_ok && n < a.o_pitch) {
          int rs, rchunk;
          a.rp.stage_of(n >> 3, &rs, &rchunk);                               // piece 0: an even chunk of its stage (n % 16 == 0)
          const size_t a0 = a.rp.entry_piece(pb, res_ent, rs, rchunk);
          const uint32_t lo_off = a.rp.st_plane[rs];
          if (n + 8 < a.o_pitch) {
            // both pieces: chunks c and c + 1 of the same swizzled row differ only in address bit 4 -> ONE aligned 32-byte sector
            // per plane (coherent 256-bit loads: the image was written by an earlier kernel, PDL)
            const uint8_t* base = a.r_img + (a0 & ~(size_t)31);
            const bool swapped = (a0 & 16) != 0;
            uint4 h0, h1, l0, l1;
            ldg256_cg(base, h0, h1); ldg256_cg(base + lo_off, l0, l1);
            rvh[0] = swapped ? h1 : h0; rvh[1] = swapped ? h0 : h1; rvl[0] = swapped ? l1 : l0; rvl[1] = swapped ? l0 : l1;
          } else {
            rvh[0] = __ldcg(reinterpret_cast<const uint4*>(a.r_img + a0));
            rvl[0] = __ldcg(reinterpret_cast<const uint4*>(a.r_img + a0 + lo_off));
          }
        }
      };
      if (csub * 16 < bn) load_res(n0 + csub * 16);
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) PW2_TRACE(5, tcount);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * kAccCols;
      for (int c0 = csub * 16; c0 < bn; c0 += cstep) {
        const int n = n0 + c0;
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        if (c0 != csub * 16) load_res(n);
        tmem_ld_wait();
        for (int ac = 1; ac < a.n_acc; ++ac) {             // add the other accumulators of the tile (fixed order)
          uint32_t r2[16];
          tmem_ld16(taddr + (uint32_t)(ac * bn + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        }
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bz = __ldg(reinterpret_cast<const float4*>(a.bias + n + 4 * q));      // bias is zero-padded past n_pad
          float4 o;
          o.x = __uint_as_float(r[4 * q + 0]) + bz.x; o.y = __uint_as_float(r[4 * q + 1]) + bz.y;
          o.z = __uint_as_float(r[4 * q + 2]) + bz.z; o.w = __uint_as_float(r[4 * q + 3]) + bz.w;
          if (a.act == ACT_SILU) silu4(o.x, o.y, o.z, o.w);
          else if (a.act == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          v[4 * q] = o.x; v[4 * q + 1] = o.y; v[4 * q + 2] = o.z; v[4 * q + 3] = o.w;
        }
        if (row_ok && plane_mode && n < a.o_pitch) {
          const bool two = n + 8 < a.o_pitch;              // past the last piece of the pitch nothing is stored (pad columns INSIDE a piece get exact zeros: zero weights, zero bias)
          uint4 ho[2], lo[2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            float* w = v + 8 * p;
            if (a.r_img != nullptr) {
              const float2 q0 = join2(rvh[p].x, rvl[p].x), q1 = join2(rvh[p].y, rvl[p].y), q2 = join2(rvh[p].z, rvl[p].z), q3 = join2(rvh[p].w, rvl[p].w);
              w[0] += q0.x; w[1] += q0.y; w[2] += q1.x; w[3] += q1.y; w[4] += q2.x; w[5] += q2.y; w[6] += q3.x; w[7] += q3.y;
            }
            split2(w[0], w[1], ho[p].x, lo[p].x); split2(w[2], w[3], ho[p].y, lo[p].y); split2(w[4], w[5], ho[p].z, lo[p].z); split2(w[6], w[7], ho[p].w, lo[p].w);
          }
          if (a.out_mode == 1) {
            *reinterpret_cast<uint4*>(a.oh + (size_t)mrow * a.o_pitch + n) = ho[0];
            *reinterpret_cast<uint4*>(a.ol + (size_t)mrow * a.o_pitch + n) = lo[0];
            if (two) {
              *reinterpret_cast<uint4*>(a.oh + (size_t)mrow * a.o_pitch + n + 8) = ho[1];
              *reinterpret_cast<uint4*>(a.ol + (size_t)mrow * a.o_pitch + n + 8) = lo[1];
            }
          } else {
            // the next block's PatchTiles image: the pixel goes into every tile whose halo patch contains it (<= 4, from the table).
            // The two pieces are chunks c, c + 1 (c even) of one swizzled row = the two halves of one aligned 32-byte sector, in
            // either order: ONE 256-bit store per plane and destination instead of two half-sector stores (the front layers'
            // epilogue was bound by store sectors: r02 timeline of block 1, 8.3 k cycles per m-tile)
            int os, ochunk;
            a.op.stage_of(n >> 3, &os, &ochunk);
            const uint32_t lo_off = a.op.st_plane[os];
            const uint32_t e4[4] = {dst_ent.x, dst_ent.y, dst_ent.z, dst_ent.w};
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              if (e4[d] != 0xffffffffu) {
                const size_t a0 = a.op.entry_piece(pb, e4[d], os, ochunk);
                if (two) {
                  uint8_t* dst = a.o_img + (a0 & ~(size_t)31);
                  const bool swapped = (a0 & 16) != 0;
                  stg256(dst, swapped ? ho[1] : ho[0], swapped ? ho[0] : ho[1]);
                  stg256(dst + lo_off, swapped ? lo[1] : lo[0], swapped ? lo[0] : lo[1]);
                } else {
                  *reinterpret_cast<uint4*>(a.o_img + a0) = ho[0];
                  *reinterpret_cast<uint4*>(a.o_img + a0 + lo_off) = lo[0];
                }
              }
            }
          }
        } else if (row_ok) {
          // fp32 output: this thread's 16 consecutive columns of row mrow
          float* dst = a.out32 + (size_t)mrow * a.N + n;
          if (a.out_vec == 4 && (a.N & 7) == 0 && c0 + 16 <= bn && n + 16 <= a.N) {        // whole 64-byte run: two 256-bit stores
            stg256(dst, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                   make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
            stg256(dst + 8, make_uint4(__float_as_uint(v[8]), __float_as_uint(v[9]), __float_as_uint(v[10]), __float_as_uint(v[11])),
                   make_uint4(__float_as_uint(v[12]), __float_as_uint(v[13]), __float_as_uint(v[14]), __float_as_uint(v[15])));
          } else
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ncol = n + 4 * q;
            if (c0 + 4 * q >= bn) continue;
            if (a.out_vec == 4) { if (ncol + 4 <= a.N) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]); }
            else if (a.out_vec == 2) {
              if (ncol + 2 <= a.N) *reinterpret_cast<float2*>(dst + 4 * q) = make_float2(v[4 * q], v[4 * q + 1]);
              if (ncol + 4 <= a.N) *reinterpret_cast<float2*>(dst + 4 * q + 2) = make_float2(v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) if (ncol + i < a.N) dst[4 * q + i] = v[4 * q + i];
            }
          }
        }
        __syncwarp();                                        // reconverge before the next (warp-aligned) TMEM load
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (threadIdx.x == 0) PW2_TRACE(6, tcount);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) PW2_TRACE(7, 0);
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kAccCols));
  }
}

// b_res_stages = number of K stages held resident (0 = weights stream with the A tiles)
size_t smem_for(int bn, int stages, int b_res_stages) {
  const size_t stage = 2 * (size_t)kABytes + (b_res_stages ? 0 : 2 * (size_t)bn * 128);
  return (size_t)stages * stage + (size_t)b_res_stages * 2 * (size_t)bn * 128 + 1024 /*alignment*/ +
         ((8 * (3 * stages + 5) + 16 + 15) & ~15);
}

}  // namespace

void pw2_tiling(const PwTcLayer& L, int M, bool conv, int* bn_out, int* stages_out, size_t* smem_out, int* b_res_out) {
  const int m_tiles = (M + kBM - 1) / kBM;
  const size_t budget = 227 * 1024;
  int n_tiles = (L.n_pad + 255) / 256;
  auto bn_of = [&](int nt) { return nt == 1 ? L.n_pad : ((L.n_pad + nt - 1) / nt + 15) / 16 * 16; };   // 16-column epilogue chunks must not spill into the next n-tile
  int bn = bn_of(n_tiles);
  while (smem_for(bn, 2, 0) > budget) { ++n_tiles; bn = bn_of(n_tiles); }
  if (L.k_stages >= 3) while (bn > 128) { ++n_tiles; bn = bn_of(n_tiles); }
  while (m_tiles * ((L.n_pad + bn - 1) / bn) < kNumSMs && bn > 64) { ++n_tiles; bn = bn_of(n_tiles); }
  // weights resident (every K stage of the CTA's n-tile) when that still leaves a 3-deep A ring: N-stationary CTAs then
  // read them once instead of once per m-tile (profiles/r02: the 288 -> 72 projects moved 100 KB of weights per 160 KB of A)
  const bool b_res = smem_for(bn, 3, L.k_stages) <= budget && (size_t)L.k_stages * 2 * bn * 128 < (1u << 20);
  int stages = conv ? 6 : 5;
  while (stages > 2 && smem_for(bn, stages, b_res ? L.k_stages : 0) > budget) --stages;
  if (stages > L.k_stages * 3 && stages > 2) stages = L.k_stages * 3 > 2 ? L.k_stages * 3 : 2;   // no point in a ring far deeper than the work
  *bn_out = bn; *stages_out = stages; *smem_out = smem_for(bn, stages, b_res ? L.k_stages : 0);
  if (b_res_out) *b_res_out = b_res ? 1 : 0;
}

void pw2_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(pw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void launch_pw2(const PwTcLayer& L, const Pw2Launch& p, cudaStream_t s, LaunchCounter& lc) {
  const int n_out = (p.out32 != nullptr) + (p.oh != nullptr) + (p.o_img != nullptr);
  if (n_out != 1) throw std::runtime_error("pw2: exactly one output (fp32, plain planes, patch image) must be set");
  if (p.oh && (p.o_pitch % 8 || p.o_pitch < p.N)) throw std::runtime_error("pw2: plane pitch must be a multiple of 8 and >= N");
  if (p.o_img && (p.o_patch.C != p.N || !p.o_patch.dst_tbl)) throw std::runtime_error("pw2: output patch layout does not match N or has no device tables");
  if (p.r_img && !p.r_patch.res_tbl) throw std::runtime_error("pw2: residual patch layout has no device table");
  if (p.r_img && (p.r_patch.C != p.N || p.r_patch.S != 1 || p.out32)) throw std::runtime_error("pw2: residual needs a stride-1 patch layout with N channels and a plane output");
  if (p.gate && (p.K % 8 || p.rows_per_chunk <= 0)) throw std::runtime_error("pw2: gated layers need K % 8 == 0");
  int bn = 0, stages = 0, b_res = 0;
  size_t smem_bytes = 0;
  pw2_tiling(L, p.M, p.gate != nullptr, &bn, &stages, &smem_bytes, &b_res);
  if (smem_bytes > 227 * 1024) throw std::runtime_error("pw2: shared memory budget exceeded");
  if (smem_bytes < 116 * 1024) smem_bytes = 116 * 1024;          // one CTA per SM (all 512 TMEM columns)
  Pw2Args a{};
  a.a_img = p.a_img; a.Wimg = p.Wimg; a.bias = p.bias; a.gate = p.gate; a.r_img = p.r_img; a.oh = p.oh; a.ol = p.ol; a.out32 = p.out32; a.o_img = p.o_img;
  a.M = p.M; a.N = p.N; a.K = p.K; a.rows_per_chunk = p.rows_per_chunk > 0 ? p.rows_per_chunk : 1; a.act = p.act;
  a.out_mode = p.out32 ? 0 : (p.oh ? 1 : 2);
  a.o_pitch = p.oh ? p.o_pitch : (p.N + 7) / 8 * 8;
  a.a_tile_bytes = RowTiles::make(p.K).tile_bytes;
  a.rp = p.r_patch; a.op = p.o_patch;
  a.n_pad = L.n_pad; a.k_pad = L.k_pad; a.bn = bn; a.n_tiles = (L.n_pad + bn - 1) / bn; a.stages = stages; a.b_res = b_res;
  a.conv = p.gate != nullptr ? 1 : 0;
  a.epi_teams = bn <= 64 ? 2 : 1;
  { static const int et = getenv("BNB_PW2_TEAMS") ? atoi(getenv("BNB_PW2_TEAMS")) : 0; if (et == 1 || et == 2) a.epi_teams = et; }
  // experiment knob: A stage blocks by 2048 per-thread cp.async instead of one bulk copy.  Measured (r02 run k16): the blocks
  // still land ~2.5 k cycles apart and the step is 1 % slower, so one SM ingests ~13 B/clk whichever engine asks: default off.
  { static const int ld = getenv("BNB_PW2_LDGSTS") ? atoi(getenv("BNB_PW2_LDGSTS")) : 0; a.a_ldgsts = ld; }
  // one accumulator by default; BNB_PW2_NACC=2|3 deals the products to independent accumulators (experiment knob: no gain measured)
  a.n_acc = 1;
  { static const int forced = getenv("BNB_PW2_NACC") ? atoi(getenv("BNB_PW2_NACC")) : 0;
    if (forced > 1) a.n_acc = std::min(forced, 3 * bn <= kAccCols ? 3 : (2 * bn <= kAccCols ? 2 : 1)); }
  a.out_vec = (p.N % 4 == 0) ? 4 : ((p.N % 2 == 0) ? 2 : 1);
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int tiles = m_tiles * a.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : (kNumSMs / a.n_tiles) * a.n_tiles;
  static std::atomic<long long> launch_idx{0};
  static const char* trace_path = getenv("BNB_PW2_TRACE");
  static const long long trace_idx = getenv("BNB_PW2_TRACE_IDX") ? atoll(getenv("BNB_PW2_TRACE_IDX")) : 0;
  const long long my_idx = launch_idx.fetch_add(1);
  long long* trace = nullptr;
  if (trace_path && my_idx == trace_idx) { BNB_CUDA(cudaMalloc(&trace, 2 * 8 * 64 * sizeof(long long))); BNB_CUDA(cudaMemsetAsync(trace, 0, 2 * 8 * 64 * sizeof(long long), s)); a.trace = trace; }
  launch_k(pw2_kernel, dim3(grid), dim3(kThreads), smem_bytes, s, a);
  if (trace) {
    std::vector<long long> h(2 * 8 * 64);
    BNB_CUDA(cudaStreamSynchronize(s));
    BNB_CUDA(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "w");
    if (f) {
      fprintf(f, "# pw2 M=%d N=%d K=%d conv=%d bn=%d stages=%d b_res=%d grid=%d m_tiles=%d out_mode=%d residual=%d\n", p.M, p.N, p.K, a.conv, bn, stages, b_res, grid, m_tiles, a.out_mode, p.r_img != nullptr);
      fprintf(f, "# cta idx start load_issue landed converted mma_committed acc_seen tile_stored end   (cycles since the CTA's start)\n");
      for (int c = 0; c < 2; ++c) {
        const long long t0 = h[(c * 8 + 0) * 64];
        for (int i = 0; i < 64; ++i) {
          bool any = false;
          for (int e = 0; e < 8; ++e) any = any || h[(c * 8 + e) * 64 + i] != 0;
          if (!any) continue;
          fprintf(f, "%d %d", c, i);
          for (int e = 0; e < 8; ++e) { const long long v = h[(c * 8 + e) * 64 + i]; fprintf(f, " %lld", v ? v - t0 : -1); }
          fprintf(f, "\n");
        }
      }
      fclose(f);
    }
    cudaFree(trace);
  }
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
