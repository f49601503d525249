// mbconv2.cu — the front half of an MBConv block in ONE kernel on fp16 hi/lo activation planes, sm_100a only:
//     1x1 expand (tcgen05.mma, 3-term hi/lo split, fp32 accumulate in TMEM) + bias + SiLU
//  -> 3x3 depthwise (+bias +SiLU, stride 1 or 2, zero padding 1) -> hi/lo planes of the result + per-tile SE sums.
//
// Replaces the CONV_2D(1x1) -> LOGISTIC/MUL -> [PAD] -> DEPTHWISE_CONV_2D -> LOGISTIC/MUL -> MEAN op groups the
// reference executes inside TFLite (/root/reference/internal/inference/tflite/classifier.go:107).
//
// Round-2 design (what changed against mbconv_tc.cu, and why — VERDICT r1 "weak" #4):
//   * the GEMM is TRANSPOSED: M = 128 expanded CHANNELS (weights are the A operand), N = the positions of one halo
//     patch (<= 160), so a TMEM lane is a channel and an epilogue thread owns ONE channel of the whole patch.  The
//     depthwise 3x3 then runs out of registers (three patch rows per thread), with per-thread bias / taps: no shared
//     memory round trip, no group barriers, no bias shuffles, no halo recomputation between the rows of a tile.
//   * activations arrive as fp16 hi and lo planes (x ~= hi + lo), written that way by the producer: the patch goes
//     TMA (4-D box, hardware swizzle, zero fill outside the image) -> shared memory -> tcgen05.mma with no converter
//     warps; the result leaves as hi/lo planes too, so the project GEMM (pw2.cu) needs no converter either.
//   * positions outside the image are skipped (whole rows: warp-uniform) or zeroed (first / last column) instead of
//     evaluated and masked; SiLU shares one reciprocal between four values.
//
// Roles (15 warps): 12 epilogue warps = 3 groups x 4 TMEM lane quarters (group g owns accumulator buffer g),
// 1 MMA issuer, 1 patch loader (TMA), 1 weight loader (cp.async.bulk of pre-swizzled slabs).
#include "mbconv2.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <stdexcept>
#include <string>

#include "tc_common.cuh"
#include "tma_host.h"

namespace bnb {

namespace {

using namespace tc;

constexpr int kGroups = 3;
constexpr int kEpiWarps = 4 * kGroups, kMmaWarp = kEpiWarps, kLoadBWarp = kEpiWarps + 1, kLoadAWarp = kEpiWarps + 2;
constexpr int kThreads = (kEpiWarps + 3) * 32;
constexpr int kTmemCols = 512;

struct Mb2Args {
  const uint8_t* Wimg; const float* bias_e; const float* w_dw; const float* bias_dw;
  __half* dh; __half* dl; float* partial;
  int B, H, W, C, Ho, Wo;
  int TH, PH, n_mma, tiles_h, tiles_w;
  int n_units, k_stages, rot_mode;       // rot_mode: 0 none, 1 last unit replicated over the four lane quarters, 2 four rotated versions of the only unit
  int a_slots, b_slots, a_resident, n_img_units;
  uint32_t a_slot_bytes, a_region_bytes, b_slot_bytes, img_unit_bytes, b_tx_bytes;
  uint32_t st_rb[kMb2MaxStages], st_ksteps[kMb2MaxStages], st_k0[kMb2MaxStages], st_aoff[kMb2MaxStages],
      st_aplane[kMb2MaxStages], st_boff[kMb2MaxStages], st_bplane[kMb2MaxStages], st_map[kMb2MaxStages];
};

template <int PW>
__device__ __forceinline__ void load_row(float (&dst)[PW], uint32_t taddr, bool row_in, bool left_oob, bool right_oob, float be) {
  if (!row_in) {                                  // warp-uniform: the whole patch row lies outside the image
#pragma unroll
    for (int i = 0; i < PW; ++i) dst[i] = 0.f;
    return;
  }
  uint32_t raw[PW];
  tmem_ld_n<PW>(taddr, raw);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < PW; ++i) dst[i] = __uint_as_float(raw[i]) + be;
  silu_n<PW>(dst);
  if (left_oob) dst[0] = 0.f;                     // zero padding lives in the EXPANDED domain
  if (right_oob) dst[PW - 1] = 0.f;
}

template <int S, int TW, int PW>
__device__ __forceinline__ void out_row(const float (&r0)[PW], const float (&r1)[PW], const float (&r2)[PW], const float (&wd)[9],
                                        float bd, __half* ph, __half* pl, int C, bool active, float& lsum) {
#pragma unroll
  for (int o = 0; o < TW; o += 4) {
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = (o + i) * S;
      float v = bd;
      v = fmaf(r0[c0], wd[0], v); v = fmaf(r0[c0 + 1], wd[1], v); v = fmaf(r0[c0 + 2], wd[2], v);
      v = fmaf(r1[c0], wd[3], v); v = fmaf(r1[c0 + 1], wd[4], v); v = fmaf(r1[c0 + 2], wd[5], v);
      v = fmaf(r2[c0], wd[6], v); v = fmaf(r2[c0 + 1], wd[7], v); v = fmaf(r2[c0 + 2], wd[8], v);
      acc[i] = v;
    }
    silu4(acc[0], acc[1], acc[2], acc[3]);
    if (active) {
      lsum += acc[0]; lsum += acc[1]; lsum += acc[2]; lsum += acc[3];        // fixed order: deterministic SE sums
      uint32_t h01, l01, h23, l23;
      split2(acc[0], acc[1], h01, l01); split2(acc[2], acc[3], h23, l23);
      unsigned short* qh = reinterpret_cast<unsigned short*>(ph) + (size_t)o * C;
      unsigned short* ql = reinterpret_cast<unsigned short*>(pl) + (size_t)o * C;
      qh[0] = (unsigned short)(h01 & 0xffffu); ql[0] = (unsigned short)(l01 & 0xffffu);
      qh[C] = (unsigned short)(h01 >> 16);     ql[C] = (unsigned short)(l01 >> 16);
      qh[2 * C] = (unsigned short)(h23 & 0xffffu); ql[2 * C] = (unsigned short)(l23 & 0xffffu);
      qh[3 * C] = (unsigned short)(h23 >> 16);     ql[3 * C] = (unsigned short)(l23 >> 16);
    }
  }
}

template <int S, int TW>
__global__ void __launch_bounds__(kThreads, 1)
mbconv2_kernel(const Mb2Args a, const __grid_constant__ CUtensorMap m_hi0, const __grid_constant__ CUtensorMap m_lo0,
               const __grid_constant__ CUtensorMap m_hi1, const __grid_constant__ CUtensorMap m_lo1) {
  constexpr int PW = (TW - 1) * S + 3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_ring = base;
  const uint32_t b_ring = a_ring + a.a_region_bytes;
  // weight slab address: resident slabs are packed exactly like the image (unit-major, stages inside); the streaming ring has
  // one max-sized slot per slab in flight
  auto a_addr = [&](int slot) {
    return a.a_resident ? a_ring + (uint32_t)(slot / a.k_stages) * a.img_unit_bytes + a.st_aoff[slot % a.k_stages]
                        : a_ring + (uint32_t)slot * a.a_slot_bytes;
  };
  const uint32_t bars = b_ring + (uint32_t)a.b_slots * a.b_slot_bytes;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (a.a_slots + s); };
  const uint32_t bb = bars + 16u * a.a_slots;
  auto b_full = [&](int s) { return bb + 8u * s; };
  auto b_empty = [&](int s) { return bb + 8u * (a.b_slots + s); };
  const uint32_t tb = bb + 16u * a.b_slots;
  auto t_full = [&](int g) { return tb + 8u * g; };
  auto t_empty = [&](int g) { return tb + 8u * (kGroups + g); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (tb - base) + 16 * kGroups);

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.a_slots; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < a.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int g = 0; g < kGroups; ++g) { mbar_init(t_full(g), 1); mbar_init(t_empty(g), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_chunk = a.tiles_h * a.tiles_w;
  const int total_tiles = a.B * tiles_per_chunk;

  if (warp == kLoadBWarp) {
    // ============================== patch loader: per tile, one 4-D TMA box per (stage, plane) ======================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int b = tile / tiles_per_chunk, tt = tile - b * tiles_per_chunk;
        const int ty = tt / a.tiles_w, tx = tt - ty * a.tiles_w;
        const int hi0 = ty * a.TH * S - 1, wi0 = tx * TW * S - 1;
        const int bs = it % a.b_slots; const uint32_t ph = (it / a.b_slots) & 1;
        mbar_wait_relaxed(b_empty(bs), ph ^ 1);
        mbar_arrive_expect_tx(b_full(bs), a.b_tx_bytes);
        for (int s = 0; s < a.k_stages; ++s) {
          const uint32_t dst = b_ring + (uint32_t)bs * a.b_slot_bytes + a.st_boff[s];
          const CUtensorMap* mh = a.st_map[s] ? &m_hi1 : &m_hi0;
          const CUtensorMap* ml = a.st_map[s] ? &m_lo1 : &m_lo0;
          tma_load_4d(dst, mh, (int)a.st_k0[s], wi0, hi0, b, b_full(bs));
          tma_load_4d(dst + a.st_bplane[s], ml, (int)a.st_k0[s], wi0, hi0, b, b_full(bs));
        }
      }
    }
  } else if (warp == kLoadAWarp) {
    // ============================== weight loader: pre-swizzled (unit, stage) slabs, hi | lo ========================
    if (lane == 0) {
      if (a.a_resident) {
        for (int iu = 0; iu < a.n_img_units; ++iu)
          for (int s = 0; s < a.k_stages; ++s) {
            const int slot = iu * a.k_stages + s;
            mbar_arrive_expect_tx(a_full(slot), 2u * a.st_aplane[s]);
            bulk_g2s(a_addr(slot), a.Wimg + (size_t)iu * a.img_unit_bytes + a.st_aoff[s], 2u * a.st_aplane[s], a_full(slot));
          }
      } else {
        uint32_t q = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x)
          for (int u = 0; u < a.n_units; ++u)
            for (int s = 0; s < a.k_stages; ++s, ++q) {
              const int slot = q % a.a_slots; const uint32_t ph = (q / a.a_slots) & 1;
              mbar_wait_relaxed(a_empty(slot), ph ^ 1);
              mbar_arrive_expect_tx(a_full(slot), 2u * a.st_aplane[s]);
              bulk_g2s(a_addr(slot), a.Wimg + (size_t)u * a.img_unit_bytes + a.st_aoff[s], 2u * a.st_aplane[s], a_full(slot));
            }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ======================================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(128u, (uint32_t)a.n_mma);
      uint32_t it = 0, q = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int bs = it % a.b_slots;
        mbar_wait(b_full(bs), (it / a.b_slots) & 1);
        const int rot = it & 3;
        for (int u = 0; u < a.n_units; ++u) {
          const uint32_t seq = it * a.n_units + u;
          const int buf = seq % kGroups;
          mbar_wait(t_empty(buf), ((seq / kGroups) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)buf * kMb2BufCols;
          for (int s = 0; s < a.k_stages; ++s, ++q) {
            int slot; uint32_t ph;
            if (a.a_resident) { slot = (a.rot_mode == 2 ? rot : u) * a.k_stages + s; ph = 0; }
            else { slot = q % a.a_slots; ph = (q / a.a_slots) & 1; }
            mbar_wait(a_full(slot), ph);
            tc_fence_after();
            const uint32_t sa = a_addr(slot);
            const uint32_t sb = b_ring + (uint32_t)bs * a.b_slot_bytes + a.st_boff[s];
            const uint64_t d_whi = make_desc_rb(sa, a.st_rb[s]), d_wlo = make_desc_rb(sa + a.st_aplane[s], a.st_rb[s]);
            const uint64_t d_xhi = make_desc_rb(sb, a.st_rb[s]), d_xlo = make_desc_rb(sb + a.st_bplane[s], a.st_rb[s]);
            for (uint32_t kk = 0; kk < a.st_ksteps[s]; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);           // +32 bytes (>> 4) inside the swizzle row
              umma(d_tmem, d_whi + adv, d_xhi + adv, idesc, (s | kk) != 0);
              umma(d_tmem, d_wlo + adv, d_xhi + adv, idesc, 1);
              umma(d_tmem, d_whi + adv, d_xlo + adv, idesc, 1);
            }
            if (!a.a_resident) umma_commit(a_empty(slot));
          }
          umma_commit(t_full(buf));
        }
        umma_commit(b_empty(bs));
      }
    }
  } else {
    // ============================== epilogue groups (3 x 4 warps): thread = one expanded channel ====================
    const int g = warp >> 2, quarter = warp & 3;
    const uint32_t tbuf = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * kMb2BufCols;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int b = tile / tiles_per_chunk, tt = tile - b * tiles_per_chunk;
      const int ty = tt / a.tiles_w, tx = tt - ty * a.tiles_w;
      const int ho0 = ty * a.TH, wo0 = tx * TW;
      const int hi0 = ho0 * S - 1, wi0 = wo0 * S - 1;
      const bool left_oob = wi0 < 0, right_oob = wi0 + PW - 1 >= a.W;
      const int rot = it & 3;
      for (int u = 0; u < a.n_units; ++u) {
        const uint32_t seq = it * a.n_units + u;
        if ((int)(seq % kGroups) != g) continue;
        const uint32_t par = (seq / kGroups) & 1;
        int qeff = quarter;
        bool warp_active;
        const int rows_u = min(128, a.C - u * 128);
        if (a.rot_mode == 1 && u == a.n_units - 1) { warp_active = quarter == rot; qeff = 0; }
        else if (a.rot_mode == 2) { qeff = (quarter - rot) & 3; warp_active = qeff * 32 < rows_u; }
        else warp_active = qeff * 32 < rows_u;
        if (!warp_active) {                                       // nothing in this lane quarter: handshake only
          mbar_wait(t_full(g), par);
          mbar_arrive(t_empty(g));
          continue;
        }
        const int c = u * 128 + qeff * 32 + lane;
        const bool active = c < a.C;
        float wd[9], bd = 0.f, be = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) wd[t] = active ? __ldg(a.w_dw + (size_t)t * a.C + c) : 0.f;
        if (active) { bd = __ldg(a.bias_dw + c); be = __ldg(a.bias_e + c); }
        mbar_wait(t_full(g), par);
        tc_fence_after();

        float lsum = 0.f;
        float rA[PW], rB[PW], rC[PW];
        auto ld = [&](float (&dst)[PW], int r) {
          const int hi = hi0 + r;
          load_row<PW>(dst, tbuf + (uint32_t)(r * PW), hi >= 0 && hi < a.H, left_oob, right_oob, be);
          if (r == a.PH - 1) { tc_fence_before(); mbar_arrive(t_empty(g)); }     // accumulator fully read: hand the buffer back
        };
        const size_t out_base = (((size_t)b * a.Ho + ho0) * a.Wo + wo0) * a.C + c;
        const size_t row_elems = (size_t)a.Wo * a.C;
        auto out = [&](const float (&r0)[PW], const float (&r1)[PW], const float (&r2)[PW], int oh) {
          out_row<S, TW, PW>(r0, r1, r2, wd, bd, a.dh + out_base + (size_t)oh * row_elems, a.dl + out_base + (size_t)oh * row_elems, a.C, active, lsum);
        };
        if (S == 1) {
          ld(rA, 0); ld(rB, 1);
          for (int oh = 0; oh < a.TH; oh += 3) {
            ld(rC, oh + 2); out(rA, rB, rC, oh);
            if (oh + 1 < a.TH) { ld(rA, oh + 3); out(rB, rC, rA, oh + 1); }
            if (oh + 2 < a.TH) { ld(rB, oh + 4); out(rC, rA, rB, oh + 2); }
          }
        } else {
          ld(rA, 0);
          for (int oh = 0; oh < a.TH; oh += 3) {
            ld(rB, 2 * oh + 1); ld(rC, 2 * oh + 2); out(rA, rB, rC, oh);
            if (oh + 1 < a.TH) { ld(rA, 2 * oh + 3); ld(rB, 2 * oh + 4); out(rC, rA, rB, oh + 1); }
            if (oh + 2 < a.TH) { ld(rC, 2 * oh + 5); ld(rA, 2 * oh + 6); out(rB, rC, rA, oh + 2); }
          }
        }
        if (a.partial != nullptr && active) a.partial[((size_t)b * tiles_per_chunk + tt) * a.C + c] = lsum;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
Mb2Plan mb2_plan(int H, int W, int Ho, int Wo, int stride, int Cin, int C, bool sw_small) {
  Mb2Plan P;
  P.S = stride; P.Cin = Cin; P.C = C;
  if (stride != 1 && stride != 2) return P;
  if (stride == 1) { if (H != Ho || W != Wo) return P; } else { if (H != 2 * Ho || W != 2 * Wo) return P; }
  if (stride == 1) P.TW = (Wo % 16 == 0) ? 16 : ((Wo % 8 == 0) ? 8 : 0);
  else P.TW = (Wo % 8 == 0) ? 8 : 0;
  if (P.TW == 0 || C % 8 || Cin % 4 || Cin > 64 * kMb2MaxStages) return P;
  P.PW = (P.TW - 1) * stride + 3;
  for (int th = Ho; th >= 1; --th) {
    if (Ho % th) continue;
    const int ph = (th - 1) * stride + 3;
    if (ph * P.PW > kMb2BufCols) continue;
    P.TH = th; P.PH = ph; break;
  }
  if (P.TH == 0) return P;
  P.n_mma = (P.PH * P.PW + 15) / 16 * 16;
  P.tiles_h = Ho / P.TH; P.tiles_w = Wo / P.TW;
  P.n_units = (C + 127) / 128;
  // K stages
  int k = 0, ns = 0;
  while (k < Cin) {
    const int left = Cin - k;
    int rb = 128, kw = 64;
    if (left < 64 && sw_small) { if (left <= 16) { rb = 32; kw = 16; } else if (left <= 32) { rb = 64; kw = 32; } }
    P.st_rb[ns] = rb; P.st_k0[ns] = k; P.st_kw[ns] = kw;
    P.st_ksteps[ns] = (std::min(left, kw) + 15) / 16;
    k += kw; ++ns;
  }
  P.k_stages = ns;
  uint32_t aoff = 0, boff = 0;
  P.b_tx_bytes = 0;
  for (int s = 0; s < ns; ++s) {
    P.st_aplane[s] = 128u * (uint32_t)P.st_rb[s];
    P.st_aoff[s] = aoff; aoff += 2 * P.st_aplane[s];
    P.st_bplane[s] = (uint32_t)round_up((size_t)P.n_mma * P.st_rb[s], 1024);
    P.st_boff[s] = boff; boff += 2 * P.st_bplane[s];
    P.b_tx_bytes += 2u * (uint32_t)(P.PH * P.PW * P.st_kw[s] * 2);
  }
  P.img_unit_bytes = aoff;
  P.a_slot_bytes = 2 * P.st_aplane[0];
  P.b_slot_bytes = boff;
  // shared-memory plan: all weight slabs resident when they fit beside a double-buffered patch, else a streaming ring
  const size_t budget = kMb2SmemLimit - 1024 /*alignment*/ - 512 /*barriers*/;
  const bool ragged = (C % 128) != 0;
  const size_t plain_bytes = (size_t)P.n_units * P.img_unit_bytes;           // resident slabs are packed
  if (P.n_units == 1 && ragged && 4 * plain_bytes + 2 * (size_t)P.b_slot_bytes <= budget) {
    P.a_resident = 1; P.a_slots = 4 * ns; P.b_slots = 2; P.a_region_bytes = (uint32_t)(4 * plain_bytes);   // single ragged unit: four rotated versions
  } else if (plain_bytes + 2 * (size_t)P.b_slot_bytes <= budget) {
    P.a_resident = 1; P.a_slots = P.n_units * ns; P.b_slots = 2; P.a_region_bytes = (uint32_t)plain_bytes;
  } else {
    P.a_resident = 0;
    P.b_slots = 2;
    size_t left = budget > 2 * (size_t)P.b_slot_bytes ? budget - 2 * (size_t)P.b_slot_bytes : 0;
    int slots = (int)(left / P.a_slot_bytes);
    if (slots < std::max(2, ns + 1)) { P.b_slots = 1; left = budget - P.b_slot_bytes; slots = (int)(left / P.a_slot_bytes); }
    if (slots < 2) return P;
    P.a_slots = std::min(slots, 6);
    P.a_region_bytes = (uint32_t)P.a_slots * P.a_slot_bytes;
  }
  P.smem_bytes = 1024 + (size_t)P.a_region_bytes + (size_t)P.b_slots * P.b_slot_bytes + 16 * (size_t)(P.a_slots + P.b_slots + kGroups) + 64;
  if (P.smem_bytes < 116 * 1024) P.smem_bytes = 116 * 1024;     // one CTA per SM: the kernel allocates all 512 TMEM columns
  P.ok = P.smem_bytes <= kMb2SmemLimit;
  (void)H; (void)W;
  return P;
}

namespace {
// rotation mode of a plan (see Mb2Args.rot_mode)
int rot_mode_of(const Mb2Plan& P) {
  const int rag = P.C % 128;
  if (rag == 0) return 0;
  if (P.n_units == 1) return (P.a_resident && P.a_slots == 4 * P.k_stages) ? 2 : 0;
  return rag <= 32 ? 1 : 0;
}
}  // namespace

void mb2_prepare_weights(const Mb2Plan& P, const float* w, std::vector<uint8_t>* image) {
  const int rm = rot_mode_of(P);
  const int n_img = rm == 2 ? 4 : P.n_units;
  image->assign((size_t)n_img * P.img_unit_bytes, 0);
  for (int iu = 0; iu < n_img; ++iu) {
    for (int r = 0; r < 128; ++r) {
      int ch;                                             // channel held by row r of this image unit (-1: none)
      if (rm == 2) { const int blk = ((r >> 5) - iu) & 3; ch = blk * 32 + (r & 31); }
      else if (rm == 1 && iu == P.n_units - 1) ch = iu * 128 + (r & 31);       // the <= 32 ragged channels, replicated in every lane quarter
      else ch = iu * 128 + r;
      if (ch >= P.C) ch = -1;
      if (ch < 0) continue;
      for (int s = 0; s < P.k_stages; ++s) {
        uint8_t* hi = image->data() + (size_t)iu * P.img_unit_bytes + P.st_aoff[s];
        uint8_t* lo = hi + P.st_aplane[s];
        for (int kc = 0; kc < P.st_kw[s]; ++kc) {
          const int k = P.st_k0[s] + kc;
          if (k >= P.Cin) break;
          const float x = w[(size_t)ch * P.Cin + k];
          const __half h = __float2half_rn(x);
          const __half l = __float2half_rn(x - __half2float(h));
          const uint32_t off = tc::swz_off((uint32_t)r, (uint32_t)(kc >> 3), (uint32_t)P.st_rb[s]) + (uint32_t)(kc & 7) * 2u;
          const unsigned short hb = __half_as_ushort(h), lb = __half_as_ushort(l);
          memcpy(hi + off, &hb, 2); memcpy(lo + off, &lb, 2);
        }
      }
    }
  }
}

namespace {
template <int S, int TW>
void launch_t(const Mb2Args& a, const CUtensorMap* m, int grid, size_t smem, cudaStream_t s) {
  mbconv2_kernel<S, TW><<<grid, kThreads, smem, s>>>(a, m[0], m[1], m[2], m[3]);
}
}  // namespace

void mb2_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
}

void launch_mbconv2(const Mb2Plan& P, const Mb2Launch& L, cudaStream_t s, LaunchCounter& lc) {
  if (!P.ok) throw std::runtime_error("mbconv2: layer has no plan");
  if (L.x_pitch % 8) throw std::runtime_error("mbconv2: input plane pitch must be a multiple of 8 channels");
  Mb2Args a{};
  a.Wimg = L.Wimg; a.bias_e = L.bias_e; a.w_dw = L.w_dw; a.bias_dw = L.bias_dw; a.dh = L.dh; a.dl = L.dl; a.partial = L.partial;
  a.B = L.B; a.H = L.H; a.W = L.W; a.C = P.C; a.Ho = L.Ho; a.Wo = L.Wo;
  a.TH = P.TH; a.PH = P.PH; a.n_mma = P.n_mma; a.tiles_h = P.tiles_h; a.tiles_w = P.tiles_w;
  a.n_units = P.n_units; a.k_stages = P.k_stages; a.rot_mode = rot_mode_of(P);
  a.a_slots = P.a_slots; a.b_slots = P.b_slots; a.a_resident = P.a_resident; a.n_img_units = a.rot_mode == 2 ? 4 : P.n_units;
  a.a_slot_bytes = P.a_slot_bytes; a.a_region_bytes = P.a_region_bytes; a.b_slot_bytes = P.b_slot_bytes; a.img_unit_bytes = P.img_unit_bytes; a.b_tx_bytes = P.b_tx_bytes;
  CUtensorMap maps[4];
  int n_maps = 0, map_rb[2] = {0, 0}, map_kw[2] = {0, 0};
  for (int st = 0; st < P.k_stages; ++st) {
    a.st_rb[st] = P.st_rb[st]; a.st_ksteps[st] = P.st_ksteps[st]; a.st_k0[st] = P.st_k0[st]; a.st_aoff[st] = P.st_aoff[st];
    a.st_aplane[st] = P.st_aplane[st]; a.st_boff[st] = P.st_boff[st]; a.st_bplane[st] = P.st_bplane[st];
    int mi = -1;
    for (int j = 0; j < n_maps; ++j) if (map_rb[j] == P.st_rb[st] && map_kw[j] == P.st_kw[st]) mi = j;
    if (mi < 0) {
      if (n_maps == 2) throw std::runtime_error("mbconv2: more than two distinct stage shapes");
      mi = n_maps++; map_rb[mi] = P.st_rb[st]; map_kw[mi] = P.st_kw[st];
      const uint64_t dims[4] = {(uint64_t)P.Cin, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.B};
      const uint64_t strides[3] = {(uint64_t)L.x_pitch * 2, (uint64_t)L.W * L.x_pitch * 2, (uint64_t)L.H * L.W * L.x_pitch * 2};
      const uint32_t box[4] = {(uint32_t)P.st_kw[st], (uint32_t)P.PW, (uint32_t)P.PH, 1};
      maps[2 * mi] = tma_encode(L.xh, 2, 4, dims, strides, box, P.st_rb[st]);
      maps[2 * mi + 1] = tma_encode(L.xl, 2, 4, dims, strides, box, P.st_rb[st]);
    }
    a.st_map[st] = (uint32_t)mi;
  }
  if (n_maps == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
  const long long tiles = (long long)L.B * P.tiles_h * P.tiles_w;
  const int grid = tiles < kNumSMs ? (int)tiles : kNumSMs;
  if (P.S == 1 && P.TW == 16) launch_t<1, 16>(a, maps, grid, P.smem_bytes, s);
  else if (P.S == 1 && P.TW == 8) launch_t<1, 8>(a, maps, grid, P.smem_bytes, s);
  else if (P.S == 2 && P.TW == 8) launch_t<2, 8>(a, maps, grid, P.smem_bytes, s);
  else throw std::runtime_error("mbconv2: unsupported tile shape");
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
