// mbconv2.cu — the front half of an MBConv block in ONE kernel on fp16 hi/lo activation planes, sm_100a only:
//     1x1 expand (tcgen05.mma, 3-term hi/lo split, fp32 accumulate in TMEM) + bias + SiLU
//  -> 3x3 depthwise (+bias +SiLU, stride 1 or 2, zero padding 1) -> hi/lo planes of the result + per-tile SE sums.
//
// Replaces the CONV_2D(1x1) -> LOGISTIC/MUL -> [PAD] -> DEPTHWISE_CONV_2D -> LOGISTIC/MUL -> MEAN op groups the
// reference executes inside TFLite (/root/reference/internal/inference/tflite/classifier.go:107).
//
// Round-2 design (what changed against mbconv_tc.cu, and why — VERDICT r1 "weak" #4; measurements in profiles/r02_*):
//   * the GEMM is TRANSPOSED: M = 128 expanded CHANNELS (weights are the A operand), N = patch positions, so a TMEM lane is
//     a channel and an epilogue thread owns ONE channel of the whole patch.  The depthwise 3x3 then runs out of registers
//     (three rotating patch rows per thread), with per-thread bias / taps: no shared memory round trip, no group barriers,
//     no halo recomputation between the rows of a tile.
//   * activations arrive as fp16 hi and lo planes (x ~= hi + lo) in a PatchTiles image (layouts.h): the producer already
//     wrote every tile's halo patch in this kernel's shared-memory order, so a patch plane is ONE cp.async.bulk (a 4-D TMA
//     box is served row by row, ~10 ns per row: request-bound) and goes into tcgen05.mma with no converter warps.  The result
//     leaves as a RowTiles image, i.e. as the project GEMM's (pw2.cu) A operand in its shared-memory order.
//   * TWO tiles per MMA: a slot holds the patches of two consecutive tiles side by side per plane (N = 2 * n_mma <= 256) —
//     a tcgen05.mma costs ~130 cycles here whatever its N (N = 48 / 80 / 112 all gave 15 MMAs = 2.0 k cycles), so the MMA
//     count per tile had to halve.  Accumulators rotate over 2..4 TMEM buffers (512 / (2 * n_mma)).
//   * one ELECTED lane issues MMAs and bulk copies (elect.sync): behind `lane == 0` every UTCHMMA / UBLKCP sat in a waterfall
//     loop (tc_common.cuh: elect_one); the issue loops use counters, no runtime divisions (uniform datapath).
//   * positions outside the image are skipped (whole rows: warp-uniform) or zeroed (first / last column) BEFORE the SiLU;
//     SiLU shares one reciprocal between four values and uses packed fp32 multiplies / adds.
//   * programmatic dependent launch: barrier init, TMEM allocation and the resident weight loads run while the previous
//     kernel drains (common.cuh).
//
// Roles (19 warps, 608 threads, 96 registers): 16 epilogue warps = 4 groups x 4 TMEM lane quarters (group g: pair-units of
// parity g & 1, tile g >> 1 of the pair), 1 MMA issuer, 1 patch loader, 1 weight loader (resident slabs once, or a streaming ring).
#include "mbconv2.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include <stdexcept>
#include <string>

#include "tc_common.cuh"

namespace bnb {

namespace {

using namespace tc;

constexpr int kGroups = 4;          // epilogue groups = accumulator buffers in flight (profiles/r02: one warp per lane quarter and unit is
                                    // latency-bound at ~0.25 IPC, so throughput comes from units in flight: 4 x narrower tiles beat 3 x wide ones)
constexpr int kEpiWarps = 4 * kGroups, kMmaWarp = kEpiWarps, kLoadBWarp = kEpiWarps + 1, kLoadAWarp = kEpiWarps + 2;
constexpr int kThreads = (kEpiWarps + 3) * 32;
constexpr int kTmemCols = 512;

struct Mb2Args {
  const uint8_t* Wimg; const float* bias_e; const float* w_dw; const float* bias_dw;
  const uint8_t* x_img; uint8_t* d_img; uint32_t d_tile_bytes; float* partial;
  int B, H, W, C, Ho, Wo;
  int TH, PH, n_mma, tiles_h, tiles_w;
  int n_units, k_stages, rot_mode;       // rot_mode: 0 none, 1 last unit replicated over the four lane quarters, 2 four rotated versions of the only unit
  int a_slots, b_slots, a_resident, n_img_units;
  int dbg;                                // BNB_MB2_DBG timing experiments (wrong results): 1 = epilogue skips its TMEM loads, 2 = no MMAs issued, 4 = epilogue skips its stores
  uint32_t a_slot_bytes, a_region_bytes, b_slot_bytes, img_unit_bytes;
  int n_bufs;                             // accumulator buffers in TMEM: 512 / (2 * n_mma) columns each, 2..4
  uint32_t b_pair_bytes;                 // shared-memory bytes of one PAIR slot (two tiles' patches side by side per plane)
  uint32_t st_poff[kMb2MaxStages], st_pplane[kMb2MaxStages];   // stage offset / plane bytes inside a pair slot
  long long* trace;                      // debug timeline (BNB_MB2_TRACE): [2 CTAs][8 events][64 slots] clock64 stamps, else null
  uint32_t st_rb[kMb2MaxStages], st_ksteps[kMb2MaxStages], st_k0[kMb2MaxStages], st_aoff[kMb2MaxStages],
      st_aplane[kMb2MaxStages], st_boff[kMb2MaxStages], st_bplane[kMb2MaxStages];
};

// events: 0 kernel start, 1 patch load issued, 2 patch landed (MMA saw it), 3 unit MMAs committed, 4 accumulator seen by
// the epilogue group, 5 accumulator released (all rows read), 6 unit done (stores issued), 7 kernel end.  Index = tile
// iteration (events 1, 2) or unit sequence number (3..6).  CTA 0 and the last CTA are recorded.
#define MB2_TRACE(ev, i) do { if (a.trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (i) < 64) \
    a.trace[((blockIdx.x == 0 ? 0 : 1) * 10 + (ev)) * 64 + (i)] = clock64(); } while (0)

template <int PW, int POLY>
__device__ __forceinline__ void load_row(float (&dst)[PW], uint32_t taddr, bool row_in, bool left_oob, bool right_oob, float be, int dbg = 0) {
  if (!row_in) {                                  // warp-uniform: the whole patch row lies outside the image
#pragma unroll
    for (int i = 0; i < PW; ++i) dst[i] = 0.f;
    return;
  }
  uint32_t raw[PW];
  if (!(dbg & 1)) { tmem_ld_n<PW>(taddr, raw); tmem_ld_wait(); }
  else {
#pragma unroll
    for (int i = 0; i < PW; ++i) raw[i] = 0x3f000000u + (uint32_t)i;
  }
#pragma unroll
  for (int i = 0; i < PW; ++i) dst[i] = __uint_as_float(raw[i]) + be;
  // zero padding lives in the EXPANDED domain.  The columns outside the image hold whatever the patch image held before
  // (they are never written): zero them BEFORE the SiLU — silu(0) = 0, and the shared reciprocal of a group of four must
  // not depend on stale data (it would change the last bits of the three valid neighbours: results must be bit-identical
  // whatever ran in the buffer before).
  if (left_oob) dst[0] = 0.f;
  if (right_oob) dst[PW - 1] = 0.f;
  silu_n<PW, POLY>(dst);
}

// TW outputs of one output row from three SiLU'd patch rows.  `prow` points at this thread's 2-byte slot of output 0 in the hi
// plane of the RowTiles image (row r0 = multiple of TW; the swizzle term of output o is ((o + (r0 & 7)) & 7): r0_lo = r0 & 7
// is 0 for TW = 8 and 0 or 4 for TW = 4); chunk16 = this channel's 16-byte chunk index << 4; the lo plane is + 16384.
template <int S, int TW, int PW, int POLY>
__device__ __forceinline__ void out_row(const float (&r0)[PW], const float (&r1)[PW], const float (&r2)[PW], const float (&wd)[9],
                                        float bd, uint8_t* prow, uint32_t chunk16, uint32_t r0_lo, bool active, float& lsum) {
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
    silu4<POLY>(acc[0], acc[1], acc[2], acc[3]);
    if (active) {
      lsum += acc[0]; lsum += acc[1]; lsum += acc[2]; lsum += acc[3];        // fixed order: deterministic SE sums
      uint32_t h01, l01, h23, l23;
      split2(acc[0], acc[1], h01, l01); split2(acc[2], acc[3], h23, l23);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = (uint32_t)(o + i) * 128u + (chunk16 ^ ((((uint32_t)(o + i) + r0_lo) & 7u) << 4));
        const uint32_t hw = i < 2 ? h01 : h23, lw = i < 2 ? l01 : l23;
        *reinterpret_cast<unsigned short*>(prow + off) = (unsigned short)((i & 1) ? (hw >> 16) : (hw & 0xffffu));
        *reinterpret_cast<unsigned short*>(prow + off + 16384) = (unsigned short)((i & 1) ? (lw >> 16) : (lw & 0xffffu));
      }
    }
  }
}

template <int S, int TW, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
mbconv2_kernel(const Mb2Args a) {
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
  auto a_res_addr = [&](int iu, int s) { return a_ring + (uint32_t)iu * a.img_unit_bytes + a.st_aoff[s]; };
  const uint32_t bars = b_ring + (uint32_t)a.b_slots * a.b_pair_bytes;
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
    for (int g = 0; g < a.n_bufs; ++g) { mbar_init(t_full(g), 1); mbar_init(t_empty(g), 256); }
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
  if (threadIdx.x == 0) MB2_TRACE(0, 0);
  // PDL (common.cuh): everything above is independent of the previous kernel.  The weight loader never touches activations, so
  // it alone skips the wait: resident weights stream in while the previous kernel's last tiles drain.
  pdl_trigger();
  if (warp != kLoadAWarp) pdl_wait();

  if (warp == kLoadBWarp) {
    // ============================== patch loader: the tile's whole shared-memory slot is ONE contiguous image in memory ===
    // (whole warp walks the loop; one elected lane issues: see elect_one())
    // A slot holds a PAIR of consecutive tiles of this CTA: per K stage [hi plane: tile 0 rows | tile 1 rows][lo plane: ...],
    // so that ONE MMA covers both patches (N = 2 * n_mma).  Why: a tcgen05.mma costs ~130 cycles here whatever its N
    // (profiles/r02: N = 48 / 80 / 112 all gave 15 MMAs = 2.0 k cycles), so the MMA count per tile had to come down.
    uint32_t pit = 0, ph = 0; int bs = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += 2 * gridDim.x, ++pit) {
      mbar_wait_relaxed(b_empty(bs), ph ^ 1);
      if (elect_one()) {
        const bool two = tile + (int)gridDim.x < total_tiles;
        uint32_t tx = 0;
        for (int s = 0; s < a.k_stages; ++s) tx += 2u * (uint32_t)a.n_mma * a.st_rb[s];
        mbar_arrive_expect_tx(b_full(bs), two ? 2u * tx : tx);
        const uint32_t slot = b_ring + (uint32_t)bs * a.b_pair_bytes;
        for (int h = 0; h < (two ? 2 : 1); ++h) {
          const uint8_t* src = a.x_img + (size_t)(tile + h * (int)gridDim.x) * a.b_slot_bytes;
          for (int s = 0; s < a.k_stages; ++s) {
            const uint32_t bytes = (uint32_t)a.n_mma * a.st_rb[s];
            bulk_g2s(slot + a.st_poff[s] + (uint32_t)h * bytes, src + a.st_boff[s], bytes, b_full(bs));
            bulk_g2s(slot + a.st_poff[s] + a.st_pplane[s] + (uint32_t)h * bytes, src + a.st_boff[s] + a.st_bplane[s], bytes, b_full(bs));
          }
        }
        MB2_TRACE(1, pit);
      }
      __syncwarp();
      if (++bs == a.b_slots) { bs = 0; ph ^= 1; }
    }
  } else if (warp == kLoadAWarp) {
    // ============================== weight loader: pre-swizzled (unit, stage) slabs, hi | lo ========================
    if (a.a_resident) {
      for (int iu = 0; iu < a.n_img_units; ++iu)
        for (int s = 0; s < a.k_stages; ++s) {
          const int slot = iu * a.k_stages + s;
          if (elect_one()) {
            mbar_arrive_expect_tx(a_full(slot), 2u * a.st_aplane[s]);
            bulk_g2s(a_res_addr(iu, s), a.Wimg + (size_t)iu * a.img_unit_bytes + a.st_aoff[s], 2u * a.st_aplane[s], a_full(slot));
          }
          __syncwarp();
        }
    } else {
      int slot = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += 2 * gridDim.x)            // one pass over the slabs per PAIR of tiles
        for (int u = 0; u < a.n_units; ++u)
          for (int s = 0; s < a.k_stages; ++s) {
            mbar_wait_relaxed(a_empty(slot), ph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(a_full(slot), 2u * a.st_aplane[s]);
              bulk_g2s(a_ring + (uint32_t)slot * a.a_slot_bytes, a.Wimg + (size_t)u * a.img_unit_bytes + a.st_aoff[s], 2u * a.st_aplane[s], a_full(slot));
            }
            __syncwarp();
            if (++slot == a.a_slots) { slot = 0; ph ^= 1; }
          }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ======================================================================
    // The WHOLE warp runs this loop with warp-uniform values (barrier waits, descriptor arithmetic: counters instead of runtime
    // divisions so that ptxas keeps them on the uniform datapath), and the tcgen05 instructions are issued by one ELECTED lane:
    // a stage's MMAs are then consecutive UTCHMMA instructions.  Guarded by `lane == 0` each of them sat in a waterfall loop and
    // the issuer needed ~200 cycles per MMA (profiles/r02 timeline: 15 MMAs = 3 k cycles per unit, the kernel's critical path).
    {
      const uint32_t idesc = make_idesc_mn(128u, 2u * (uint32_t)a.n_mma);          // both tiles of the pair
      uint32_t pit = 0, seq = 0, bph = 0, aph = 0, tph = 0; int bs = 0, aslot = 0, buf = 0;
      const uint32_t buf_cols = 2u * (uint32_t)a.n_mma;                        // one accumulator = one unit of a pair of tiles
      for (int tile = blockIdx.x; tile < total_tiles; tile += 2 * gridDim.x, ++pit) {
        mbar_wait(b_full(bs), bph);
        if (elect_one()) MB2_TRACE(2, pit);
        const int rot = pit & 3;
        const uint32_t sb0 = b_ring + (uint32_t)bs * a.b_pair_bytes;
        for (int u = 0; u < a.n_units; ++u, ++seq) {
          mbar_wait(t_empty(buf), tph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)buf * buf_cols;
          if (elect_one()) MB2_TRACE(8, seq);
          const int iu = a.rot_mode == 2 ? rot : u;
          for (int s = 0; s < a.k_stages; ++s) {
            uint32_t sa;
            if (a.a_resident) { mbar_wait(a_full(iu * a.k_stages + s), 0); sa = a_res_addr(iu, s); }
            else { mbar_wait(a_full(aslot), aph); sa = a_ring + (uint32_t)aslot * a.a_slot_bytes; }
            tc_fence_after();
            const uint32_t sb = sb0 + a.st_poff[s], rb = a.st_rb[s];
            const uint64_t d_whi = make_desc_rb(sa, rb), d_wlo = make_desc_rb(sa + a.st_aplane[s], rb);
            const uint64_t d_xhi = make_desc_rb(sb, rb), d_xlo = make_desc_rb(sb + a.st_pplane[s], rb);
            const uint32_t nk = a.st_ksteps[s];
            if (elect_one()) {
              if (!(a.dbg & 2)) {
#pragma unroll
                for (uint32_t kk = 0; kk < 4; ++kk) {
                  if (kk < nk) {
                    const uint64_t adv = (uint64_t)(kk * 2);           // +32 bytes (>> 4) inside the swizzle row
                    umma(d_tmem, d_whi + adv, d_xhi + adv, idesc, (s | (int)kk) != 0);
                    umma(d_tmem, d_wlo + adv, d_xhi + adv, idesc, 1);
                    umma(d_tmem, d_whi + adv, d_xlo + adv, idesc, 1);
                  }
                }
              }
              if (!a.a_resident) umma_commit(a_empty(aslot));
            }
            __syncwarp();
            if (!a.a_resident && ++aslot == a.a_slots) { aslot = 0; aph ^= 1; }
          }
          if (elect_one()) { MB2_TRACE(9, seq); umma_commit(t_full(buf)); MB2_TRACE(3, seq); }
          __syncwarp();
          if (++buf == a.n_bufs) { buf = 0; tph ^= 1; }
        }
        if (elect_one()) umma_commit(b_empty(bs));
        __syncwarp();
        if (++bs == a.b_slots) { bs = 0; bph ^= 1; }
      }
    }
  } else {
    // ============================== epilogue groups (3 x 4 warps): thread = one expanded channel ====================
    // group g takes the pair-units whose sequence number has parity g & 1, and of those tile g >> 1 of the pair; the
    // accumulators rotate over n_bufs TMEM buffers (seq % n_bufs, tracked with counters: no runtime division)
    const int g = warp >> 2, quarter = warp & 3;
    const int gpar = g & 1, half = g >> 1;
    int gbuf = gpar; uint32_t par = 0;                              // buffer / barrier parity of this group's NEXT pair-unit
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * a.n_mma);
    uint32_t pit = 0;
    for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += 2 * gridDim.x, ++pit) {
      const int tile = tile0 + half * (int)gridDim.x;
      const bool have = tile < total_tiles;                        // an odd tail: the pair's second tile does not exist
      const int b = tile / tiles_per_chunk, tt = tile - b * tiles_per_chunk;
      const int ty = tt / a.tiles_w, tx = tt - ty * a.tiles_w;
      const int ho0 = ty * a.TH, wo0 = tx * TW;
      const int hi0 = ho0 * S - 1, wi0 = wo0 * S - 1;
      const bool left_oob = wi0 < 0, right_oob = wi0 + PW - 1 >= a.W;
      const int rot = pit & 3;
      for (int u = 0; u < a.n_units; ++u) {
        const uint32_t seq = pit * a.n_units + u;
        if ((int)(seq & 1) != gpar) continue;
        const int mybuf = gbuf; const uint32_t mypar = par;       // this item's buffer; advance the counters for the next one
        gbuf += 2; if (gbuf >= a.n_bufs) { gbuf -= a.n_bufs; par ^= 1; }
        const uint32_t tbuf = tlane + (uint32_t)mybuf * (2u * (uint32_t)a.n_mma);
        int qeff = quarter;
        bool warp_active;
        const int rows_u = min(128, a.C - u * 128);
        if (a.rot_mode == 1 && u == a.n_units - 1) { warp_active = quarter == rot; qeff = 0; }
        else if (a.rot_mode == 2) { qeff = (quarter - rot) & 3; warp_active = qeff * 32 < rows_u; }
        else warp_active = qeff * 32 < rows_u;
        if (!warp_active || !have) {                              // nothing for this warp: handshake only
          mbar_wait(t_full(mybuf), mypar);
          mbar_arrive(t_empty(mybuf));
          continue;
        }
        const int c = u * 128 + qeff * 32 + lane;
        const bool active = c < a.C;
        float wd[9], bd = 0.f, be = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) wd[t] = active ? __ldg(a.w_dw + (size_t)t * a.C + c) : 0.f;
        if (active) { bd = __ldg(a.bias_dw + c); be = __ldg(a.bias_e + c); }
        mbar_wait(t_full(mybuf), mypar);
        tc_fence_after();
        if (lane == 0 && quarter == 0 && half == 0) MB2_TRACE(4, seq);

        float lsum = 0.f;
        float rA[PW], rB[PW], rC[PW];
        auto ld = [&](float (&dst)[PW], int r) {
          const int hi = hi0 + r;
          load_row<PW, POLY>(dst, tbuf + (uint32_t)(r * PW), hi >= 0 && hi < a.H, left_oob, right_oob, be, a.dbg);
          if (r == a.PH - 1) { tc_fence_before(); mbar_arrive(t_empty(mybuf)); if (lane == 0 && quarter == 0 && half == 0) MB2_TRACE(5, seq); }     // accumulator fully read: hand the buffer back
        };
        // RowTiles image of the result: row m = (b*Ho + ho)*Wo + wo, 64-channel stage c >> 6, chunk (c >> 3) & 7
        const uint32_t m_tile0 = ((uint32_t)b * a.Ho + ho0) * a.Wo + wo0;
        uint8_t* const out_c = a.d_img + (size_t)(c >> 6) * 32768u + (size_t)(c & 7) * 2u;
        const uint32_t chunk16 = (uint32_t)((c >> 3) & 7) << 4;
        auto out = [&](const float (&r0)[PW], const float (&r1)[PW], const float (&r2)[PW], int oh) {
          const uint32_t m = m_tile0 + (uint32_t)oh * a.Wo;
          out_row<S, TW, PW, POLY>(r0, r1, r2, wd, bd, out_c + (size_t)(m >> 7) * a.d_tile_bytes + (m & 127u) * 128u, chunk16, m & 7u, active && !(a.dbg & 4), lsum);
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
        if (lane == 0 && quarter == 0 && half == 0) MB2_TRACE(6, seq);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) MB2_TRACE(7, 0);
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
  P.TW = stride == 1 ? ((Wo % 8 == 0) ? 8 : 0) : ((Wo % 4 == 0) ? 4 : 0);
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
  uint32_t poff = 0;
  for (int s = 0; s < ns; ++s) {
    P.st_pplane[s] = 2u * (uint32_t)P.n_mma * (uint32_t)P.st_rb[s];           // multiple of 1024 (n_mma % 16 == 0): swizzle atoms stay aligned
    P.st_poff[s] = poff; poff += 2 * P.st_pplane[s];
  }
  P.b_pair_bytes = poff;
  // shared-memory plan: all weight slabs resident when they fit beside a double-buffered patch, else a streaming ring
  const size_t budget = kMb2SmemLimit - 1024 /*alignment*/ - 512 /*barriers*/;
  const bool ragged = (C % 128) != 0;
  const size_t plain_bytes = (size_t)P.n_units * P.img_unit_bytes;           // resident slabs are packed
  const size_t pair = P.b_pair_bytes;
  if (P.n_units == 1 && ragged && 4 * plain_bytes + 2 * pair <= budget) {
    P.a_resident = 1; P.a_slots = 4 * ns; P.b_slots = 2; P.a_region_bytes = (uint32_t)(4 * plain_bytes);   // single ragged unit: four rotated versions
  } else if (plain_bytes + 2 * pair <= budget) {
    P.a_resident = 1; P.a_slots = P.n_units * ns; P.b_slots = 2; P.a_region_bytes = (uint32_t)plain_bytes;
  } else {
    P.a_resident = 0;
    P.b_slots = 2;
    size_t left = budget > 2 * pair ? budget - 2 * pair : 0;
    int slots = (int)(left / P.a_slot_bytes);
    if (slots < std::max(2, ns + 1)) { P.b_slots = 1; left = budget > pair ? budget - pair : 0; slots = (int)(left / P.a_slot_bytes); }
    if (slots < 2) return P;
    P.a_slots = std::min(slots, 6);
    P.a_region_bytes = (uint32_t)P.a_slots * P.a_slot_bytes;
  }
  P.smem_bytes = 1024 + (size_t)P.a_region_bytes + (size_t)P.b_slots * P.b_pair_bytes + 16 * (size_t)(P.a_slots + P.b_slots + kGroups) + 64;
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

PatchTiles mb2_patch_layout(const Mb2Plan& P, int H, int W) {
  PatchTiles t;
  t.H = H; t.W = W; t.C = P.Cin; t.S = P.S; t.TH = P.TH; t.TW = P.TW; t.PH = P.PH; t.PW = P.PW;
  t.tiles_h = P.tiles_h; t.tiles_w = P.tiles_w;
  auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  if ((W & (W - 1)) || (P.TW & (P.TW - 1))) throw std::runtime_error("mbconv2: map width and tile width must be powers of two");
  t.log2W = ilog2(W); t.log2TW = ilog2(P.TW);
  t.n_stages = P.k_stages;
  for (int s = 0; s < P.k_stages; ++s) { t.st_rb[s] = P.st_rb[s]; t.st_k0[s] = P.st_k0[s]; t.st_off[s] = P.st_boff[s]; t.st_plane[s] = P.st_bplane[s]; }
  t.tile_bytes = P.b_slot_bytes;
  const uint32_t T = (uint32_t)(P.TH * P.S);
  t.magic_t = (65536u + T - 1) / T;
  t.hw = (uint32_t)(H * W);
  t.magic_hw = (uint32_t)((1ull << 32) / t.hw);
  return t;
}

void mb2_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<2, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<1, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
  BNB_CUDA(cudaFuncSetAttribute(mbconv2_kernel<2, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMb2SmemLimit));
}

void launch_mbconv2(const Mb2Plan& P, const Mb2Launch& L, cudaStream_t s, LaunchCounter& lc) {
  if (!P.ok) throw std::runtime_error("mbconv2: layer has no plan");
  Mb2Args a{};
  a.Wimg = L.Wimg; a.bias_e = L.bias_e; a.w_dw = L.w_dw; a.bias_dw = L.bias_dw; a.x_img = L.x_img; a.d_img = L.d_img;
  a.d_tile_bytes = RowTiles::make(P.C).tile_bytes; a.partial = L.partial;
  a.B = L.B; a.H = L.H; a.W = L.W; a.C = P.C; a.Ho = L.Ho; a.Wo = L.Wo;
  a.TH = P.TH; a.PH = P.PH; a.n_mma = P.n_mma; a.tiles_h = P.tiles_h; a.tiles_w = P.tiles_w;
  { static const int n_forced = getenv("BNB_MB2_NMMA") ? atoi(getenv("BNB_MB2_NMMA")) : 0; if (n_forced > 0) a.n_mma = n_forced; }   // timing experiment (wrong results)
  a.n_units = P.n_units; a.k_stages = P.k_stages; a.rot_mode = rot_mode_of(P);
  a.a_slots = P.a_slots; a.b_slots = P.b_slots; a.a_resident = P.a_resident; a.n_img_units = a.rot_mode == 2 ? 4 : P.n_units;
  { static const int dbg = getenv("BNB_MB2_DBG") ? atoi(getenv("BNB_MB2_DBG")) : 0; a.dbg = dbg; }
  a.a_slot_bytes = P.a_slot_bytes; a.a_region_bytes = P.a_region_bytes; a.b_slot_bytes = P.b_slot_bytes; a.img_unit_bytes = P.img_unit_bytes; a.b_pair_bytes = P.b_pair_bytes;
  a.n_bufs = std::max(2, std::min(4, kTmemCols / (2 * P.n_mma)));
  { static const int nb_forced = getenv("BNB_MB2_NBUFS") ? atoi(getenv("BNB_MB2_NBUFS")) : 0; if (nb_forced >= 2) a.n_bufs = std::min(a.n_bufs, nb_forced); }
  for (int st = 0; st < P.k_stages; ++st) {
    a.st_rb[st] = P.st_rb[st]; a.st_ksteps[st] = P.st_ksteps[st]; a.st_k0[st] = P.st_k0[st]; a.st_aoff[st] = P.st_aoff[st];
    a.st_aplane[st] = P.st_aplane[st]; a.st_boff[st] = P.st_boff[st]; a.st_bplane[st] = P.st_bplane[st];
    a.st_poff[st] = P.st_poff[st]; a.st_pplane[st] = P.st_pplane[st];
  }
  const long long tiles = (long long)L.B * P.tiles_h * P.tiles_w;
  const int grid = tiles < kNumSMs ? (int)tiles : kNumSMs;
  // debug timeline: BNB_MB2_TRACE=<file> BNB_MB2_TRACE_IDX=<n-th mbconv2 launch of the process>
  static std::atomic<long long> launch_idx{0};
  static const char* trace_path = getenv("BNB_MB2_TRACE");
  static const long long trace_idx = getenv("BNB_MB2_TRACE_IDX") ? atoll(getenv("BNB_MB2_TRACE_IDX")) : 0;
  const long long my_idx = launch_idx.fetch_add(1);
  long long* trace = nullptr;
  if (trace_path && my_idx == trace_idx) { BNB_CUDA(cudaMalloc(&trace, 2 * 10 * 64 * sizeof(long long))); BNB_CUDA(cudaMemsetAsync(trace, 0, 2 * 10 * 64 * sizeof(long long), s)); a.trace = trace; }
  // BNB_MB2_POLY=1: one of every four SiLU exponentials on the FMA pipe instead of MUFU (tc_common.cuh: ex2_poly).  Experiment
  // knob, default off: ncu shows the XU pipe 78 % busy, yet moving exponentials over made the kernel SLOWER (r02 run k20:
  // pw_expand 1.453 ms -> 1.514 with one of four, 1.608 with two of four) — the epilogue is issue-bound, not MUFU-bound.
  static const int poly = getenv("BNB_MB2_POLY") ? atoi(getenv("BNB_MB2_POLY")) : 0;
  const bool s1 = P.S == 1 && P.TW == 8, s2 = P.S == 2 && P.TW == 4;
  if (!s1 && !s2) throw std::runtime_error("mbconv2: unsupported tile shape");
  if (poly <= 0) { if (s1) launch_k(mbconv2_kernel<1, 8, 0>, dim3(grid), dim3(kThreads), P.smem_bytes, s, a); else launch_k(mbconv2_kernel<2, 4, 0>, dim3(grid), dim3(kThreads), P.smem_bytes, s, a); }
  else { if (s1) launch_k(mbconv2_kernel<1, 8, 1>, dim3(grid), dim3(kThreads), P.smem_bytes, s, a); else launch_k(mbconv2_kernel<2, 4, 1>, dim3(grid), dim3(kThreads), P.smem_bytes, s, a); }
  if (trace) {
    std::vector<long long> h(2 * 10 * 64);
    BNB_CUDA(cudaStreamSynchronize(s));
    BNB_CUDA(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "w");
    if (f) {
      fprintf(f, "# mbconv2 B=%d H=%d W=%d Cin=%d C=%d S=%d tile=%dx%d patch=%dx%d n_mma=%d units=%d k_stages=%d resident=%d a_slots=%d b_slots=%d grid=%d tiles=%lld\n",
              L.B, L.H, L.W, P.Cin, P.C, P.S, P.TH, P.TW, P.PH, P.PW, P.n_mma, P.n_units, P.k_stages, P.a_resident, P.a_slots, P.b_slots, grid, tiles);
      fprintf(f, "# cta idx start patch_issue patch_landed mma_committed acc_seen acc_released unit_done end mma_first_issue mma_last_issue   (cycles since the CTA's start)\n");
      for (int c = 0; c < 2; ++c) {
        const long long t0 = h[(c * 10 + 0) * 64];
        for (int i = 0; i < 64; ++i) {
          bool any = false;
          for (int e = 0; e < 10; ++e) any = any || h[(c * 10 + e) * 64 + i] != 0;
          if (!any) continue;
          fprintf(f, "%d %d", c, i);
          for (int e = 0; e < 10; ++e) { const long long v = h[(c * 10 + e) * 64 + i]; fprintf(f, " %lld", v ? v - t0 : -1); }
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
