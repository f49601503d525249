// mbconv_tc.cu — the front half of an MBConv block in ONE kernel, sm_100a only:
//     1x1 expand (tcgen05.mma, fp16 hi/lo 3-term split, fp32 accumulate in TMEM) + bias + SiLU
//  -> 3x3 depthwise (+bias +SiLU, stride 1 or 2, zero padding 1) -> per-channel row sums for squeeze-excite.
//
// Replaces the CONV_2D(1x1) -> LOGISTIC/MUL -> [PAD] -> DEPTHWISE_CONV_2D -> LOGISTIC/MUL -> MEAN op groups the
// reference executes inside TFLite (/root/reference/internal/inference/tflite/classifier.go:107).
// The expanded tensor (4-8x the block input, 9.4 MB per chunk over the network, the largest HBM/L2 consumer of
// the unfused chain) is never written: it lives in TMEM, then in an 18 KB shared-memory tile, then dies.
//
// Work decomposition
//   tile   = one chunk x a TH x TW patch of depthwise OUTPUT pixels; its input patch with halo
//            PH x PW = ((TH-1)s+3) x ((TW-1)s+3) <= 128 positions = the 128 rows (TMEM lanes) of the expand GEMM.
//            TH x TW comes from a cost model of the epilogue (mbconv_geometry), evaluated for the classifier's nominal
//            launch size so that results never depend on the batch composition.
//   A      = block input [B][H][W][Cin] fp32, fetched by ONE 4-D TMA box per 64 input channels (out-of-image
//            coordinates arrive as zeros), converted once per tile by 4 warps to fp16 hi/lo 128B-swizzled K-major tiles.
//   units  = the expanded channels in groups of 96 = three 32-channel slices: ONE MMA group (N = 96) per unit into one
//            of two 128-column TMEM buffers (N = 32 MMAs cost the same ~90 cycles each and starved the epilogue);
//            B (the pre-split weight image of pw_tc.cu) streams per unit via cp.async.bulk, two units in flight.
//   epilogue groups (3 x 4 warps): group g owns slice g of every unit: tcgen05.ld -> +bias -> SiLU -> zero outside the
//            image (padding is zero in the EXPANDED domain) -> fp32 smem tile [128 pos][36] (STS.128) -> group barrier
//            -> depthwise with lane = channel, four outputs per step from a 3 x (3s+3) register window read at immediate
//            offsets, coalesced 128 B stores of the output, per-tile SE sums in a fixed order (deterministic).
// How each of these choices was measured into place: profiles/README.md (optimisation log of this kernel).
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <type_traits>
#include <vector>
#include <mutex>
#include <tuple>

#include "kernels.h"
#include "mbconv_tc.h"
#include "tc_common.cuh"

namespace bnb {

namespace {

using namespace tc;

constexpr int kGroups = 3;          // epilogue groups of 4 warps; slice j belongs to group j % kGroups (the CUDA-core work of the
                                    // epilogue, not the MMA, bounds this kernel: 12 of the 19 warps do it)
constexpr int kThreads = 608;       // 19 warps: 12 epilogue (3 groups), MMA, A loader, 4 converters, B loader
constexpr int kMmaWarp = 12, kLoadAWarp = 13, kConvWarp0 = 14, kConvThreads = 128, kLoadBWarp = 18;
constexpr int kTmemCols = 256;      // kGroups x 2 accumulators x 32 columns, rounded to a power of two
constexpr int kRowsPerPass = kConvThreads / 8, kPasses = 128 / kRowsPerPass;
constexpr int kSliceCols = 32;      // expanded channels per slice
constexpr int kABytes = 32768;      // one A stage: raw fp32 [128][64] -> in place hi | lo fp16 tiles (16 KB each)
constexpr int kUnitCols = kGroups * kSliceCols;   // one MMA unit: a slice for every group (N = 96)
constexpr int kBufCols = 128;       // TMEM columns per accumulator buffer (unit rounded up)
constexpr int kBStage = 2 * kUnitCols * 128;     // one (unit, k-stage) weight slab: hi 96 x 128 B | lo 96 x 128 B (24 KB)
constexpr int kSPitch = 36;         // floats per position row (144 B): 16 B-aligned rows for STS.128 by the writers (lane = pos, 8 lanes cover all banks);
                                    // readers (lane = channel) hit consecutive banks
constexpr int kSBytes = 128 * kSPitch * 4 + 1024;  // expanded tile of one group: [128 pos][36] fp32 + 1 KB slack that the depthwise window may over-read

#define MB_TRACE(ev, iter) do { if (a.trace && blockIdx.x == 0 && (iter) < 64) a.trace[(ev) * 64 + (iter)] = clock64(); } while (0)

__device__ __forceinline__ float silu1(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__global__ void __launch_bounds__(kThreads, 1)
mbconv_tc_kernel(const MbArgs a, const __grid_constant__ CUtensorMap x_map) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up ---------------------------------------------------------------------------------
  const uint32_t a_ring = base;                                               // a_slots x 32 KB
  const uint32_t b_slot_bytes = (uint32_t)a.k_stages * kBStage;
  const uint32_t b_ring = a_ring + (uint32_t)a.a_slots * kABytes;             // b_slots x k_stages x 8 KB
  const uint32_t s_tiles = b_ring + (uint32_t)a.b_slots * b_slot_bytes;       // 2 x 16 KB
  const uint32_t s_red = s_tiles + kGroups * kSBytes;                         // kGroups x 4 warps x 32 floats
  const uint32_t bars = s_red + 2048;
  auto a_raw = [&](int s) { return bars + 8u * s; };
  auto a_full = [&](int s) { return bars + 8u * (a.a_slots + s); };
  auto a_empty = [&](int s) { return bars + 8u * (2 * a.a_slots + s); };
  const uint32_t bb = bars + 8u * (3 * a.a_slots);
  auto b_full = [&](int s) { return bb + 8u * s; };
  auto b_empty = [&](int s) { return bb + 8u * (a.b_slots + s); };
  const uint32_t tb = bb + 8u * (2 * a.b_slots);
  auto t_full = [&](int buf) { return tb + 8u * buf; };
  auto t_empty = [&](int buf) { return tb + 8u * (2 + buf); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (tb - base) + 8 * 4 * kGroups);

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.a_slots; ++s) { mbar_init(a_raw(s), 1); mbar_init(a_full(s), kConvThreads); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < a.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(t_full(b), 1); mbar_init(t_empty(b), kGroups * 128); }
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
  const int n_slices = (a.n_pad + kSliceCols - 1) / kSliceCols;
  const int n_units = (n_slices + kGroups - 1) / kGroups;
  const int P = a.ph * a.pw;
  const int pitch = a.box_c * 4;

  if (warp >= kConvWarp0 && warp < kLoadBWarp) {
    // ============================== A converters (4 warps): raw fp32 patch -> fp16 hi / lo tiles, in place ==
    const int pt = threadIdx.x - kConvWarp0 * 32;
    const int c = pt & 7, r0 = pt >> 3;
    uint32_t ia = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int ks = 0; ks < a.k_stages; ++ks, ++ia) {
        const int s = ia % a.a_slots; const uint32_t ph = (ia / a.a_slots) & 1;
        uint8_t* buf = base_ptr + (a_ring - base) + (size_t)s * kABytes;
        const int k = ks * 64 + c * 8;
        const bool k_live = k < a.k_pad;
        float v[kPasses][8];
        mbar_wait_relaxed(a_raw(s), ph);
        if (k_live) {
#pragma unroll
          for (int q = 0; q < kPasses; ++q) {
            const int r = r0 + kRowsPerPass * q;
            if (r < P && k < a.Cin) {
              const float4 x0 = *reinterpret_cast<const float4*>(buf + r * pitch + c * 32);
              float4 x1 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k + 4 < a.Cin) x1 = *reinterpret_cast<const float4*>(buf + r * pitch + c * 32 + 16);
              v[q][0] = x0.x; v[q][1] = x0.y; v[q][2] = x0.z; v[q][3] = x0.w;
              v[q][4] = x1.x; v[q][5] = x1.y; v[q][6] = x1.z; v[q][7] = x1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[q][i] = 0.f;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");     // all raw reads done before the in-place overwrite
        if (k_live) {
#pragma unroll
          for (int q = 0; q < kPasses; ++q) {
            const int r = r0 + kRowsPerPass * q;
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_pair(v[q][2 * i], v[q][2 * i + 1], hw[i], lw[i]);
            const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(buf + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(buf + 16384 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(a_full(s));
      }
    }
  } else if (warp == kLoadAWarp) {
    // ============================== A loader: one 4-D TMA box per (tile, 64 input channels) ==================
    if (lane == 0) {
      uint32_t ia = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_chunk, tt = tile - b * tiles_per_chunk;
        const int ty = tt / a.tiles_w, tx = tt - ty * a.tiles_w;
        const int hi0 = ty * a.th * a.stride - 1, wi0 = tx * a.tw * a.stride - 1;
        for (int ks = 0; ks < a.k_stages; ++ks, ++ia) {
          const int s = ia % a.a_slots; const uint32_t ph = (ia / a.a_slots) & 1;
          mbar_wait_relaxed(a_empty(s), ph ^ 1);
          mbar_arrive_expect_tx(a_raw(s), (uint32_t)(a.box_c * 4 * P));
          tma_load_4d(a_ring + (uint32_t)s * kABytes, &x_map, ks * 64, wi0, hi0, b, a_raw(s));
        }
      }
    }
  } else if (warp == kLoadBWarp) {
    // ============================== B loader: weight slabs of one unit (kGroups slices = 96 channels), all k-stages ==
    if (lane == 0) {
      uint32_t ib = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int u = 0; u < n_units; ++u, ++ib) {
          const int s = ib % a.b_slots; const uint32_t ph = (ib / a.b_slots) & 1;
          const int rows = min(kUnitCols, a.n_pad - u * kUnitCols);
          mbar_wait_relaxed(b_empty(s), ph ^ 1);
          MB_TRACE(0, ib);
          mbar_arrive_expect_tx(b_full(s), (uint32_t)(a.k_stages * 2 * rows * 128));
          for (int ks = 0; ks < a.k_stages; ++ks) {
            const uint8_t* src = a.Wimg + ((size_t)ks * 2) * (size_t)a.n_pad * 128 + (size_t)u * kUnitCols * 128;
            const uint32_t dst = b_ring + (uint32_t)s * b_slot_bytes + (uint32_t)ks * kBStage;
            bulk_g2s(dst, src, (uint32_t)rows * 128u, b_full(s));
            bulk_g2s(dst + kBStage / 2, src + (size_t)a.n_pad * 128, (uint32_t)rows * 128u, b_full(s));
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer ================================================================
    if (lane == 0) {
      // One MMA covers a unit = kGroups adjacent slices (N = 96): with N = 32 the instruction rate was bound by re-reading
      // the A tile from shared memory for every slice (~93 cycles per MMA whatever N).  Two units (the two TMEM buffers) are
      // issued interleaved k-step by k-step so that consecutive MMAs never accumulate into the same columns.
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < a.k_stages; ++ks) {                 // the whole converted patch must be in place
          const uint32_t i = ia + ks;
          mbar_wait(a_full(i % a.a_slots), (i / a.a_slots) & 1);
        }
        const int batch = min(2, a.b_slots);
        for (int u0 = 0; u0 < n_units; u0 += batch) {
          const int nb = min(batch, n_units - u0);
          uint32_t d_tmem[2], idesc[2], bslot[2];
          int bufs[2];
          for (int v = 0; v < nb; ++v) {
            const uint32_t seq = ib + v;
            const int ncols = min(kUnitCols, a.n_pad - (u0 + v) * kUnitCols);
            bufs[v] = seq & 1; bslot[v] = seq % a.b_slots;
            idesc[v] = (1u << 4) | ((uint32_t)(ncols >> 3) << 17) | (8u << 24);    // f32 accum, f16 x f16, M = 128
            d_tmem[v] = tmem_base + (uint32_t)bufs[v] * kBufCols;
            mbar_wait(t_empty(bufs[v]), ((seq >> 1) & 1) ^ 1);
            if (v == 0) MB_TRACE(1, ib);
            mbar_wait(b_full(bslot[v]), (seq / a.b_slots) & 1);
          }
          MB_TRACE(2, ib);
          tc_fence_after();
          for (int ks = 0; ks < a.k_stages; ++ks) {
            const uint32_t sa = a_ring + (uint32_t)((ia + ks) % a.a_slots) * kABytes;
            const uint64_t d_ahi = make_desc(sa), d_alo = make_desc(sa + 16384);
            const int kk_n = min(64, a.k_pad - ks * 64) / 16;
            for (int kk = 0; kk < kk_n; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);
              for (int v = 0; v < nb; ++v) {
                const uint32_t sbk = b_ring + bslot[v] * b_slot_bytes + (uint32_t)ks * kBStage;
                const uint64_t d_bhi = make_desc(sbk), d_blo = make_desc(sbk + kBStage / 2);
                umma(d_tmem[v], d_ahi + adv, d_bhi + adv, idesc[v], (ks | kk) != 0);
                umma(d_tmem[v], d_alo + adv, d_bhi + adv, idesc[v], 1);
                umma(d_tmem[v], d_ahi + adv, d_blo + adv, idesc[v], 1);
              }
            }
          }
          for (int v = 0; v < nb; ++v) { umma_commit(b_empty(bslot[v])); umma_commit(t_full(bufs[v])); }
          MB_TRACE(3, ib);
          ib += nb;
        }
        for (int ks = 0; ks < a.k_stages; ++ks) umma_commit(a_empty((ia + ks) % a.a_slots));   // patch consumed by every slice
        ia += a.k_stages;
      }
    }
  } else {
    // ============================== epilogue groups (3 x 4 warps) ==============================================
    const int g = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;                                        // patch position = TMEM lane of this thread
    const int prow = row / a.pw, pcol = row - prow * a.pw;
    uint8_t* S = base_ptr + (s_tiles - base) + (size_t)g * kSBytes;
    float* red = reinterpret_cast<float*>(base_ptr + (s_red - base)) + g * 128;
    const int n_out = a.th * a.tw;
    uint32_t sgl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_chunk, tt = tile - b * tiles_per_chunk;
      const int ty = tt / a.tiles_w, tx = tt - ty * a.tiles_w;
      const int ho0 = ty * a.th, wo0 = tx * a.tw;
      const int hi = ho0 * a.stride - 1 + prow, wi = wo0 * a.stride - 1 + pcol;
      const bool inside = row < P && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;   // padding is ZERO in the expanded domain
      for (int u = 0; u < n_units; ++u, ++sgl) {
        const int buf = sgl & 1;
        const int j = u * kGroups + g;
        if (j >= n_slices) {                                               // ragged last unit: handshake only
          mbar_wait(t_full(buf), (sgl >> 1) & 1);
          mbar_arrive(t_empty(buf));
          continue;
        }
        const int ch0 = j * kSliceCols;
        // depthwise taps + both biases of this slice (issued before the accumulator wait: latency hidden)
        const bool ch_ok = ch0 + lane < a.C;
        float wd[9], bd = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) wd[t] = ch_ok ? __ldg(a.w_dw + (size_t)t * a.C + ch0 + lane) : 0.f;
        if (ch_ok) bd = __ldg(a.bias_dw + ch0 + lane);
        const float bev = __ldg(a.bias_e + ch0 + lane);                   // lane i holds the expand bias of column i (padded at upload)
        mbar_wait(t_full(buf), (sgl >> 1) & 1);
        if (threadIdx.x == 0) MB_TRACE(4, sgl);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kBufCols + (uint32_t)g * kSliceCols, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(t_empty(buf));                                        // accumulator is in registers: free it early
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float4 o;
          o.x = __uint_as_float(r[4 * i4 + 0]) + __shfl_sync(0xffffffffu, bev, 4 * i4 + 0);
          o.y = __uint_as_float(r[4 * i4 + 1]) + __shfl_sync(0xffffffffu, bev, 4 * i4 + 1);
          o.z = __uint_as_float(r[4 * i4 + 2]) + __shfl_sync(0xffffffffu, bev, 4 * i4 + 2);
          o.w = __uint_as_float(r[4 * i4 + 3]) + __shfl_sync(0xffffffffu, bev, 4 * i4 + 3);
          silu2(o.x, o.y); silu2(o.z, o.w);
          if (!inside) o = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(S) + row * kSPitch + 4 * i4) = o;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");       // expanded tile of this group complete
        if (threadIdx.x == 0) MB_TRACE(5, sgl);
        // ---- depthwise 3x3 from shared memory: lane = channel; the tile's output rows are cut into segments that the
        // four warps take round-robin; a warp produces FOUR adjacent outputs at a time from one 3 x (3s+3) register window
        // (independent accumulator chains: the single-output version was latency-bound at ~260 cycles per pixel).
        float lsum = 0.f;
        const float* Sl = reinterpret_cast<const float*>(S) + lane;
        const int nseg = a.th >= 4 ? 1 : (a.th >= 2 ? 2 : 4);
        const int segw = (a.tw + nseg - 1) / nseg;
        const int wlim = min(a.tw, a.Wo - wo0);                            // output columns of this tile that exist
        auto dw_segments = [&](auto stride_tag) {
          constexpr int ST = decltype(stride_tag)::value;
          constexpr int NC = 3 * ST + 3;                                   // window columns for four outputs
          const int rowp = a.pw * kSPitch;
          for (int sidx = q; sidx < a.th * nseg; sidx += 4) {
            const int oh = sidx / nseg, ws = (sidx - oh * nseg) * segw, we = min(wlim, ws + segw);
            const int ho = ho0 + oh;
            if (ho >= a.Ho || ws >= we) continue;
            float* dp = a.D + (((size_t)b * a.Ho + ho) * a.Wo + wo0 + ws) * a.C + ch0 + lane;
            const float* rp = Sl + (oh * ST * a.pw + ws * ST) * kSPitch;   // window origin; columns at immediate offsets
            for (int nleft = we - ws; nleft > 0; nleft -= 4, rp += 4 * ST * kSPitch, dp += 4 * (size_t)a.C) {
              float xw[3][NC];
#pragma unroll
              for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) xw[kh][cc] = rp[kh * rowp + cc * kSPitch];   // may over-read past the patch: unused
              float acc[4] = {bd, bd, bd, bd};
#pragma unroll
              for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                  for (int i = 0; i < 4; ++i) acc[i] = fmaf(xw[kh][i * ST + kw], wd[kh * 3 + kw], acc[i]);
              silu2(acc[0], acc[1]); silu2(acc[2], acc[3]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const bool ok = ch_ok && i < nleft;
                if (ok) dp[i * (size_t)a.C] = acc[i];
                lsum += ok ? acc[i] : 0.f;
              }
            }
          }
        };
        if (a.stride == 1) dw_segments(std::integral_constant<int, 1>{}); else dw_segments(std::integral_constant<int, 2>{});
        if (a.partial != nullptr) {
          red[q * 32 + lane] = lsum;
          asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
          if (q == 0 && ch_ok)
            a.partial[((size_t)b * tiles_per_chunk + tt) * a.C + ch0 + lane] = (red[lane] + red[32 + lane]) + (red[64 + lane] + red[96 + lane]);
        }
        if (threadIdx.x == 0) MB_TRACE(6, sgl);
        asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");       // tile S and `red` may be overwritten now
        if (threadIdx.x == 0) MB_TRACE(7, sgl);
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

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
MbGeom mbconv_geometry(int H, int W, int Ho, int Wo, int stride, int Cin, int C, int B, int max_tiles) {
  MbGeom g{};
  // Tile search.  The kernel is bound by the CUDA-core work of its epilogue warps, so candidates are ranked by a small
  // cost model (cycles, calibrated on clock64 traces of the 12x32 and 24x64 blocks):
  //   slice = E phase (all 128 TMEM lanes whatever the patch size) + depthwise blocks of 4 outputs on the busiest warp
  //   tile  = units x slice + patch load/convert bubble;   launch = ceil(B x tiles / SMs) x tile
  // With B = 0 / C = 0 (unknown) this degenerates to "fewest tiles, then smallest patch".
  const int n_slices = C > 0 ? (C + kSliceCols - 1) / kSliceCols : kGroups;
  const int units = (n_slices + kGroups - 1) / kGroups;
  long long best_cost = -1; int best_tiles = 0;
  for (int th = 1; th <= Ho; ++th)
    for (int tw = 1; tw <= Wo; ++tw) {
      const int ph = (th - 1) * stride + 3, pw = (tw - 1) * stride + 3;
      if (ph * pw > 128 || pw > 256 || ph > 256) continue;
      const int tiles_h = (Ho + th - 1) / th, tiles_w = (Wo + tw - 1) / tw, tiles = tiles_h * tiles_w;
      if (max_tiles > 0 && tiles > max_tiles) continue;
      const int nseg = th >= 4 ? 1 : (th >= 2 ? 2 : 4), segw = (tw + nseg - 1) / nseg;
      const int blocks = ((th * nseg + 3) / 4) * ((segw + 3) / 4);
      const long long rounds = B > 0 ? ((long long)B * tiles + kNumSMs - 1) / kNumSMs : tiles;
      const long long cost = rounds * ((long long)units * (1950 + 700 * blocks) + 3000) * 1024 + ph * pw;
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && tiles < best_tiles)) {
        best_cost = cost; best_tiles = tiles;
        g.th = th; g.tw = tw; g.ph = ph; g.pw = pw; g.tiles_h = tiles_h; g.tiles_w = tiles_w;
      }
    }
  g.k_stages = (Cin + 63) / 64;
  g.box_c = Cin < 64 ? Cin : 64;
  // shared-memory plan: the patch ring (a tile uses k_stages slots; two tiles in flight only when k_stages == 1) and the
  // weight-unit ring (one slot = 96 channels x all k-stages; the MMA issuer interleaves min(2, b_slots) units)
  g.a_slots = g.k_stages == 1 ? 2 : g.k_stages;
  g.b_slots = g.k_stages <= 2 ? 2 : 1;
  g.smem_bytes = (size_t)g.a_slots * kABytes + (size_t)g.b_slots * g.k_stages * kBStage + kGroups * kSBytes + 2048 +
                 8 * (3 * (size_t)g.a_slots + 2 * (size_t)g.b_slots + 4 * kGroups) + 64 + 16 + 1024 /*alignment*/;
  (void)H; (void)W;
  return g;
}

static CUtensorMap encode_x_map(const float* x, int B, int H, int W, int C, int box_c, int pw, int ph) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::mutex mu;
  static std::map<std::tuple<const float*, int, int, int, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    BNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("mbconv_tc: cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeFn>(p);
  }
  auto key = std::make_tuple(x, B, H, W, C, pw, ph);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)pw, (cuuint32_t)ph, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("mbconv_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

void mbconv_tc_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(mbconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMbSmemLimit));
}

void launch_mbconv_tc(const MbLaunch& L, cudaStream_t s, LaunchCounter& lc) {
  const MbGeom g = mbconv_geometry(L.H, L.W, L.Ho, L.Wo, L.stride, L.Cin, L.C, L.B_nominal, L.max_tiles);
  if (g.th == 0) throw std::runtime_error("mbconv_tc: no tile geometry fits 128 patch positions");
  if (g.smem_bytes > kMbSmemLimit) throw std::runtime_error("mbconv_tc: shared memory budget exceeded");
  MbArgs a{};
  a.Wimg = L.Wimg; a.bias_e = L.bias_e; a.w_dw = L.w_dw; a.bias_dw = L.bias_dw; a.D = L.D; a.partial = L.partial;
  a.B = L.B; a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.C = L.C; a.Ho = L.Ho; a.Wo = L.Wo; a.stride = L.stride;
  a.th = g.th; a.tw = g.tw; a.ph = g.ph; a.pw = g.pw; a.tiles_h = g.tiles_h; a.tiles_w = g.tiles_w;
  a.n_pad = (L.C + 15) / 16 * 16; a.k_pad = (L.Cin + 15) / 16 * 16; a.k_stages = g.k_stages; a.box_c = g.box_c;
  a.a_slots = g.a_slots; a.b_slots = g.b_slots;
  const CUtensorMap xmap = encode_x_map(L.x, L.B, L.H, L.W, L.Cin, g.box_c, g.pw, g.ph);
  const long long tiles = (long long)L.B * g.tiles_h * g.tiles_w;
  const int grid = tiles < kNumSMs ? (int)tiles : kNumSMs;
  static std::atomic<long long> launch_idx{0};
  static const char* trace_path = getenv("BNB_MB_TRACE");
  static const long long trace_idx = getenv("BNB_MB_TRACE_IDX") ? atoll(getenv("BNB_MB_TRACE_IDX")) : 0;
  long long* trace = nullptr;
  const long long my_idx = launch_idx.fetch_add(1);
  if (trace_path && my_idx == trace_idx) { BNB_CUDA(cudaMalloc(&trace, 8 * 64 * sizeof(long long))); BNB_CUDA(cudaMemsetAsync(trace, 0, 8 * 64 * sizeof(long long), s)); a.trace = trace; }
  mbconv_tc_kernel<<<grid, kThreads, g.smem_bytes, s>>>(a, xmap);
  if (trace) {
    std::vector<long long> h(8 * 64);
    BNB_CUDA(cudaStreamSynchronize(s));
    BNB_CUDA(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "w");
    if (f) {
      fprintf(f, "# B=%d H=%d W=%d Cin=%d C=%d stride=%d tile=%dx%d patch=%dx%d tiles/chunk=%d grid=%d k_stages=%d\n# slice bload_issue mma_tmem_free mma_b_ready mma_committed epi_acc_ready epi_E_done epi_dw_done epi_slice_end\n",
              L.B, L.H, L.W, L.Cin, L.C, L.stride, g.th, g.tw, g.ph, g.pw, g.tiles_h * g.tiles_w, grid, g.k_stages);
      long long t0 = h[0];
      for (int i = 0; i < 64; ++i) { fprintf(f, "%d", i); for (int e = 0; e < 8; ++e) fprintf(f, " %lld", h[e * 64 + i] ? h[e * 64 + i] - t0 : -1); fprintf(f, "\n"); }
      fclose(f);
    }
    cudaFree(trace);
  }
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
