// layouts.h — operand images in HBM/L2, written by the producing kernel in EXACTLY the shared-memory form the consuming
// tcgen05 kernel needs, so that an operand tile travels with ONE cp.async.bulk instead of a TMA tensor box.
//
// Why (profiles/r02_*): a TMA tensor box is served row by row (~10 ns per 128-byte row per SM); the 128-row A tiles of the
// project GEMMs and the 144..153-row halo patches of the fused MBConv kernel were bound by that request rate, not by
// bytes or math.  A contiguous image is one bulk request.
//
// Two images, both holding fp16 hi and lo planes (x ~= hi + lo):
//   RowTiles  — a [M][K] matrix cut into 128-row tiles x 64-channel stages: per (tile, stage) [hi plane | lo plane], each
//               plane 128 rows x 128 B, 128-byte swizzled K-major (the A operand of pw2.cu; written by mbconv2.cu's
//               depthwise phase, the post-conv im2col kernel and the embedding kernel).
//   PatchTiles — an NHWC activation cut into the halo patches of the CONSUMING MBConv block's tiles (halo pixels are stored
//               once per tile that needs them): per (chunk, tile) the complete shared-memory slot of mbconv2.cu — every K
//               stage, [hi | lo], rows = patch positions, swizzled K-major with 128 / 64 / 32-byte rows (written by pw2.cu's
//               epilogue and by the stem kernel).
// Unwritten bytes of an image (channels past K inside the last stage, rows past the patch) are only ever multiplied by
// zero weights or discarded, so they merely have to be finite: buffers are zero-filled once at allocation and every
// kernel writes finite fp16 bit patterns.
#pragma once
#include <stdint.h>

#include <vector>

#ifndef __host__
#define __host__
#define __device__
#endif

namespace bnb {

constexpr int kPatchMaxStages = 3;

// byte offset of 16-byte chunk `chunk` of row `r` inside a K-major swizzled plane with `rb` bytes per row (128 / 64 / 32)
__host__ __device__ inline uint32_t lay_swz(uint32_t r, uint32_t chunk, uint32_t rb) {
  const uint32_t a = r * rb + chunk * 16u;
  const uint32_t mask = rb == 128 ? 7u : (rb == 64 ? 3u : 1u);
  return a ^ (((a >> 7) & mask) << 4);
}

// ---- RowTiles ----------------------------------------------------------------------------------------------------------
struct RowTiles {
  int K = 0, n_stages = 0;                         // stages of 64 channels, all 128-byte rows
  uint32_t tile_bytes = 0;                         // n_stages * 32 KB
  __host__ __device__ static RowTiles make(int K_) { RowTiles t; t.K = K_; t.n_stages = (K_ + 63) / 64; t.tile_bytes = (uint32_t)t.n_stages * 32768u; return t; }
  __host__ __device__ size_t bytes(long long M) const { return (size_t)((M + 127) / 128) * tile_bytes; }
  // offset of the 16-byte piece holding channels [8*c8, 8*c8+8) of row m in the hi plane; the lo plane is +16384
  __host__ __device__ size_t piece(long long m, int c8) const {
    return (size_t)(m >> 7) * tile_bytes + (size_t)(c8 >> 3) * 32768u + lay_swz((uint32_t)(m & 127), (uint32_t)(c8 & 7), 128u);
  }
};

// ---- PatchTiles --------------------------------------------------------------------------------------------------------
struct PatchTiles {
  int H = 0, W = 0, C = 0;                         // the activation: [B][H][W][C]
  int S = 1, TH = 0, TW = 0, PH = 0, PW = 0;       // consumer tiling: output tile TH x TW, stride S, patch PH x PW
  int tiles_h = 0, tiles_w = 0, log2W = 0, log2TW = 0;
  int n_stages = 0;
  int st_rb[kPatchMaxStages] = {0, 0, 0}, st_k0[kPatchMaxStages] = {0, 0, 0};
  uint32_t st_off[kPatchMaxStages] = {0, 0, 0}, st_plane[kPatchMaxStages] = {0, 0, 0};
  uint32_t tile_bytes = 0;
  uint32_t magic_t = 0;                            // ceil(65536 / (TH*S)): u / (TH*S) == (u * magic_t) >> 16 for u < 65536 / (TH*S)
  uint32_t hw = 0, magic_hw = 0;                   // H*W and floor(2^32 / (H*W))
  // Device lookup tables (null on the host side): where pixel p = h*W + w of a chunk goes.  dst_tbl[p] = up to four entries
  // (tile index inside the chunk << 8 | patch row n), 0xffffffff = unused; res_tbl[p] = the entry of the tile in which the
  // pixel is interior.  They replace for_each_tile / interior in the kernels' epilogues (scalar address arithmetic was the
  // critical path of the project GEMM's epilogue: profiles/r02 timelines).
  const uint4* dst_tbl = nullptr; const uint32_t* res_tbl = nullptr;

  __host__ __device__ size_t bytes(long long B) const { return (size_t)B * tiles_h * tiles_w * tile_bytes; }
  __host__ __device__ void split_pixel(uint32_t m, int* b, int* h, int* w) const {      // m = (b*H + h)*W + w
    uint32_t q = (uint32_t)(((uint64_t)m * magic_hw) >> 32);
    uint32_t r = m - q * hw;
    if (r >= hw) { ++q; r -= hw; }
    *b = (int)q; *h = (int)(r >> log2W); *w = (int)(r & (uint32_t)(W - 1));
  }
  __host__ __device__ void stage_of(int c8, int* s, int* chunk) const {                 // 8-channel piece -> (stage, 16-byte chunk)
    int st = 0;
    for (int i = 1; i < kPatchMaxStages; ++i) if (i < n_stages && 8 * c8 >= st_k0[i]) st = i;
    *s = st; *chunk = c8 - st_k0[st] / 8;
  }
  // offset of the piece inside ONE tile image (hi plane; lo = + st_plane[s]) for patch position (prow, pcol)
  __host__ __device__ uint32_t in_tile(int prow, int pcol, int s, int chunk) const {
    return st_off[s] + lay_swz((uint32_t)(prow * PW + pcol), (uint32_t)chunk, (uint32_t)st_rb[s]);
  }
  // offset of the hi-plane piece for a table entry (tile << 8 | n) of chunk b; the lo plane is + st_plane[s]
  __host__ __device__ size_t entry_piece(int b, uint32_t entry, int s, int chunk) const {
    return ((size_t)b * (size_t)(tiles_h * tiles_w) + (entry >> 8)) * tile_bytes + st_off[s] + lay_swz(entry & 255u, (uint32_t)chunk, (uint32_t)st_rb[s]);
  }
  __host__ __device__ void split_chunk(uint32_t m, int* b, uint32_t* pix) const {          // m = b*H*W + pix
    uint32_t q = (uint32_t)(((uint64_t)m * magic_hw) >> 32);
    uint32_t r = m - q * hw;
    if (r >= hw) { ++q; r -= hw; }
    *b = (int)q; *pix = r;
  }
  __host__ __device__ size_t tile_base(int b, int ty, int tx) const { return ((size_t)((size_t)b * tiles_h + ty) * tiles_w + tx) * tile_bytes; }
  // every tile that holds pixel (h, w): calls f(tile_y, tile_x, prow, pcol); at most 4 tiles
  template <class F>
  __host__ __device__ void for_each_tile(int h, int w, F&& f) const {
    // padded coordinates u = h + 1, v = w + 1; tile t covers [t*T, t*T + P - 1] with T = tile extent * stride
    const int T = TH * S, U = TW * S;
    const int u = h + 1, v = w + 1;
    const int ty_a = (int)(((uint32_t)u * magic_t) >> 16), tx_a = v >> (log2TW + (S == 2 ? 1 : 0));
    const int pr_a = u - ty_a * T, pc_a = v - tx_a * U;
    for (int dy = 0; dy < 2; ++dy) {
      const int ty = ty_a - dy, pr = pr_a + dy * T;
      if (ty < 0 || ty >= tiles_h || pr > PH - 1) continue;
      for (int dx = 0; dx < 2; ++dx) {
        const int tx = tx_a - dx, pc = pc_a + dx * U;
        if (tx < 0 || tx >= tiles_w || pc > PW - 1) continue;
        f(ty, tx, pr, pc);
      }
    }
  }
  // the tile in which pixel (h, w) is an INTERIOR position (stride-1 consumers: residual reads)
  __host__ __device__ void interior(int h, int w, int* ty, int* tx, int* prow, int* pcol) const {
    const int t = (int)(((uint32_t)h * magic_t) >> 16);
    *ty = t; *prow = h - t * TH + 1; *tx = w >> log2TW; *pcol = (w & (TW - 1)) + 1;
  }
};

// host: the two lookup tables of a layout (upload them and point dst_tbl / res_tbl at the device copies)
struct PatchTables { std::vector<uint32_t> dst /* 4 per pixel */, res /* 1 per pixel */; };
inline PatchTables patch_build_tables(const PatchTiles& t) {
  PatchTables T;
  T.dst.assign((size_t)t.H * t.W * 4, 0xffffffffu);
  T.res.assign((size_t)t.H * t.W, 0xffffffffu);
  for (int h = 0; h < t.H; ++h)
    for (int w = 0; w < t.W; ++w) {
      int k = 0;
      const size_t p = (size_t)h * t.W + w;
      t.for_each_tile(h, w, [&](int ty, int tx, int pr, int pc) {
        const uint32_t e = ((uint32_t)(ty * t.tiles_w + tx) << 8) | (uint32_t)(pr * t.PW + pc);
        T.dst[p * 4 + k++] = e;
        if (t.S == 1 && pr >= 1 && pr <= t.TH && pc >= 1 && pc <= t.TW) T.res[p] = e;
      });
    }
  return T;
}

}  // namespace bnb
