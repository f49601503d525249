// mbconv2.h — front half of an MBConv block on fp16 hi/lo activation planes (see mbconv2.cu).
#pragma once
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "kernels.h"
#include "layouts.h"

namespace bnb {

constexpr size_t kMb2SmemLimit = 227 * 1024;
constexpr int kMb2MaxStages = 3;
constexpr int kMb2BufCols = 128;      // TMEM columns per tile patch; an accumulator buffer holds a PAIR of tiles (2 buffers x 256 columns)

// Host-side plan of one layer: tile geometry, K stages, shared-memory layout.  Pure host logic (CPU-testable).
struct Mb2Plan {
  bool ok = false;
  int S = 1, TW = 0, TH = 0, PW = 0, PH = 0, n_mma = 0, tiles_h = 0, tiles_w = 0;
  int Cin = 0, C = 0, n_units = 0, k_stages = 0;
  int st_rb[kMb2MaxStages] = {0, 0, 0};        // row bytes of the stage's K-major tiles: 128 / 64 / 32
  int st_k0[kMb2MaxStages] = {0, 0, 0};        // first input channel of the stage
  int st_kw[kMb2MaxStages] = {0, 0, 0};        // channels the stage's TMA box covers (64 / 32 / 16)
  int st_ksteps[kMb2MaxStages] = {0, 0, 0};    // 16-wide MMA k-steps that hold real channels
  uint32_t st_aoff[kMb2MaxStages] = {0, 0, 0};   // offset of the stage's slab inside one unit of the weight image
  uint32_t st_aplane[kMb2MaxStages] = {0, 0, 0}; // bytes of one weight plane (128 rows)
  uint32_t st_boff[kMb2MaxStages] = {0, 0, 0};   // offset of the stage inside a patch slot
  uint32_t st_bplane[kMb2MaxStages] = {0, 0, 0}; // bytes reserved for one patch plane
  uint32_t img_unit_bytes = 0;                   // weight image bytes per unit (all stages, hi | lo)
  uint32_t a_slot_bytes = 0, a_region_bytes = 0, b_slot_bytes = 0, b_tx_bytes = 0;
  // shared-memory slot of a PAIR of tiles (the kernel multiplies two patches per MMA): per stage [hi: tile 0 | tile 1][lo: ...]
  uint32_t b_pair_bytes = 0, st_poff[kMb2MaxStages] = {0, 0, 0}, st_pplane[kMb2MaxStages] = {0, 0, 0};
  int a_slots = 0, b_slots = 0, a_resident = 0;
  size_t smem_bytes = 0;
};
// sw32_tail: allow the 32-byte-swizzle mode for K tails of <= 16 channels (else every stage is a 128-byte row stage)
Mb2Plan mb2_plan(int H, int W, int Ho, int Wo, int stride, int Cin, int C, bool sw32_tail = true);

// expand weights [C][Cin] fp32 -> per-(unit, stage) hi | lo swizzled K-major planes of 128 rows
void mb2_prepare_weights(const Mb2Plan& P, const float* w, std::vector<uint8_t>* image);

// the PatchTiles image this layer READS (its producer writes it): geometry + stage layout of the plan
PatchTiles mb2_patch_layout(const Mb2Plan& P, int H, int W);

struct Mb2Launch {
  const uint8_t* x_img;                 // block input as the PatchTiles image of THIS layer's plan (mb2_patch_layout)
  const uint8_t* Wimg;                  // mb2_prepare_weights image
  const float* bias_e;                  // expand bias, padded to n_units * 128
  const float* w_dw;                    // depthwise taps [9][C]
  const float* bias_dw;                 // [C]
  uint8_t* d_img;                       // depthwise output [B*Ho*Wo][C] as a RowTiles image (RowTiles::make(C))
  float* partial;                       // [B][tiles_h * tiles_w][C] SE sums per tile, or null
  int B, H, W, Ho, Wo;
};
void launch_mbconv2(const Mb2Plan& P, const Mb2Launch& L, cudaStream_t s, LaunchCounter& lc);

// device-global set-up shared by the tcgen05 kernels: per-device max dynamic shared memory attributes
void tc_prepare_device(int device);

}  // namespace bnb
