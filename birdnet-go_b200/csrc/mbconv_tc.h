// mbconv_tc.h — fused expand(1x1, tcgen05) + SiLU + depthwise 3x3 + SiLU + SE row sums (see mbconv_tc.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "kernels.h"

namespace bnb {

struct MbGeom {
  int th, tw, ph, pw, tiles_h, tiles_w;   // output tile, input patch (<= 128 positions), tiles per chunk
  int k_stages, box_c, a_slots, b_slots;
  size_t smem_bytes;
};
constexpr size_t kMbSmemLimit = 227 * 1024;
// tile geometry + shared-memory plan for one block (host logic, also used by the CPU tests)
// C = expanded channels, B = chunks in the launch (both feed the cost model; 0 = unknown), max_tiles = cap on tiles per
// chunk (SE partial-sum slots; 0 = none)
MbGeom mbconv_geometry(int H, int W, int Ho, int Wo, int stride, int Cin, int C = 0, int B = 0, int max_tiles = 0);

struct MbArgs {
  const uint8_t* Wimg; const float* bias_e; const float* w_dw; const float* bias_dw; float* D; float* partial;
  int B, H, W, Cin, C, Ho, Wo, stride;
  int th, tw, ph, pw, tiles_h, tiles_w;
  int n_pad, k_pad, k_stages, box_c, a_slots, b_slots;
  long long* trace;   // debug timeline of CTA 0 (BNB_MB_TRACE), else null
};

struct MbLaunch {
  const float* x;          // block input [B][H][W][Cin] fp32
  const uint8_t* Wimg;     // expand weights: the pre-split, swizzled image prepared by pw_tc_prepare
  const float* bias_e;     // expand bias (padded)
  const float* w_dw;       // depthwise taps [9][C]
  const float* bias_dw;    // [C]
  float* D;                // depthwise output [B][Ho][Wo][C]
  float* partial;          // [B][tiles_h*tiles_w][C] SE sums per tile, or null
  int B, H, W, Cin, C, Ho, Wo, stride;
  int max_tiles;           // cap on tiles per chunk (size of the SE partial-sum buffer), 0 = none
  int B_nominal;           // launch size the tile search is run for (fixed per classifier: keeps results batch-invariant)
};
void launch_mbconv_tc(const MbLaunch& L, cudaStream_t s, LaunchCounter& lc);

}  // namespace bnb
