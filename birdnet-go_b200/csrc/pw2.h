// pw2.h — dense GEMM layers (1x1 project convs, final 3x3 conv after im2col, FC head) on fp16 hi/lo planes (pw2.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "kernels.h"
#include "layouts.h"
#include "pw_tc.h"

namespace bnb {

struct Pw2Launch {
  const uint8_t* a_img;                 // A [M][K] as a RowTiles image (RowTiles::make(K)): fp16 hi | lo planes per (128-row tile, 64-channel stage)
  const uint8_t* Wimg;                  // pw_tc_prepare image of W[N][K]
  const float* bias;                    // [N] padded with zeros to n_pad + 64
  const float* gate;                    // [M / rows_per_chunk][K] SE gate applied to A on the fly, or null
  const uint8_t* r_img = nullptr;       // residual: the block INPUT, a PatchTiles image (r_patch: stride-1 geometry, pixel m interior), or null
  PatchTiles r_patch;
  // exactly one output: fp32 [M][N]; plain hi/lo planes [M][o_pitch]; or the PatchTiles image the NEXT block reads
  float* out32 = nullptr;
  __half* oh = nullptr; __half* ol = nullptr; int o_pitch = 0;
  uint8_t* o_img = nullptr; PatchTiles o_patch;
  int M, N, K, rows_per_chunk, act;
};
// tiling decision (host logic, CPU-testable): N-tile width, pipeline stages, shared-memory bytes, resident-weights flag
void pw2_tiling(const PwTcLayer& L, int M, bool conv, int* bn, int* stages, size_t* smem_bytes, int* b_res);
void launch_pw2(const PwTcLayer& L, const Pw2Launch& p, cudaStream_t s, LaunchCounter& lc);
void pw2_set_attributes();
void mb2_set_attributes();

}  // namespace bnb
