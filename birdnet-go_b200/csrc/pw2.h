// pw2.h — dense GEMM layers (1x1 project convs, final 3x3 conv after im2col, FC head) on fp16 hi/lo planes (pw2.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "kernels.h"
#include "pw_tc.h"

namespace bnb {

struct Pw2Launch {
  const __half* ah; const __half* al;   // A planes [M][a_pitch] fp16 (x ~= hi + lo), K-major
  int a_pitch;                          // elements per row in memory (multiple of 8)
  const uint8_t* Wimg;                  // pw_tc_prepare image of W[N][K]
  const float* bias;                    // [N] padded with zeros to n_pad + 64
  const float* gate;                    // [M / rows_per_chunk][K] SE gate applied to A on the fly, or null
  const __half* rh; const __half* rl;   // residual planes [M][r_pitch] or null
  int r_pitch;
  __half* oh; __half* ol; int o_pitch;  // output planes [M][o_pitch] (o_pitch multiple of 8, >= N), or null
  float* out32;                         // fp32 output [M][N] (exactly one of oh / out32 is set)
  int M, N, K, rows_per_chunk, act;
};
// tiling decision (host logic, CPU-testable): N-tile width, pipeline stages, shared-memory bytes, resident-weights flag
void pw2_tiling(const PwTcLayer& L, int M, bool conv, int* bn, int* stages, size_t* smem_bytes, int* b_res);
void launch_pw2(const PwTcLayer& L, const Pw2Launch& p, cudaStream_t s, LaunchCounter& lc);
void pw2_set_attributes();
void mb2_set_attributes();

}  // namespace bnb
