// pw_tc.h — tcgen05 pointwise-GEMM layer descriptor and launcher (see pw_tc.cu).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "kernels.h"

namespace bnb {

struct PwTcLayer {
  int N = 0, K = 0, n_pad = 0, k_pad = 0, k_stages = 0;
};

struct PwTcArgs {
  const float* A; const uint8_t* Wimg; const float* bias; float* C; const float* residual; const float* gate;
  int M, N, K, rows_per_chunk, act;
  int n_pad, k_pad, n_tiles, bn, stages, c_vec4, box_k;
  int b_res;          // 1: one-stage layer, the weight tile of the CTA's n-tile stays resident in smem
  int dbg;            // debug knobs (BNB_PWTC_DBG): 1 = no activation, 2 = no global store, 4 = no TMEM load, 8 = no staging
  long long* trace;   // debug: per-role clock64 timestamps of CTA 0 (BNB_PWTC_TRACE), else null
};

// Split W[N][K] (fp32, K-major = the OHWI / [O,I] layout of the .tflite) into fp16 hi/lo and lay both out as
// the 128B-swizzled K-major shared-memory image of every (n-tile, k-stage); returns the tiling.
PwTcLayer pw_tc_prepare(const float* w, int N, int K, std::vector<uint8_t>* image);
// tiling decision for a layer at a given M (exposed for the CPU tests)
void pw_tc_tiling(const PwTcLayer& L, int M, int* bn, int* stages, size_t* smem_bytes, int* b_res = nullptr);
void launch_pw_tc(const PwTcLayer& L, const PwArgs& p, const uint8_t* d_image, cudaStream_t s, LaunchCounter& lc);

}  // namespace bnb
