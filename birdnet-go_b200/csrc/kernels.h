// kernels.h — host-callable launchers of the sm_100a kernels (definitions in the .cu files).
//
// All activations are NHWC float32 with H = mel axis, W = time axis, exactly the layout of the
// tensors in the reference's .tflite graph, so every intermediate can be compared 1:1 with the
// oracle (tests/ use bnb_debug_read_tensor for that).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "layouts.h"

namespace bnb {

// ---------------------------------------------------------------- frontend (frontend.cu)
constexpr int kMinMaxParts = 8;       // partial min/max blocks per chunk
constexpr int kFeFramesPerCta = 32;   // frames of each spectrogram handled by one CTA

struct FrontendDev {
  // geometry (validated against these fixed values at load time)
  int n_samples;     // 144000
  int n_frames;      // 511
  int n_mel;         // 96
  float eps, center, gain;
  float pow_exp[2], bn_scale[2], bn_shift[2];
  int nk1[2];        // number of 32-/16-bin groups of the real spectrum that the mel matrix touches
  int mel_stride[2]; // padded non-zeros per mel band
  // device tables
  const float* win[2];       // [2048], [1024]
  const float2* tw[2];       // [32][32], [16][32]: exp(-2*pi*i*l*k2/N), indexed [k2][l]
  const float2* post[2];     // [N/2]: 0.5*(cos, sin)(pi*k/(N/2))  (N/2 = complex FFT length)
  const float* win_dft[2];   // [N/2]: Re DFT of the window constants (fp64 at load) for the DC compensation
  const int* mel_start[2];   // [96]
  const int* mel_cnt[2];     // [96]
  const float* mel_w[2];     // [96][mel_stride]
};

void launch_minmax(const void* pcm, int fmt, int B, int n_samples, float* partial, cudaStream_t s, LaunchCounter& lc);
void launch_frontend(const FrontendDev& fe, const void* pcm, int fmt, int B, const float* partial, float* out,
                     cudaStream_t s, LaunchCounter& lc);
size_t frontend_smem_bytes();

// ---------------------------------------------------------------- fp32 CUDA-core conv stack (conv_f32.cu)
struct StemMixDev {
  const float* w_stem;  // [kh=4][kw=8][ci=2][co=24]  (re-laid from OHWI at load)
  const float* b_stem;  // [24]
  const float* w_mix;   // [co=24][c=48]  concat order already normalised to (max, avg)
  const float* b_mix;   // [24]
  int in_h, in_w, out_h, out_w, pad_t, pad_l;   // 96, 511, 48, 256, 1, 3
};
// out_img != null: write the result as fp16 hi / lo planes into the PatchTiles image block 1 reads (F16X3 path) instead of fp32 `out`
void launch_stem_mix(const StemMixDev& p, const float* in, float* stem_out_or_null, float* out, int B,
                     cudaStream_t s, LaunchCounter& lc, uint8_t* out_img = nullptr, const PatchTiles* out_patch = nullptr);

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_SIGMOID = 3 };

// A-operand addressing modes of the generic pointwise GEMM
enum AMode : int {
  A_PLAIN = 0,      // A[m][k] = in[m*K + k]
  A_CONV3X3_ROW = 1 // post conv: kh x kw VALID over a [H=kh, W, C] map -> out [1, W-kw+1]; see conv_f32.cu
};

struct PwArgs {
  const float* A; const float* W; const float* bias; float* C;
  const float* residual;   // [M][N] or null
  const float* gate;       // [M / rows_per_chunk][K] or null: A[m][k] *= gate[m / rows_per_chunk][k]
  const float* a_mul; const float* a_add;  // optional per-k affine + ReLU on A (post block): relu(a*mul[c]+add[c]), c = k % a_ch
  int a_ch;
  int M, N, K;
  int rows_per_chunk;
  int act;
  int a_mode;
  int in_w, out_w, cin, kw;  // A_CONV3X3_ROW geometry
};
void launch_pw_conv(const PwArgs& a, cudaStream_t s, LaunchCounter& lc);

struct DwArgs {
  const float* in; const float* w; const float* bias; float* out;  // w: [3][3][C]
  int B, H, W, C, stride, Ho, Wo;
  float* partial;   // [B][parts][C] per-part channel sums of the output (for squeeze-excite) or null
  int parts;        // pixel partitions per chunk (0 = let the launcher choose via dw_parts)
  int c4_per_cta;   // set by the launcher
};
int dw_parts(int B, int Ho, int Wo, int C, int W, int stride);
void launch_dw_conv(const DwArgs& a, cudaStream_t s, LaunchCounter& lc);

struct SeArgs {
  const float* partial;  // [B][parts][C] channel sums from the depthwise kernel
  const float* w1; const float* b1;   // [Cse][C]
  const float* w2t; const float* b2;  // [Cse][C]  (second FC transposed at load)
  float* gate;      // [B][C]
  int B, HW, C, Cse, parts;
};
void launch_se_gate(const SeArgs& a, cudaStream_t s, LaunchCounter& lc);

// relu(x*mul+add) + im2col of the final KxK VALID conv input (tensor-core path: the GEMM then sees a plain [M][K] matrix)
void launch_post_prep(const float* in, const float* mul, const float* add, float* out, int B, int kh, int kw, int in_w, int out_w,
                      int cin, cudaStream_t s, LaunchCounter& lc);

// mean over `rows` consecutive rows: in [B][rows][C] -> out [B][C]
void launch_row_mean(const float* in, float* out, int B, int rows, int C, cudaStream_t s, LaunchCounter& lc);
// the same two helpers on fp16 hi/lo planes (F16X3 path)
// (outputs are RowTiles images: the A operands of the post-conv GEMM and of the FC head)
void launch_post_prep2(const __half* ih, const __half* il, const float* mul, const float* add, uint8_t* o_img, int B, int kh, int kw,
                       int in_w, int out_w, int cin, cudaStream_t s, LaunchCounter& lc);
void launch_row_mean2(const float* in, float* out, uint8_t* o_img, int B, int rows, int C, cudaStream_t s, LaunchCounter& lc);

// ---------------------------------------------------------------- post-processing (post.cu)
// conf = sigmoid(sensitivity * logit) (float64 exp like analyze.go:113-115), top-k descending, ties -> lower index
void launch_sigmoid_topk(const float* logits, int B, int n, float sensitivity, int k, int32_t* idx, float* conf,
                         cudaStream_t s, LaunchCounter& lc);
// N1: detections >= threshold of the sorted per-chunk top-k, compacted over the batch (chunk order, then confidence order)
void launch_compact_detections(const int32_t* idx, const float* conf, int B, int k, float threshold, int max_det, int32_t* det_chunk,
                               int32_t* det_idx, float* det_conf, int32_t* counts, int32_t* n_det, cudaStream_t s, LaunchCounter& lc);
// N2: dense + sigmoid head on embeddings (bat pipeline), and the ultrasonic frame-power CV filter (ultrasonic.cu)
void launch_dense_head(const float* emb, const float* w, const float* bias, int B, int n_in, int n_out, float* out, cudaStream_t s);
int ultrasonic_frames(int n_samples, int sample_rate, int fft, int hop, int split_hz);
size_t ultrasonic_workspace_bytes(int B, int n_frames, int fft);
void launch_ultrasonic_cv(const void* d_pcm, int fmt, int B, int n_samples, int sample_rate, int fft, int hop, int split_hz,
                          void* workspace, double* d_cv, cudaStream_t s);

}  // namespace bnb
