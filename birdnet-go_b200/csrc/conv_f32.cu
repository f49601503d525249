// conv_f32.cu — fp32 CUDA-core kernels of the BirdNET v2.4 conv stack (ops 95-375 of the .tflite
// the reference executes through /root/reference/internal/inference/tflite/classifier.go:107).
//
// These are (a) the truth path (BNB_PRECISION_F32) and (b) the kernels for every layer that is
// not a dense pointwise GEMM: stem 4x8/s2 conv fused with the avg/max pool + concat + 1x1 mix,
// 3x3 depthwise (+SiLU), squeeze-excite gate, global mean.  The dense 1x1 layers have a tcgen05
// implementation in pw_tc.cu; launch_pw_conv here is their fp32 reference and the fallback for
// shapes the tensor-core kernel does not take.
#include "kernels.h"
#include "tc_common.cuh"

namespace bnb {

namespace {

// =================================================================================================
// stem conv 4x8 stride 2 (SAME: pad top 1 / left 3) + ReLU  ->  {max,avg} pool 1x2 -> concat -> 1x1
// One CTA = one stem output row (256 positions = 128 pooled positions); one thread = one pooled
// position = two adjacent stem positions x 24 channels.
// Input rows 2h-1 .. 2h+2 are staged in smem de-interleaved by (column mod 4) so that the
// stride-4-column accesses of adjacent threads hit consecutive float2 words (no bank conflicts).
// =================================================================================================
constexpr int kStemCo = 24, kStemKh = 4, kStemKw = 8;
constexpr int kStemThreads = 128;
constexpr int kStemCols = 4 * kStemThreads + 8;       // padded input columns covered: [-3, 4*128+4]
constexpr int kStemQ = kStemCols / 4;                 // 130 column quads

__global__ void __launch_bounds__(kStemThreads)
stem_mix_kernel(const StemMixDev p, const float* __restrict__ in, float* __restrict__ stem_out, float* __restrict__ out,
                uint8_t* __restrict__ out_img, const PatchTiles op) {
  __shared__ __align__(16) float s_in[kStemKh][4][kStemQ][2];          // [row][col%4][col/4][ci]
  __shared__ __align__(16) float s_w[kStemKh * kStemKw * 2 * kStemCo];  // [kh][kw][ci][co]
  __shared__ __align__(16) float s_wm[kStemCo * 2 * kStemCo];           // [48][co] (transposed at load)
  __shared__ float s_b[kStemCo], s_bm[kStemCo];
  const int tid = threadIdx.x, h = blockIdx.x, b = blockIdx.y;
  for (int i = tid; i < kStemKh * kStemKw * 2 * kStemCo; i += kStemThreads) s_w[i] = __ldg(p.w_stem + i);
  for (int i = tid; i < kStemCo * 2 * kStemCo; i += kStemThreads) {     // [co][48] -> [c][co]
    const int co = i / (2 * kStemCo), c = i - co * 2 * kStemCo;
    s_wm[c * kStemCo + co] = __ldg(p.w_mix + i);
  }
  if (tid < kStemCo) { s_b[tid] = __ldg(p.b_stem + tid); s_bm[tid] = __ldg(p.b_mix + tid); }
  pdl_trigger();
  pdl_wait();                // weights above are constants; everything below reads the frontend's output / writes activations
  // stage input rows; padded column index pc = col + pad_l  (pc in [0, kStemCols))
  const float* inb = in + (size_t)b * p.in_h * p.in_w * 2;
  for (int i = tid; i < kStemKh * kStemCols; i += kStemThreads) {
    const int r = i / kStemCols, pc = i - r * kStemCols;
    const int row = 2 * h - p.pad_t + r, col = pc - p.pad_l;
    float2 v = make_float2(0.f, 0.f);
    if (row >= 0 && row < p.in_h && col >= 0 && col < p.in_w) v = __ldcg(reinterpret_cast<const float2*>(inb + ((size_t)row * p.in_w + col) * 2));   // previous kernel's output: coherent load (PDL)
    *reinterpret_cast<float2*>(&s_in[r][pc & 3][pc >> 2][0]) = v;
  }
  __syncthreads();

  // accumulators as channel PAIRS: every FMA of this loop is a packed fma.rn.f32x2 (sm_100 FFMA2: x broadcast to both
  // halves, two adjacent output channels' weights from one 64-bit register pair), halving the FMA-pipe instruction count
  float2 acc0[kStemCo / 2], acc1[kStemCo / 2];
#pragma unroll
  for (int c = 0; c < kStemCo / 2; ++c) { acc0[c] = make_float2(s_b[2 * c], s_b[2 * c + 1]); acc1[c] = acc0[c]; }
  // stem position w0 = 2*tid reads padded cols 4*tid + kw ; w1 = 2*tid+1 reads 4*tid + 2 + kw
#pragma unroll
  for (int kh = 0; kh < kStemKh; ++kh) {
#pragma unroll
    for (int kw = 0; kw < kStemKw; ++kw) {
      const int pc0 = kw, pc1 = kw + 2;   // + 4*tid
      const float2 x0 = *reinterpret_cast<const float2*>(&s_in[kh][pc0 & 3][tid + (pc0 >> 2)][0]);
      const float2 x1 = *reinterpret_cast<const float2*>(&s_in[kh][pc1 & 3][tid + (pc1 >> 2)][0]);
      const float2 x0a = make_float2(x0.x, x0.x), x0b = make_float2(x0.y, x0.y);
      const float2 x1a = make_float2(x1.x, x1.x), x1b = make_float2(x1.y, x1.y);
      const float* w = s_w + ((kh * kStemKw + kw) * 2) * kStemCo;
#pragma unroll
      for (int c4 = 0; c4 < kStemCo / 4; ++c4) {
        const float4 wa = *reinterpret_cast<const float4*>(w + 4 * c4);
        const float4 wb = *reinterpret_cast<const float4*>(w + kStemCo + 4 * c4);
        const float2 wa0 = make_float2(wa.x, wa.y), wa1 = make_float2(wa.z, wa.w);
        const float2 wb0 = make_float2(wb.x, wb.y), wb1 = make_float2(wb.z, wb.w);
        acc0[2 * c4] = __ffma2_rn(x0a, wa0, acc0[2 * c4]); acc0[2 * c4 + 1] = __ffma2_rn(x0a, wa1, acc0[2 * c4 + 1]);
        acc1[2 * c4] = __ffma2_rn(x1a, wa0, acc1[2 * c4]); acc1[2 * c4 + 1] = __ffma2_rn(x1a, wa1, acc1[2 * c4 + 1]);
        acc0[2 * c4] = __ffma2_rn(x0b, wb0, acc0[2 * c4]); acc0[2 * c4 + 1] = __ffma2_rn(x0b, wb1, acc0[2 * c4 + 1]);
        acc1[2 * c4] = __ffma2_rn(x1b, wb0, acc1[2 * c4]); acc1[2 * c4 + 1] = __ffma2_rn(x1b, wb1, acc1[2 * c4 + 1]);
      }
    }
  }
  // ReLU, optional dump of the stem tensor, then pool
  float pool[2 * kStemCo];   // [max(24) | avg(24)]
#pragma unroll
  for (int c = 0; c < kStemCo / 2; ++c) {
    acc0[c].x = fmaxf(acc0[c].x, 0.f); acc0[c].y = fmaxf(acc0[c].y, 0.f);
    acc1[c].x = fmaxf(acc1[c].x, 0.f); acc1[c].y = fmaxf(acc1[c].y, 0.f);
    pool[2 * c] = fmaxf(acc0[c].x, acc1[c].x); pool[2 * c + 1] = fmaxf(acc0[c].y, acc1[c].y);
    pool[kStemCo + 2 * c] = (acc0[c].x + acc1[c].x) * 0.5f; pool[kStemCo + 2 * c + 1] = (acc0[c].y + acc1[c].y) * 0.5f;
  }
  if (stem_out != nullptr) {
    float* so = stem_out + (((size_t)b * p.out_h + h) * p.out_w + 2 * tid) * kStemCo;
#pragma unroll
    for (int c = 0; c < kStemCo / 2; ++c) {
      so[2 * c] = acc0[c].x; so[2 * c + 1] = acc0[c].y; so[kStemCo + 2 * c] = acc1[c].x; so[kStemCo + 2 * c + 1] = acc1[c].y;
    }
  }
  // 1x1 mix 48 -> 24 (weights staged transposed [c][co]: pooled value broadcast x two adjacent outputs per FFMA2)
  float2 mix[kStemCo / 2];
#pragma unroll
  for (int c = 0; c < kStemCo / 2; ++c) mix[c] = make_float2(s_bm[2 * c], s_bm[2 * c + 1]);
#pragma unroll
  for (int c = 0; c < 2 * kStemCo; ++c) {
    const float2 pb = make_float2(pool[c], pool[c]);
    const float* w = s_wm + c * kStemCo;
#pragma unroll
    for (int q = 0; q < kStemCo / 4; ++q) {
      const float4 wv = *reinterpret_cast<const float4*>(w + 4 * q);
      mix[2 * q] = __ffma2_rn(pb, make_float2(wv.x, wv.y), mix[2 * q]);
      mix[2 * q + 1] = __ffma2_rn(pb, make_float2(wv.z, wv.w), mix[2 * q + 1]);
    }
  }
  const size_t opos = (((size_t)b * p.out_h + h) * (p.out_w / 2) + tid) * kStemCo;
  if (out_img != nullptr) {
    // F16X3 path: the block input leaves as fp16 hi / lo planes (x ~= hi + lo), written straight into the PatchTiles image that
    // block 1's fused kernel (mbconv2.cu) bulk-copies: the pixel goes into every tile whose halo patch contains it
    uint32_t hw[kStemCo / 2], lw[kStemCo / 2];
#pragma unroll
    for (int c = 0; c < kStemCo / 2; ++c) tc::split2(mix[c].x, mix[c].y, hw[c], lw[c]);
    const uint4 ent = __ldg(op.dst_tbl + (size_t)h * op.W + tid);
    const uint32_t e4[4] = {ent.x, ent.y, ent.z, ent.w};
    // 24 channels = chunks 0, 1, 2 of the pixel's 64-byte row; chunk 3 holds the pad channels 24..31 (zero expand weights), written
    // as zeros, so that each plane is two aligned 256-bit stores (chunks c, c + 1 of a swizzled row are the halves of one 32-byte
    // sector, in either order) instead of three half-sector ones
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (e4[d] == 0xffffffffu) continue;
#pragma unroll
      for (int qp = 0; qp < (kStemCo / 8 + 1) / 2; ++qp) {
        const int q0 = 2 * qp, q1 = q0 + 1;
        int st, chunk;
        op.stage_of(q0, &st, &chunk);
        const size_t a0 = op.entry_piece(b, e4[d], st, chunk);
        uint8_t* dst = out_img + (a0 & ~(size_t)31);
        const bool swapped = (a0 & 16) != 0;
        const uint4 hA = make_uint4(hw[4 * q0], hw[4 * q0 + 1], hw[4 * q0 + 2], hw[4 * q0 + 3]);
        const uint4 lA = make_uint4(lw[4 * q0], lw[4 * q0 + 1], lw[4 * q0 + 2], lw[4 * q0 + 3]);
        const uint4 hB = q1 < kStemCo / 8 ? make_uint4(hw[(4 * q1) % (kStemCo / 2)], hw[(4 * q1 + 1) % (kStemCo / 2)], hw[(4 * q1 + 2) % (kStemCo / 2)], hw[(4 * q1 + 3) % (kStemCo / 2)]) : zero4;
        const uint4 lB = q1 < kStemCo / 8 ? make_uint4(lw[(4 * q1) % (kStemCo / 2)], lw[(4 * q1 + 1) % (kStemCo / 2)], lw[(4 * q1 + 2) % (kStemCo / 2)], lw[(4 * q1 + 3) % (kStemCo / 2)]) : zero4;
        tc::stg256(dst, swapped ? hB : hA, swapped ? hA : hB);
        tc::stg256(dst + op.st_plane[st], swapped ? lB : lA, swapped ? lA : lB);
      }
    }
    return;
  }
  float* o = out + opos;
#pragma unroll
  for (int q = 0; q < kStemCo / 4; ++q)
    *reinterpret_cast<float4*>(o + 4 * q) = make_float4(mix[2 * q].x, mix[2 * q].y, mix[2 * q + 1].x, mix[2 * q + 1].y);
}

// =================================================================================================
// Generic pointwise GEMM  C[M,N] = act( f(A)[M,K] * W[N,K]^T + bias ) (+ residual)
//   f = optional per-chunk SE gate on k, or per-channel affine+ReLU (post block), and an implicit
//   im2col row addressing for the final KxK VALID conv.
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread, operands staged k-major in smem.
// =================================================================================================
constexpr int kBM = 64, kBN = 64, kBK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SILU: return silu_f(v);
    case ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(256)
pw_conv_kernel(const PwArgs a) {
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Ws[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int tx = tid & 15, ty = tid >> 4;     // thread computes rows ty*4..+3, cols tx*4..+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: each thread loads one float4 (4 consecutive k) of one row of A and of W per k-tile
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lr, wn = n0 + lr;
  const float* arow = nullptr;
  const float* grow = nullptr;
  int a_seg = 0, a_seg_stride = 0;     // im2col: k = seg*a_seg + j  ->  base + seg*a_seg_stride + j
  if (am < a.M) {
    if (a.a_mode == A_PLAIN) arow = a.A + (size_t)am * a.K;
    else {
      const int bidx = am / a.out_w, wo = am - bidx * a.out_w;
      // input map [kh rows][in_w][cin]; output (0, wo): k = (kh, kw, ci) ; row kh contributes kw*cin contiguous floats
      a_seg = a.kw * a.cin; a_seg_stride = a.in_w * a.cin;
      arow = a.A + ((size_t)bidx * (a.K / a_seg) * a.in_w + wo) * a.cin;
    }
    if (a.gate) grow = a.gate + (size_t)(am / a.rows_per_chunk) * a.K;
  }
  const float* wrow = (wn < a.N) ? a.W + (size_t)wn * a.K : nullptr;

  for (int k0 = 0; k0 < a.K; k0 += kBK) {
    const int k = k0 + lk;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow && k < a.K) {
      if (a.a_mode == A_PLAIN) av = __ldg(reinterpret_cast<const float4*>(arow + k));
      else { const int seg = k / a_seg, j = k - seg * a_seg; av = __ldg(reinterpret_cast<const float4*>(arow + (size_t)seg * a_seg_stride + j)); }
      if (grow) { const float4 g = __ldg(reinterpret_cast<const float4*>(grow + k)); av.x *= g.x; av.y *= g.y; av.z *= g.z; av.w *= g.w; }
      if (a.a_mul) {
        const int c = k % a.a_ch;
        const float4 mu = __ldg(reinterpret_cast<const float4*>(a.a_mul + c)), ad = __ldg(reinterpret_cast<const float4*>(a.a_add + c));
        av.x = fmaxf(av.x * mu.x + ad.x, 0.f); av.y = fmaxf(av.y * mu.y + ad.y, 0.f);
        av.z = fmaxf(av.z * mu.z + ad.z, 0.f); av.w = fmaxf(av.w * mu.w + ad.w, 0.f);
      }
    }
    if (wrow && k < a.K) wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 af = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 wf = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float ar[4] = {af.x, af.y, af.z, af.w}, wr[4] = {wf.x, wf.y, wf.z, wf.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue
  const int n = n0 + tx * 4;
  if (n < a.N) {   // N is a multiple of 4 except the FC head (6522 = 4*1630 + 2): handled per element
    float bs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[j] = (a.bias && n + j < a.N) ? __ldg(a.bias + n + j) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= a.M) continue;
      float* c = a.C + (size_t)m * a.N + n;
      const float* r = a.residual ? a.residual + (size_t)m * a.N + n : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < a.N) {
          float v = apply_act(acc[i][j] + bs[j], a.act);
          if (r) v += __ldg(r + j);
          c[j] = v;
        }
      }
    }
  }
}

// =================================================================================================
// 3x3 depthwise conv + bias + SiLU, NHWC, zero padding 1 on each side (stride 1 SAME, or the
// reference's explicit PAD(1,1) + VALID stride 2).  One thread = 4 channels of one output pixel.
// =================================================================================================
// Grid (Ho, B, channel groups): one CTA = one output row of one chunk.  A thread owns 4 channels (its 9 taps +
// bias live in registers) and walks along a segment of the row with a sliding 3x3 register window, so each new
// output costs 3 (stride 1) or 6 (stride 2) 16-byte loads instead of 9, all issued before the previous output's
// math.  The SiLU outputs are also summed per channel: partial[b][ho][c] is the squeeze-excite row sum
// (deterministic, no atomics), reduced by se_gate_kernel.
constexpr int kDwMaxThreads = 256;

template <int STRIDE>
__global__ void __launch_bounds__(kDwMaxThreads, 3)
dw_conv_kernel(const DwArgs a) {
  __shared__ float4 s_w[9][kDwMaxThreads];      // taps of this CTA's channels; also reused for the row-sum reduction
  const int c4_per_cta = a.c4_per_cta, lanes = blockDim.x / c4_per_cta;    // row segments per CTA
  const int cl = threadIdx.x % c4_per_cta, pl = threadIdx.x / c4_per_cta;
  const int c4 = blockIdx.z * c4_per_cta + cl, b = blockIdx.y, ho = blockIdx.x;
  const int c4n = a.C / 4;
  const int seg = (a.Wo + lanes - 1) / lanes;
  const int w0 = pl * seg, w1 = min(a.Wo, w0 + seg);
  if (pl == 0) {
#pragma unroll
    for (int t = 0; t < 9; ++t) s_w[t][cl] = __ldg(reinterpret_cast<const float4*>(a.w + t * a.C) + c4);
  }
  const float4 bz = __ldg(reinterpret_cast<const float4*>(a.bias) + c4);
  const float4* rowp[3];
  bool rv[3];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = ho * STRIDE - 1 + kh;
    rv[kh] = hi >= 0 && hi < a.H;                                          // uniform across the CTA
    rowp[kh] = reinterpret_cast<const float4*>(a.in + ((size_t)b * a.H + (rv[kh] ? hi : 0)) * a.W * a.C) + c4;
  }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ld = [&](int kh, int wi) { return (rv[kh] && wi >= 0 && wi < a.W) ? __ldg(rowp[kh] + (size_t)wi * c4n) : zero; };
  float4* outp = reinterpret_cast<float4*>(a.out + ((size_t)b * a.Ho + ho) * a.Wo * a.C) + c4;
  float4 sum = zero;
  // sliding window x0|x1|x2 per input row + the NEXT output's new columns prefetched one iteration ahead
  float4 x0[3], x1[3], x2[3], n1[3], n2[3];
  if (w0 < w1) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      if (STRIDE == 1) { x1[kh] = ld(kh, w0 - 1); x2[kh] = ld(kh, w0); n2[kh] = ld(kh, w0 + 1); n1[kh] = zero; }
      else { x2[kh] = ld(kh, 2 * w0 - 1); n1[kh] = ld(kh, 2 * w0); n2[kh] = ld(kh, 2 * w0 + 1); x1[kh] = zero; }
    }
  }
  __syncthreads();
  for (int wo = w0; wo < w1; ++wo) {
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      if (STRIDE == 1) { x0[kh] = x1[kh]; x1[kh] = x2[kh]; x2[kh] = n2[kh]; if (wo + 1 < w1) n2[kh] = ld(kh, wo + 2); }
      else { x0[kh] = x2[kh]; x1[kh] = n1[kh]; x2[kh] = n2[kh]; if (wo + 1 < w1) { n1[kh] = ld(kh, 2 * wo + 2); n2[kh] = ld(kh, 2 * wo + 3); } }
    }
    float4 acc = bz;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const float4 wa = s_w[kh * 3 + 0][cl], wb = s_w[kh * 3 + 1][cl], wc = s_w[kh * 3 + 2][cl];
      acc.x = fmaf(x0[kh].x, wa.x, acc.x); acc.y = fmaf(x0[kh].y, wa.y, acc.y); acc.z = fmaf(x0[kh].z, wa.z, acc.z); acc.w = fmaf(x0[kh].w, wa.w, acc.w);
      acc.x = fmaf(x1[kh].x, wb.x, acc.x); acc.y = fmaf(x1[kh].y, wb.y, acc.y); acc.z = fmaf(x1[kh].z, wb.z, acc.z); acc.w = fmaf(x1[kh].w, wb.w, acc.w);
      acc.x = fmaf(x2[kh].x, wc.x, acc.x); acc.y = fmaf(x2[kh].y, wc.y, acc.y); acc.z = fmaf(x2[kh].z, wc.z, acc.z); acc.w = fmaf(x2[kh].w, wc.w, acc.w);
    }
    acc.x = silu_f(acc.x); acc.y = silu_f(acc.y); acc.z = silu_f(acc.z); acc.w = silu_f(acc.w);
    outp[(size_t)wo * c4n] = acc;
    sum.x += acc.x; sum.y += acc.y; sum.z += acc.z; sum.w += acc.w;
  }
  if (a.partial != nullptr) {
    __syncthreads();                      // everyone is done with the taps: reuse row 0 of s_w for the reduction
    float4* s_red = &s_w[0][0];
    s_red[threadIdx.x] = sum;
    __syncthreads();
    if (pl == 0) {
      for (int l = 1; l < lanes; ++l) { const float4 o = s_red[l * c4_per_cta + cl]; sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w; }
      reinterpret_cast<float4*>(a.partial + ((size_t)b * a.Ho + ho) * a.C)[c4] = sum;
    }
  }
}

// =================================================================================================
// Squeeze-excite gate: gate[b][c] = sigmoid(W2 * silu(W1 * mean_hw(x[b]) + b1) + b2).
// One CTA = kSeChunks chunks so every weight row fetched from L2 is used for several chunks; 32 warps split the
// hidden units (float4 weight loads, warp-shuffle reductions); the channel means come from the per-row sums the
// depthwise kernel left behind.
// =================================================================================================
constexpr int kSeThreads = 1024;
constexpr int kSeMaxC = 1536, kSeMaxS = 64, kSeChunks = 2;

__global__ void __launch_bounds__(kSeThreads)
se_gate_kernel(const SeArgs a) {
  __shared__ __align__(16) float s_mean[kSeChunks][kSeMaxC];
  __shared__ float s_hidden[kSeChunks][kSeMaxS];
  const int b0 = blockIdx.x * kSeChunks, tid = threadIdx.x;
  const int nb = min(kSeChunks, a.B - b0);
  const float inv = 1.0f / (float)a.HW;
  pdl_trigger();
  pdl_wait();
  for (int i = tid; i < kSeChunks * a.C; i += kSeThreads) {
    const int g = i / a.C, c = i - g * a.C;
    // four independent partial accumulators: the loads of a round are in flight together (the serial version paid one L2
    // round trip per part); the summation order is fixed by the code, hence independent of batch composition
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (g < nb) {
      const float* pp = a.partial + (size_t)(b0 + g) * a.parts * a.C + c;
      int p = 0;
      for (; p + 4 <= a.parts; p += 4) {
        s0 += __ldcg(pp + (size_t)p * a.C); s1 += __ldcg(pp + (size_t)(p + 1) * a.C);       // coherent loads (PDL)
        s2 += __ldcg(pp + (size_t)(p + 2) * a.C); s3 += __ldcg(pp + (size_t)(p + 3) * a.C);
      }
      for (; p < a.parts; ++p) s0 += __ldcg(pp + (size_t)p * a.C);
    }
    s_mean[g][c] = ((s0 + s1) + (s2 + s3)) * inv;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j = warp; j < a.Cse; j += kSeThreads / 32) {
    const float4* w = reinterpret_cast<const float4*>(a.w1 + (size_t)j * a.C);
    float acc[kSeChunks];
#pragma unroll
    for (int g = 0; g < kSeChunks; ++g) acc[g] = 0.f;
    for (int c4 = lane; c4 < a.C / 4; c4 += 32) {
      const float4 wv = __ldg(w + c4);
#pragma unroll
      for (int g = 0; g < kSeChunks; ++g) {
        const float4 m = *reinterpret_cast<const float4*>(&s_mean[g][4 * c4]);
        acc[g] = fmaf(m.x, wv.x, fmaf(m.y, wv.y, fmaf(m.z, wv.z, fmaf(m.w, wv.w, acc[g]))));
      }
    }
#pragma unroll
    for (int g = 0; g < kSeChunks; ++g) {
      const float v = warp_sum(acc[g]);
      if (lane == 0) s_hidden[g][j] = silu_f(v + __ldg(a.b1 + j));
    }
  }
  __syncthreads();
  // FC2: thread = (4 adjacent channels, half of the hidden units): float4 weight loads, up to 8 in flight; the second half's
  // partial dot products go through shared memory (the channel means are dead by now) and are added in a fixed order.  The
  // kernel is a chain of three dependent phases on 32..128 CTAs: its time is load latency, so the point is few, wide batches
  // (C = 1536 took 2 rounds x 8 batches of scalar loads before: profiles/r02 launch lists, 27 us -> see r02 notes).
  {
    const int c4n = a.C / 4;                                   // C % 4 == 0 (checked at launch)
    const int slice = tid / c4n, c4 = tid - slice * c4n;
    const int jh = (a.Cse + 1) / 2;
    const int j0 = slice == 0 ? 0 : jh, j1 = slice == 0 ? jh : a.Cse;
    float4 acc[kSeChunks];
#pragma unroll
    for (int g = 0; g < kSeChunks; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slice < 2) {
      const float4* wp = reinterpret_cast<const float4*>(a.w2t) + c4;
      int j = j0;
      for (; j + 8 <= j1; j += 8) {
        float4 wv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) wv[t] = __ldg(wp + (size_t)(j + t) * c4n);
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
          for (int g = 0; g < kSeChunks; ++g) {
            const float h = s_hidden[g][j + t];
            acc[g].x = fmaf(h, wv[t].x, acc[g].x); acc[g].y = fmaf(h, wv[t].y, acc[g].y);
            acc[g].z = fmaf(h, wv[t].z, acc[g].z); acc[g].w = fmaf(h, wv[t].w, acc[g].w);
          }
      }
      for (; j + 4 <= j1; j += 4) {
        float4 wv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) wv[t] = __ldg(wp + (size_t)(j + t) * c4n);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int g = 0; g < kSeChunks; ++g) {
            const float h = s_hidden[g][j + t];
            acc[g].x = fmaf(h, wv[t].x, acc[g].x); acc[g].y = fmaf(h, wv[t].y, acc[g].y);
            acc[g].z = fmaf(h, wv[t].z, acc[g].z); acc[g].w = fmaf(h, wv[t].w, acc[g].w);
          }
      }
      for (; j < j1; ++j) {
        const float4 wv = __ldg(wp + (size_t)j * c4n);
#pragma unroll
        for (int g = 0; g < kSeChunks; ++g) {
          const float h = s_hidden[g][j];
          acc[g].x = fmaf(h, wv.x, acc[g].x); acc[g].y = fmaf(h, wv.y, acc[g].y);
          acc[g].z = fmaf(h, wv.z, acc[g].z); acc[g].w = fmaf(h, wv.w, acc[g].w);
        }
      }
      if (slice == 1) {
#pragma unroll
        for (int g = 0; g < kSeChunks; ++g) *reinterpret_cast<float4*>(&s_mean[g][4 * c4]) = acc[g];
      }
    }
    __syncthreads();
    if (slice == 0) {
      const float4 bz = __ldg(reinterpret_cast<const float4*>(a.b2) + c4);
#pragma unroll
      for (int g = 0; g < kSeChunks; ++g) {
        if (g < nb) {
          const float4 o = *reinterpret_cast<const float4*>(&s_mean[g][4 * c4]);
          float4 r;
          r.x = sigmoid_f(bz.x + (acc[g].x + o.x)); r.y = sigmoid_f(bz.y + (acc[g].y + o.y));
          r.z = sigmoid_f(bz.z + (acc[g].z + o.z)); r.w = sigmoid_f(bz.w + (acc[g].w + o.w));
          *reinterpret_cast<float4*>(a.gate + (size_t)(b0 + g) * a.C + 4 * c4) = r;
        }
      }
    }
  }
}

// relu(x*mul+add) + im2col for the final KxK VALID conv: in [B][kh][in_w][cin] -> out [B*out_w][kh*kw*cin]
__global__ void __launch_bounds__(256)
post_prep_kernel(const float* __restrict__ in, const float* __restrict__ mul, const float* __restrict__ add, float* __restrict__ out,
                 int B, int kh, int kw, int in_w, int out_w, int cin) {
  const int c4n = cin / 4, K4 = kh * kw * c4n;
  const long long total = (long long)B * out_w * K4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(idx % K4);
    const long long row = idx / K4;
    const int wo = (int)(row % out_w), b = (int)(row / out_w);
    const int c4 = k4 % c4n, t = k4 / c4n, x = t % kw, y = t / kw;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * kh + y) * in_w + wo + x) * cin) + c4);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mul) + c4), a = __ldg(reinterpret_cast<const float4*>(add) + c4);
    float4 o;
    o.x = fmaxf(fmaf(v.x, m.x, a.x), 0.f); o.y = fmaxf(fmaf(v.y, m.y, a.y), 0.f);
    o.z = fmaxf(fmaf(v.z, m.z, a.z), 0.f); o.w = fmaxf(fmaf(v.w, m.w, a.w), 0.f);
    reinterpret_cast<float4*>(out)[idx] = o;
  }
}

__global__ void __launch_bounds__(256)
row_mean_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int rows, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx - b * C;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += __ldg(in + ((size_t)b * rows + r) * C + c);
  out[idx] = s / (float)rows;
}

// ---- F16X3 path: the same two helpers on fp16 hi/lo planes -------------------------------------------------------------
// relu(x*mul+add) + im2col: in planes [B][kh][in_w][cin] -> out planes [B*out_w][kh*kw*cin], 8 channels (16 B per plane) per thread
__global__ void __launch_bounds__(256)
post_prep2_kernel(const __half* __restrict__ ih, const __half* __restrict__ il, const float* __restrict__ mul, const float* __restrict__ add,
                  uint8_t* __restrict__ o_img, const RowTiles ot, int B, int kh, int kw, int in_w, int out_w, int cin) {
  const int c8n = cin / 8, K8 = kh * kw * c8n;
  const long long total = (long long)B * out_w * K8;
  pdl_trigger();
  pdl_wait();
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(idx % K8);
    const long long row = idx / K8;
    const int wo = (int)(row % out_w), b = (int)(row / out_w);
    const int c8 = k8 % c8n, t = k8 / c8n, x = t % kw, y = t / kw;
    const size_t src = ((((size_t)b * kh + y) * in_w + wo + x) * cin) + 8 * c8;
    const uint4 h = __ldcg(reinterpret_cast<const uint4*>(ih + src)), l = __ldcg(reinterpret_cast<const uint4*>(il + src));   // coherent (PDL)
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mul + 8 * c8)), m1 = __ldg(reinterpret_cast<const float4*>(mul + 8 * c8 + 4));
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(add + 8 * c8)), a1 = __ldg(reinterpret_cast<const float4*>(add + 8 * c8 + 4));
    const float2 v0 = tc::join2(h.x, l.x), v1 = tc::join2(h.y, l.y), v2 = tc::join2(h.z, l.z), v3 = tc::join2(h.w, l.w);
    uint4 ho, lo;
    tc::split2(fmaxf(fmaf(v0.x, m0.x, a0.x), 0.f), fmaxf(fmaf(v0.y, m0.y, a0.y), 0.f), ho.x, lo.x);
    tc::split2(fmaxf(fmaf(v1.x, m0.z, a0.z), 0.f), fmaxf(fmaf(v1.y, m0.w, a0.w), 0.f), ho.y, lo.y);
    tc::split2(fmaxf(fmaf(v2.x, m1.x, a1.x), 0.f), fmaxf(fmaf(v2.y, m1.y, a1.y), 0.f), ho.z, lo.z);
    tc::split2(fmaxf(fmaf(v3.x, m1.z, a1.z), 0.f), fmaxf(fmaf(v3.y, m1.w, a1.w), 0.f), ho.w, lo.w);
    uint8_t* dst = o_img + ot.piece(row, k8);              // RowTiles image of the im2col matrix [B*out_w][kh*kw*cin]
    *reinterpret_cast<uint4*>(dst) = ho;
    *reinterpret_cast<uint4*>(dst + 16384) = lo;
  }
}

// mean over `rows` consecutive rows -> fp32 [B][C] (the embedding the API returns) and its hi/lo planes (the FC head's A operand)
__global__ void __launch_bounds__(256)
row_mean2_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ o_img, const RowTiles ot, int B, int rows, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one thread = two adjacent channels
  pdl_trigger();
  pdl_wait();
  if (idx >= B * (C / 2)) return;
  const int b = idx / (C / 2), c = 2 * (idx - b * (C / 2));
  float s0 = 0.f, s1 = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(in + ((size_t)b * rows + r) * C + c));   // coherent (PDL)
    s0 += v.x; s1 += v.y;
  }
  s0 /= (float)rows; s1 /= (float)rows;
  *reinterpret_cast<float2*>(out + (size_t)b * C + c) = make_float2(s0, s1);
  uint32_t h, l;
  tc::split2(s0, s1, h, l);
  uint8_t* dst = o_img + ot.piece(b, c >> 3) + (size_t)(c & 7) * 2u;     // RowTiles image of the embedding matrix [B][C]
  *reinterpret_cast<uint32_t*>(dst) = h;
  *reinterpret_cast<uint32_t*>(dst + 16384) = l;
}

}  // namespace

void launch_post_prep2(const __half* ih, const __half* il, const float* mul, const float* add, uint8_t* o_img, int B, int kh, int kw,
                       int in_w, int out_w, int cin, cudaStream_t s, LaunchCounter& lc) {
  if (cin % 8) throw std::runtime_error("post_prep2: channel count must be a multiple of 8");
  const long long total = (long long)B * out_w * kh * kw * (cin / 8);
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  launch_k(post_prep2_kernel, dim3((unsigned)blocks), dim3(256), 0, s, ih, il, mul, add, o_img, RowTiles::make(kh * kw * cin), B, kh, kw, in_w, out_w, cin);
  lc.n++;
}

void launch_row_mean2(const float* in, float* out, uint8_t* o_img, int B, int rows, int C, cudaStream_t s, LaunchCounter& lc) {
  if (C % 2) throw std::runtime_error("row_mean2: channel count must be even");
  launch_k(row_mean2_kernel, dim3(ceil_div(B * (C / 2), 256)), dim3(256), 0, s, in, out, o_img, RowTiles::make(C), B, rows, C);
  lc.n++;
}

void launch_stem_mix(const StemMixDev& p, const float* in, float* stem_out_or_null, float* out, int B,
                     cudaStream_t s, LaunchCounter& lc, uint8_t* out_img, const PatchTiles* out_patch) {
  dim3 grid(p.out_h, B);
  if (out_img && (!out_patch || !out_patch->dst_tbl || out_patch->C != kStemCo || out_patch->H != p.out_h || out_patch->W != p.out_w / 2)) throw std::runtime_error("stem_mix: patch layout does not match the pooled stem output");
  if (out_img && (out_patch->n_stages != 1 || out_patch->st_rb[0] < 64)) throw std::runtime_error("stem_mix: the 24-channel patch image must be one stage with at least four 16-byte chunks per row");
  launch_k(stem_mix_kernel, grid, dim3(kStemThreads), 0, s, p, in, stem_out_or_null, out, out_img, out_patch ? *out_patch : PatchTiles());
  lc.n++;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("BNB_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

void launch_pw_conv(const PwArgs& a, cudaStream_t s, LaunchCounter& lc) {
  dim3 grid(ceil_div(a.M, kBM), ceil_div(a.N, kBN));
  pw_conv_kernel<<<grid, 256, 0, s>>>(a);
  BNB_LAUNCH_CHECK(lc);
}

int dw_parts(int B, int Ho, int Wo, int C, int W, int stride) { (void)B; (void)Wo; (void)C; (void)W; (void)stride; return Ho; }   // one partial sum per output row

void launch_dw_conv(const DwArgs& a0, cudaStream_t s, LaunchCounter& lc) {
  DwArgs a = a0;
  const int c4n = a.C / 4;
  const int groups = (c4n + kDwMaxThreads - 1) / kDwMaxThreads;
  if (c4n % groups) throw std::runtime_error("dw_conv: channel count not divisible into CTA groups");
  a.c4_per_cta = c4n / groups;
  int lanes = kDwMaxThreads / a.c4_per_cta > 0 ? kDwMaxThreads / a.c4_per_cta : 1;
  if (lanes > a.Wo) lanes = a.Wo;
  dim3 grid(a.Ho, a.B, groups);
  if (a.stride == 1) dw_conv_kernel<1><<<grid, lanes * a.c4_per_cta, 0, s>>>(a);
  else if (a.stride == 2) dw_conv_kernel<2><<<grid, lanes * a.c4_per_cta, 0, s>>>(a);
  else throw std::runtime_error("dw_conv: unsupported stride");
  BNB_LAUNCH_CHECK(lc);
}

void launch_se_gate(const SeArgs& a, cudaStream_t s, LaunchCounter& lc) {
  if (a.C > kSeMaxC || a.Cse > kSeMaxS || a.C % 4 || a.C / 2 > kSeThreads) throw std::runtime_error("se_gate: channel count exceeds kernel limits");
  launch_k(se_gate_kernel, dim3((a.B + kSeChunks - 1) / kSeChunks, 1), dim3(kSeThreads), 0, s, a);
  lc.n++;
}

void launch_post_prep(const float* in, const float* mul, const float* add, float* out, int B, int kh, int kw, int in_w, int out_w,
                      int cin, cudaStream_t s, LaunchCounter& lc) {
  const long long total = (long long)B * out_w * kh * kw * (cin / 4);
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  post_prep_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, mul, add, out, B, kh, kw, in_w, out_w, cin);
  BNB_LAUNCH_CHECK(lc);
}

void launch_row_mean(const float* in, float* out, int B, int rows, int C, cudaStream_t s, LaunchCounter& lc) {
  row_mean_kernel<<<ceil_div(B * C, 256), 256, 0, s>>>(in, out, B, rows, C);
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
