// frontend.cu — the in-graph mel frontend of BirdNET v2.4 as sm_100a kernels.
//
// What the reference computes (ops 0-94 of the .tflite it runs through
// /root/reference/internal/inference/tflite/classifier.go:107 `Invoke`; NOT
// /root/reference/internal/spectrogram, which only renders PNGs — SURVEY.md §0.1):
//   x <- ((x - min x) / (max(x - min x) + 1e-6) - 0.5) * 2                     per 3 s chunk
//   spec s in {0,1}: frames (2048/hop 278 | 1024/hop 280) x 511, * periodic Hann,
//       RFFT -> keep the REAL part only, mel[96 x bins] (sparse triangles), y*y, y^p, flip mel axis
//   concat on channel, per-channel affine (folded BatchNorm)  ->  [B, 96, 511, 2] NHWC
//
// Kernel design (HBM/L2-bound, fp32 on CUDA cores — the frontend is the precision-critical
// part):
//   * minmax_partial_kernel: 8 CTAs per chunk, vectorised 16 B loads, warp-shuffle reductions.
//   * frontend_kernel: one CTA = 32 consecutive frames of BOTH spectrograms of one chunk; the
//     ~10.7 k-sample PCM span is staged once in shared memory (each sample is reused by ~7
//     frames x 2 spectrograms).  One warp = one frame: a 2048-point real FFT is a 1024-point
//     complex FFT done as 32 x 32 "four-step": every lane runs a 32-point radix-2 FFT entirely in
//     registers (compile-time twiddles), a twiddle multiply, a 32x32 transpose through a padded
//     per-warp scratch, a second in-register 32-point FFT; the real-input split needs the
//     mirrored bin, which lives in lane (32-l)&31 and is fetched with one warp shuffle.  The
//     1024-point frames of the second spectrogram run two per warp (16 x 32 decomposition) so all
//     lanes stay busy.  Only the bins the mel matrix touches are post-processed (127 of 1025 and
//     309 of 513).  Results are staged in shared memory and written as 256 B rows.
#include <math.h>

#include "kernels.h"

namespace bnb {

namespace {

constexpr int kSpanMax = 10672;                 // floats of PCM staged per CTA (>= 31*278+2048, multiple of 4)
constexpr int kWarps = 16;
constexpr int kScratchPerWarp = 32 * 33;        // padded 32x32 transpose tile
constexpr int kPost1 = 128;                     // post-processing table entries kept in smem (spec 0)
constexpr int kPost2 = 512;
constexpr int kSmemFloats = kSpanMax + 2048 + 1024 + 2048 + 1024 + 2 * kPost1 + 2 * kPost2 + kPost1 + kPost2 + kWarps * kScratchPerWarp + 96 * 32 * 2;

__host__ __device__ constexpr int brev(int k, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((k >> i) & 1) << (bits - 1 - i);
  return r;
}
// cos/sin(2*pi*k/32), k = 0..15
__host__ __device__ constexpr float w32c(int k) {
  constexpr float t[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                           0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                           0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                           -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
  return t[k];
}
__host__ __device__ constexpr float w32s(int k) {
  constexpr float t[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                           0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                           1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                           0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
  return t[k];
}

// In-place radix-2 decimation-in-frequency FFT of N points held in registers a[OFF..OFF+N).
// Result: a[OFF + brev(k)] = X[k].  Every index and twiddle is a compile-time constant.
template <int N, int OFF, int LEN>
struct DifStage {
  static __device__ __forceinline__ void run(float2* a) {
    constexpr int H = LEN / 2;
#pragma unroll
    for (int i = 0; i < N; i += LEN) {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        // complex add / subtract as ONE packed instruction each (sm_100 FADD2 / FFMA2 on the (re, im) register pair;
        // fma(v, -1, u) rounds once, exactly like u - v)
        const float2 u = a[OFF + i + j], v = a[OFF + i + j + H];
        a[OFF + i + j] = __fadd2_rn(u, v);
        const float2 d = __ffma2_rn(v, make_float2(-1.f, -1.f), u);
        const float dx = d.x, dy = d.y;
        const int tk = j * (32 / LEN);          // W_LEN^j = W_32^tk = (cos, -sin)
        if (tk == 0) a[OFF + i + j + H] = make_float2(dx, dy);
        else if (tk == 8) a[OFF + i + j + H] = make_float2(dy, -dx);
        else {
          const float c = w32c(tk), s = w32s(tk);
          a[OFF + i + j + H] = make_float2(fmaf(dx, c, dy * s), fmaf(dy, c, -dx * s));
        }
      }
    }
    DifStage<N, OFF, LEN / 2>::run(a);
  }
};
template <int N, int OFF>
struct DifStage<N, OFF, 1> { static __device__ __forceinline__ void run(float2*) {} };

template <int N, int OFF>
__device__ __forceinline__ void fft_dif(float2* a) { DifStage<N, OFF, N>::run(a); }

__device__ __forceinline__ float2 cmul(float2 v, float2 t) {
  return make_float2(fmaf(v.x, t.x, -v.y * t.y), fmaf(v.x, t.y, v.y * t.x));
}

struct Smem {
  float* span; float* win1; float* win2; float2* tw1; float2* tw2; float2* post1; float2* post2; float* dc1; float* dc2; float* scratch; float* stage;
  __device__ explicit Smem(float* base) {
    span = base; win1 = span + kSpanMax; win2 = win1 + 2048;
    tw1 = reinterpret_cast<float2*>(win2 + 1024); tw2 = tw1 + 1024;
    post1 = tw2 + 512; post2 = post1 + kPost1;
    dc1 = reinterpret_cast<float*>(post2 + kPost2); dc2 = dc1 + kPost1;
    scratch = dc2 + kPost2; stage = scratch + kWarps * kScratchPerWarp;
  }
};

// 32x32 transpose across the warp through a padded scratch: in: lane l holds v[k] for k=0..31 (value
// destined to lane k, slot l); out: lane k holds all 32 values, slot index = source lane.
template <class GET, class PUT>
__device__ __forceinline__ void warp_transpose(float* sc, int lane, GET get, PUT put) {
#pragma unroll
  for (int k = 0; k < 32; ++k) sc[k * 33 + lane] = get(k);
  __syncwarp();
#pragma unroll
  for (int l = 0; l < 32; ++l) put(l, sc[lane * 33 + l]);
  __syncwarp();
}

__device__ __forceinline__ float mel_out(const float* bins, const int* start, const int* cnt, const float* w, int stride,
                                         int m, float pw, float sc, float sh) {
  const int s = __ldg(start + m), c = __ldg(cnt + m);
  float acc = 0.f;
  for (int i = 0; i < c; ++i) acc = fmaf(bins[s + i], __ldg(w + m * stride + i), acc);
  const float y = acc * acc;
  // y^p as exp2(p * log2 y) on the MUFU pipe (lg2.approx + ex2.approx: ~3e-7 relative on the result for p ~ 0.2, the same
  // order as powf's own rounding; y = 0 -> exp2(-inf) = 0).  powf's argument reduction and special cases cost ~100 instructions
  // per call and there are 196 k calls per chunk: 18 % of this kernel's instruction stream.
  return fmaf(__powf(y, pw), sc, sh);
}

// ---- spectrogram 0: one 2048-sample frame per warp ---------------------------------------------
__device__ __forceinline__ void spec0_frame(const FrontendDev& P, const Smem& S, int warp, int lane, int tl) {
  float2 a[32];
  const float* xs = S.span + 278 * tl;
  // DC compensation: X[k] = mu*W[k] + DFT((x - mu) w)[k] holds for any mu; with mu = frame mean the fp32
  // rounding noise of the FFT scales with the AC part only (silent / DC-offset frames would otherwise drown
  // the 1e-6-level bins the x^0.23 compression amplifies).  W = DFT of the model's own window constants (fp64 at load).
  float mu = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    a[j] = *reinterpret_cast<const float2*>(xs + 2 * (lane + 32 * j));
    mu += a[j].x + a[j].y;
  }
  mu = warp_sum(mu) * (1.0f / 2048.0f);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float2 w = *reinterpret_cast<const float2*>(S.win1 + 2 * (lane + 32 * j));
    a[j] = make_float2((a[j].x - mu) * w.x, (a[j].y - mu) * w.y);
  }
  fft_dif<32, 0>(a);                                    // a[brev5(k2)] = sum_j z[l+32j] W32^(j k2)
#pragma unroll
  for (int k2 = 1; k2 < 32; ++k2) a[brev(k2, 5)] = cmul(a[brev(k2, 5)], S.tw1[k2 * 32 + lane]);
  float* sc = S.scratch + warp * kScratchPerWarp;
  warp_transpose(sc, lane, [&](int k2) { return a[brev(k2, 5)].x; }, [&](int l, float v) { a[l].x = v; });
  // NOTE: the .x slots were overwritten in natural order l; the .y values are still in bit-reversed slots.
  {
    float ay[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) ay[k2] = a[brev(k2, 5)].y;
    warp_transpose(sc, lane, [&](int k2) { return ay[k2]; }, [&](int l, float v) { a[l].y = v; });
  }
  fft_dif<32, 0>(a);                                    // lane = k2: a[brev5(k1)] = Z[k2 + 32 k1]
  const int src = (32 - lane) & 31;
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    if (k1 < P.nk1[0]) {                                // warp-uniform
      const float2 zk = a[brev(k1, 5)];
      const float2 mine = (lane == 0) ? a[brev((32 - k1) & 31, 5)] : a[brev(31 - k1, 5)];
      const float znx = __shfl_sync(0xffffffffu, mine.x, src), zny = __shfl_sync(0xffffffffu, mine.y, src);
      const int k = lane + 32 * k1;
      const float2 cs = S.post1[k];
      sc[k] = fmaf(mu, S.dc1[k], 0.5f * (zk.x + znx) + cs.x * (zk.y + zny) - cs.y * (zk.x - znx));
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int m = lane + 32 * q;
    S.stage[((95 - m) * 32 + tl) * 2 + 0] =
        mel_out(sc, P.mel_start[0], P.mel_cnt[0], P.mel_w[0], P.mel_stride[0], m, P.pow_exp[0], P.bn_scale[0], P.bn_shift[0]);
  }
  __syncwarp();
}

// ---- spectrogram 1: two 1024-sample frames per warp ----------------------------------------------
__device__ __forceinline__ void spec1_pair(const FrontendDev& P, const Smem& S, int warp, int lane, int t0, int ta, bool b_valid) {
  float2 a[32];
  const float* xa = S.span + 2 * t0 + 280 * ta;         // 280*(t0+ta) - 278*t0
  const float* xb = b_valid ? xa + 280 : xa;
  float mua = 0.f, mub = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int n = lane + 32 * j;
    a[j] = *reinterpret_cast<const float2*>(xa + 2 * n);
    a[16 + j] = *reinterpret_cast<const float2*>(xb + 2 * n);
    mua += a[j].x + a[j].y; mub += a[16 + j].x + a[16 + j].y;
  }
  mua = warp_sum(mua) * (1.0f / 1024.0f); mub = warp_sum(mub) * (1.0f / 1024.0f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 w = *reinterpret_cast<const float2*>(S.win2 + 2 * (lane + 32 * j));
    a[j] = make_float2((a[j].x - mua) * w.x, (a[j].y - mua) * w.y);
    a[16 + j] = make_float2((a[16 + j].x - mub) * w.x, (a[16 + j].y - mub) * w.y);
  }
  fft_dif<16, 0>(a);                                    // a[f*16 + brev4(k2)] = A_f[l][k2]
  fft_dif<16, 16>(a);
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float2 t = S.tw2[k2 * 32 + lane];
    a[brev(k2, 4)] = cmul(a[brev(k2, 4)], t);
    a[16 + brev(k2, 4)] = cmul(a[16 + brev(k2, 4)], t);
  }
  float* sc = S.scratch + warp * kScratchPerWarp;
  // destination lane d = k2 + 16 f receives A_f[l][k2] in slot l
  {
    float ay[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) ay[d] = a[(d >> 4) * 16 + brev(d & 15, 4)].y;
    warp_transpose(sc, lane, [&](int d) { return a[(d >> 4) * 16 + brev(d & 15, 4)].x; }, [&](int l, float v) { a[l].x = v; });
    warp_transpose(sc, lane, [&](int d) { return ay[d]; }, [&](int l, float v) { a[l].y = v; });
  }
  fft_dif<32, 0>(a);                                    // lane = k2 + 16 f: a[brev5(k1)] = Z_f[k2 + 16 k1]
  const int f = lane >> 4, k2 = lane & 15;
  const int src = ((16 - k2) & 15) + 16 * f;
  const float mu = f ? mub : mua;
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    if (k1 < P.nk1[1]) {
      const float2 zk = a[brev(k1, 5)];
      const float2 mine = (k2 == 0) ? a[brev((32 - k1) & 31, 5)] : a[brev(31 - k1, 5)];
      const float znx = __shfl_sync(0xffffffffu, mine.x, src), zny = __shfl_sync(0xffffffffu, mine.y, src);
      const int k = k2 + 16 * k1;
      const float2 cs = S.post2[k];
      sc[f * 512 + k] = fmaf(mu, S.dc2[k], 0.5f * (zk.x + znx) + cs.x * (zk.y + zny) - cs.y * (zk.x - znx));
    }
  }
  __syncwarp();
#pragma unroll
  for (int ff = 0; ff < 2; ++ff) {
    if (ff == 0 || b_valid) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int m = lane + 32 * q;
        S.stage[((95 - m) * 32 + ta + ff) * 2 + 1] =
            mel_out(sc + ff * 512, P.mel_start[1], P.mel_cnt[1], P.mel_w[1], P.mel_stride[1], m, P.pow_exp[1], P.bn_scale[1], P.bn_shift[1]);
      }
    }
  }
  __syncwarp();
}

template <int FMT>
__global__ void __launch_bounds__(kWarps * 32, 1)
frontend_kernel(const FrontendDev P, const void* __restrict__ pcm, const float* __restrict__ partial, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_f[];
  Smem S(smem_f);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * kFeFramesPerCta;
  const int nfr = min(kFeFramesPerCta, P.n_frames - t0);

  // tables (constants: loaded before the dependency wait so they overlap the previous kernel's tail)
  for (int i = tid; i < 2048; i += blockDim.x) S.win1[i] = __ldg(P.win[0] + i);
  for (int i = tid; i < 1024; i += blockDim.x) S.win2[i] = __ldg(P.win[1] + i);
  for (int i = tid; i < 1024; i += blockDim.x) S.tw1[i] = __ldg(P.tw[0] + i);
  for (int i = tid; i < 512; i += blockDim.x) S.tw2[i] = __ldg(P.tw[1] + i);
  for (int i = tid; i < kPost1; i += blockDim.x) S.post1[i] = __ldg(P.post[0] + i);
  for (int i = tid; i < kPost2; i += blockDim.x) S.post2[i] = __ldg(P.post[1] + i);
  for (int i = tid; i < kPost1; i += blockDim.x) S.dc1[i] = __ldg(P.win_dft[0] + i);
  for (int i = tid; i < kPost2; i += blockDim.x) S.dc2[i] = __ldg(P.win_dft[1] + i);
  pdl_trigger();
  pdl_wait();
  // chunk min / max from the partials
  // (coherent loads: the partials were written by the previous kernel, which may still have been running when this CTA
  // started — PDL, common.cuh)
  float mn = __ldcg(partial + (b * kMinMaxParts) * 2), mx = __ldcg(partial + (b * kMinMaxParts) * 2 + 1);
#pragma unroll
  for (int i = 1; i < kMinMaxParts; ++i) {
    mn = fminf(mn, __ldcg(partial + (b * kMinMaxParts + i) * 2));
    mx = fmaxf(mx, __ldcg(partial + (b * kMinMaxParts + i) * 2 + 1));
  }
  const float den = (mx - mn) + P.eps;                  // max(x - min) + eps, as the graph computes it

  // stage the PCM span (normalised) -------------------------------------------------------------
  const int span0 = 278 * t0;
  int span_len = (nfr - 1) * 278 + 2048;
  { const int e2 = 2 * t0 + (nfr - 1) * 280 + 1024; span_len = max(span_len, e2); }
  span_len = min((span_len + 3) & ~3, P.n_samples - span0);
  auto norm = [&](float x) { return (__fdiv_rn(x - mn, den) - P.center) * P.gain; };
  if (FMT == 0) {
    const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(pcm) + (size_t)b * P.n_samples + span0);
    for (int i = tid; i < span_len / 4; i += blockDim.x) {
      const float4 v = __ldg(src + i);
      *reinterpret_cast<float4*>(S.span + 4 * i) = make_float4(norm(v.x), norm(v.y), norm(v.z), norm(v.w));
    }
  } else {
    const int2* src = reinterpret_cast<const int2*>(static_cast<const int16_t*>(pcm) + (size_t)b * P.n_samples + span0);
    for (int i = tid; i < span_len / 4; i += blockDim.x) {
      const int2 v = __ldg(src + i);
      const float k = 1.0f / 32768.0f;
      const float x0 = (float)(short)(v.x & 0xffff) * k, x1 = (float)(short)(v.x >> 16) * k;
      const float x2 = (float)(short)(v.y & 0xffff) * k, x3 = (float)(short)(v.y >> 16) * k;
      *reinterpret_cast<float4*>(S.span + 4 * i) = make_float4(norm(x0), norm(x1), norm(x2), norm(x3));
    }
  }
  for (int i = span_len + tid; i < kSpanMax; i += blockDim.x) S.span[i] = 0.f;
  __syncthreads();

  // spectrogram 0: frames warp and warp+16; spectrogram 1: frames (2 warp, 2 warp + 1)
#pragma unroll 1
  for (int rep = 0; rep < 2; ++rep) {
    const int tl = warp + kWarps * rep;
    if (tl < nfr) spec0_frame(P, S, warp, lane, tl);
  }
  {
    const int ta = 2 * warp;
    if (ta < nfr) spec1_pair(P, S, warp, lane, t0, ta, ta + 1 < nfr);
  }
  __syncthreads();

  // coalesced store: for each mel row h, nfr frames x 2 channels are contiguous in the output
  for (int i = tid; i < 96 * kFeFramesPerCta; i += blockDim.x) {
    const int h = i >> 5, tl = i & 31;
    if (tl < nfr) {
      const float2 v = *reinterpret_cast<const float2*>(S.stage + (h * 32 + tl) * 2);
      *reinterpret_cast<float2*>(out + (((size_t)b * 96 + h) * P.n_frames + t0 + tl) * 2) = v;
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256)
minmax_partial_kernel(const void* __restrict__ pcm, int n_samples, float* __restrict__ partial) {
  const int b = blockIdx.y, part = blockIdx.x;
  const int per = n_samples / kMinMaxParts;            // 18000, multiple of 8
  float mn = INFINITY, mx = -INFINITY;
  pdl_trigger();
  pdl_wait();
  if (FMT == 0) {
    const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(pcm) + (size_t)b * n_samples + (size_t)part * per);
    for (int i = threadIdx.x; i < per / 4; i += blockDim.x) {
      const float4 v = __ldg(src + i);
      mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
  } else {
    const int4* src = reinterpret_cast<const int4*>(static_cast<const int16_t*>(pcm) + (size_t)b * n_samples + (size_t)part * per);
    int imn = 32767, imx = -32768;
    for (int i = threadIdx.x; i < per / 8; i += blockDim.x) {
      const int4 v = __ldg(src + i);
      const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int lo = (short)(w[q] & 0xffff), hi = (short)(w[q] >> 16);
        imn = min(imn, min(lo, hi)); imx = max(imx, max(lo, hi));
      }
    }
    mn = (float)imn * (1.0f / 32768.0f); mx = (float)imx * (1.0f / 32768.0f);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  __shared__ float smn[8], smx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    partial[(b * kMinMaxParts + part) * 2] = mn;
    partial[(b * kMinMaxParts + part) * 2 + 1] = mx;
  }
}

}  // namespace

size_t frontend_smem_bytes() { return (size_t)kSmemFloats * sizeof(float); }

// per-device attribute (called once per device by tc_prepare_device, engine.cu)
void frontend_set_attributes() {
  BNB_CUDA(cudaFuncSetAttribute(frontend_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frontend_smem_bytes()));
  BNB_CUDA(cudaFuncSetAttribute(frontend_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frontend_smem_bytes()));
}

void launch_minmax(const void* pcm, int fmt, int B, int n_samples, float* partial, cudaStream_t s, LaunchCounter& lc) {
  dim3 grid(kMinMaxParts, B);
  if (fmt == 0) launch_k(minmax_partial_kernel<0>, grid, dim3(256), 0, s, pcm, n_samples, partial);
  else launch_k(minmax_partial_kernel<1>, grid, dim3(256), 0, s, pcm, n_samples, partial);
  BNB_LAUNCH_CHECK(lc);
}

void launch_frontend(const FrontendDev& fe, const void* pcm, int fmt, int B, const float* partial, float* out,
                     cudaStream_t s, LaunchCounter& lc) {
  const size_t smem = frontend_smem_bytes();
  dim3 grid(ceil_div(fe.n_frames, kFeFramesPerCta), B);
  if (fmt == 0) launch_k(frontend_kernel<0>, grid, dim3(kWarps * 32), smem, s, fe, pcm, partial, out);
  else launch_k(frontend_kernel<1>, grid, dim3(kWarps * 32), smem, s, fe, pcm, partial, out);
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
