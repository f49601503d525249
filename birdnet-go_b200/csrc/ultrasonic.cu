// ultrasonic.cu — SURVEY §8(f) N2: the bat pipeline's post-detection validation filter on the GPU.
//
// What the reference computes (/root/reference/internal/audiocore/ultrasonic/filter.go:20-66, called once per chunk with bat
// detections from internal/analysis/processor/processor.go:892-935 on the chunk's int16 PCM at the SOURCE rate, converted by
// convert.BytesToFloat64PCM16 = int16 / 32768 in float64):
//   window = symmetric Hann 0.5 (1 - cos(2 pi i / (n - 1)))            filter.go:139-145
//   frames = 1 + (len - fft) / hop, each: x * window -> complex FFT (float64, radix-2)   :49-54
//   power[frame] = sum over bins split..nyquist of |X|^2, doubled for 0 < bin < nyquist      :56-64
//   CV = sqrt(mean((p - mean p)^2)) / mean p   (0 when mean <= 0)                           :76-97
//   ok = false for short input, fft not a power of two, hop <= 0, split outside [0, rate/2), fewer than two frames   :21-39
// Defaults: FFT 8192, hop 4096, split 20 kHz, threshold 0.15 (conf/defaults.go:108-112) -> 34 frames per 144000-sample chunk.
//
// Kernel design (float64 like the reference: the CV of a flat tone is ~1e-3, far below what fp32 butterflies would keep):
//   us_twiddle_kernel      exp(-2 pi i j / fft), j < fft / 2, with sincospi (the reference multiplies a running twiddle, which
//                          drifts by ~1e-13; direct evaluation is the more accurate of the two and the test tolerance says so)
//   us_frame_power_kernel  one CTA = one frame of one chunk: window + bit reversal into a shared-memory complex128 buffer
//                          (fft * 16 B = 128 KB for 8192), log2(fft) butterfly stages, band power by a fixed-order tree
//   us_cv_kernel           one thread per chunk walks its frame powers sequentially — exactly the reference's summation order
#include <math.h>

#include <stdexcept>

#include "kernels.h"

namespace bnb {

namespace {

constexpr int kUsThreads = 512;

__global__ void us_twiddle_kernel(double2* tw, int fft) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= fft / 2) return;
  double s, c;
  sincospi(-2.0 * (double)j / (double)fft, &s, &c);
  tw[j] = make_double2(c, s);
}

// pcm: int16 (fmt 1, scaled by 1/32768 as BytesToFloat64PCM16 does) or float32 (fmt 0) samples, [B][n_samples]
__global__ void __launch_bounds__(kUsThreads)
us_frame_power_kernel(const void* __restrict__ pcm, int fmt, int n_samples, int fft, int log2n, int hop, int n_frames, int split_bin,
                      const double2* __restrict__ tw, double* __restrict__ power) {
  extern __shared__ double2 us_buf[];
  __shared__ double s_red[kUsThreads];
  const int frame = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const size_t off = (size_t)b * n_samples + (size_t)frame * hop;
  const double inv_nm1 = 1.0 / (double)(fft - 1);
  for (int i = tid; i < fft; i += kUsThreads) {
    const double x = fmt == 1 ? (double)static_cast<const int16_t*>(pcm)[off + i] / 32768.0 : (double)static_cast<const float*>(pcm)[off + i];
    const double w = 0.5 * (1.0 - cospi(2.0 * (double)i * inv_nm1));
    us_buf[__brev((unsigned)i) >> (32 - log2n)] = make_double2(x * w, 0.0);
  }
  __syncthreads();
  for (int lh = 0; lh < log2n; ++lh) {                      // half = 1 << lh, size = 2 * half
    const int half = 1 << lh;
    for (int t = tid; t < fft / 2; t += kUsThreads) {
      const int k = t & (half - 1), start = (t >> lh) << (lh + 1);
      const double2 w = tw[(size_t)k << (log2n - 1 - lh)];   // exp(-2 pi i k / size)
      const double2 u = us_buf[start + k], x = us_buf[start + k + half];
      const double2 v = make_double2(w.x * x.x - w.y * x.y, w.x * x.y + w.y * x.x);
      us_buf[start + k] = make_double2(u.x + v.x, u.y + v.y);
      us_buf[start + k + half] = make_double2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
  }
  const int nyq = fft / 2;
  double acc = 0.0;
  for (int bin = split_bin + tid; bin <= nyq; bin += kUsThreads) {
    const double2 z = us_buf[bin];
    double p = z.x * z.x + z.y * z.y;
    if (bin > 0 && bin < nyq) p *= 2.0;
    acc += p;
  }
  s_red[tid] = acc;
  __syncthreads();
  for (int s = kUsThreads / 2; s > 0; s >>= 1) {             // fixed-order tree: the same bits whatever else runs
    if (tid < s) s_red[tid] += s_red[tid + s];
    __syncthreads();
  }
  if (tid == 0) power[(size_t)b * n_frames + frame] = s_red[0];
}

__global__ void us_cv_kernel(const double* __restrict__ power, int n_frames, int B, double* __restrict__ cv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* p = power + (size_t)b * n_frames;
  const double n = (double)n_frames;
  double sum = 0.0;
  for (int i = 0; i < n_frames; ++i) sum += p[i];
  const double mean = sum / n;
  double out = 0.0;
  if (mean > 0.0) {
    double sq = 0.0;
    for (int i = 0; i < n_frames; ++i) { const double d = p[i] - mean; sq += d * d; }
    out = sqrt(sq / n) / mean;
  }
  cv[b] = out;
}

}  // namespace

// Parameter checks of ComputeUSFrameCV (filter.go:21-39): returns the frame count, or 0 when the reference would return ok = false.
int ultrasonic_frames(int n_samples, int sample_rate, int fft, int hop, int split_hz) {
  if (n_samples < fft || sample_rate <= 0 || fft < 2 || hop <= 0) return 0;
  if (fft & (fft - 1)) return 0;
  if (split_hz < 0 || split_hz >= sample_rate / 2) return 0;
  const int frames = 1 + (n_samples - fft) / hop;
  return frames < 2 ? 0 : frames;
}

size_t ultrasonic_workspace_bytes(int B, int n_frames, int fft) {
  return (size_t)(fft / 2) * sizeof(double2) + (size_t)B * n_frames * sizeof(double) + (size_t)B * sizeof(double);
}

// d_pcm: device samples [B][n_samples]; workspace: ultrasonic_workspace_bytes(); d_cv: [B] doubles (inside the workspace is fine)
void launch_ultrasonic_cv(const void* d_pcm, int fmt, int B, int n_samples, int sample_rate, int fft, int hop, int split_hz,
                          void* workspace, double* d_cv, cudaStream_t s) {
  const int n_frames = ultrasonic_frames(n_samples, sample_rate, fft, hop, split_hz);
  if (n_frames == 0) throw std::invalid_argument("ultrasonic filter: parameters for which the reference reports ok = false");
  if (fft > 8192) throw std::invalid_argument("ultrasonic filter: FFT sizes above 8192 do not fit the shared-memory buffer");
  int log2n = 0;
  while ((1 << log2n) < fft) ++log2n;
  const double bin_width = (double)sample_rate / (double)fft;
  const int split_bin = (int)((double)split_hz / bin_width);
  double2* tw = static_cast<double2*>(workspace);
  double* power = reinterpret_cast<double*>(tw + fft / 2);
  const size_t smem = (size_t)fft * sizeof(double2);
  BNB_CUDA(cudaFuncSetAttribute(us_frame_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8192 * sizeof(double2))));
  us_twiddle_kernel<<<ceil_div(fft / 2, 256), 256, 0, s>>>(tw, fft);
  us_frame_power_kernel<<<dim3(n_frames, B), kUsThreads, smem, s>>>(d_pcm, fmt, n_samples, fft, log2n, hop, n_frames, split_bin, tw, power);
  us_cv_kernel<<<ceil_div(B, 128), 128, 0, s>>>(power, n_frames, B, d_cv);
  BNB_CUDA(cudaGetLastError());
}

}  // namespace bnb
