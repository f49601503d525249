// post.cu — BirdNET.Predict's post-processing on the device.
//
// Restates /root/reference/internal/classifier/analyze.go:82-99:
//   confidence[i] = float32( 1 / (1 + exp(-sensitivity * float64(logit[i]))) )      (:113-115, :197-208)
//   top-k (k = 10) by descending confidence                                           (:220-253)
// The reference's quick-select is unstable among exactly equal confidences; here ties go to the
// lower label index (documented deviation, same rule as the oracle).
// One CTA per chunk; the 6522 confidences live in shared memory, k arg-max rounds.
#include "kernels.h"

namespace bnb {

namespace {

constexpr int kTopkThreads = 256;

__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

__global__ void __launch_bounds__(kTopkThreads)
sigmoid_topk_kernel(const float* __restrict__ logits, int n, float sensitivity, int k, int32_t* __restrict__ idx,
                    float* __restrict__ conf) {
  extern __shared__ float s_conf[];
  __shared__ float s_v[kTopkThreads / 32];
  __shared__ int s_i[kTopkThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)b * n;
  const double sens = (double)sensitivity;
  pdl_trigger();
  pdl_wait();
  // NaN logits (a NaN/Inf PCM sample propagates to every logit) become confidence NaN in the reference too; here they
  // rank below every real confidence (-1 sentinel) so the arg-max rounds below always find an in-range index
  for (int i = tid; i < n; i += kTopkThreads) {
    const float c = (float)(1.0 / (1.0 + exp(-sens * (double)__ldcg(x + i))));
    s_conf[i] = (c == c) ? c : -1.f;
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bv = -3.f; int bi = 0x7fffffff;
    for (int i = tid; i < n; i += kTopkThreads) { const float v = s_conf[i]; if (better(v, i, bv, bi)) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kTopkThreads / 32; ++w) if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
      if (r < n && bi >= 0 && bi < n) {
        idx[(size_t)b * k + r] = bi;
        conf[(size_t)b * k + r] = bv < 0.f ? __int_as_float(0x7fc00000) : bv;      // a consumed / NaN slot reports NaN, like the reference would
        s_conf[bi] = -2.f;
      } else { idx[(size_t)b * k + r] = -1; conf[(size_t)b * k + r] = 0.f; }
    }
    __syncthreads();
  }
}

}  // namespace

void launch_sigmoid_topk(const float* logits, int B, int n, float sensitivity, int k, int32_t* idx, float* conf,
                         cudaStream_t s, LaunchCounter& lc) {
  launch_k(sigmoid_topk_kernel, dim3(B), dim3(kTopkThreads), (size_t)n * sizeof(float), s, logits, n, sensitivity, k, idx, conf);
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
