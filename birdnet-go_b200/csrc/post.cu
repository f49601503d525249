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

// N1: compaction of the per-chunk top-k into one detection list.  conf[b][0..k) is sorted descending (NaN slots last), so the
// detections of a chunk are the entries >= threshold; chunk b's run starts at the exclusive prefix sum of the counts — chunk
// order, then confidence order: what the reference's per-chunk loop over Results produces (processor.go:820-876), deterministic.
__global__ void __launch_bounds__(1024)
compact_detections_kernel(const int32_t* __restrict__ idx, const float* __restrict__ conf, int B, int k, float threshold, int max_det,
                          int32_t* __restrict__ det_chunk, int32_t* __restrict__ det_idx, float* __restrict__ det_conf,
                          int32_t* __restrict__ counts, int32_t* __restrict__ n_det) {
  __shared__ int s_scan[1024];
  __shared__ int s_base;
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    int cnt = 0;
    if (b < B) for (int j = 0; j < k; ++j) cnt += __ldcg(conf + (size_t)b * k + j) >= threshold ? 1 : 0;
    s_scan[tid] = cnt;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {                     // inclusive Hillis-Steele scan
      const int v = tid >= d ? s_scan[tid - d] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int base = s_base;
    if (b < B) {
      counts[b] = cnt;
      int o = base + s_scan[tid] - cnt;
      for (int j = 0; j < k; ++j) {
        const float c = __ldcg(conf + (size_t)b * k + j);
        if (c >= threshold) {
          if (o < max_det) { det_chunk[o] = b; det_idx[o] = __ldcg(idx + (size_t)b * k + j); det_conf[o] = c; }
          ++o;
        }
      }
    }
    __syncthreads();
    if (tid == 1023) s_base = base + s_scan[1023];
    __syncthreads();
  }
  if (tid == 0) *n_det = s_base;
}

// bat pipeline's custom head (N2): conf[b][c] = sigmoid(bias[c] + sum_i emb[b][i] * w[c][i]); one warp per output, fp32 with
// a fixed lane-strided summation order + shuffle tree (reference: a small ONNX dense graph whose raw outputs get a plain
// sigmoid in Go, internal/inference/onnx/custom_classifier.go:147-173)
__global__ void __launch_bounds__(256)
dense_head_kernel(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ bias, int B, int n_in, int n_out,
                  float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * n_out) return;
  const int b = warp / n_out, c = warp - b * n_out;
  float acc = 0.f;
  for (int i = lane; i < n_in; i += 32) acc = fmaf(emb[(size_t)b * n_in + i], __ldg(w + (size_t)c * n_in + i), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * n_out + c] = 1.0f / (1.0f + expf(-(acc + __ldg(bias + c))));
}

}  // namespace

void launch_compact_detections(const int32_t* idx, const float* conf, int B, int k, float threshold, int max_det, int32_t* det_chunk,
                               int32_t* det_idx, float* det_conf, int32_t* counts, int32_t* n_det, cudaStream_t s, LaunchCounter& lc) {
  launch_k(compact_detections_kernel, dim3(1), dim3(1024), 0, s, idx, conf, B, k, threshold, max_det, det_chunk, det_idx, det_conf, counts, n_det);
  lc.n++;
}

void launch_dense_head(const float* emb, const float* w, const float* bias, int B, int n_in, int n_out, float* out, cudaStream_t s) {
  const long long warps = (long long)B * n_out;
  dense_head_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(emb, w, bias, B, n_in, n_out, out);
  BNB_CUDA(cudaGetLastError());
}

void launch_sigmoid_topk(const float* logits, int B, int n, float sensitivity, int k, int32_t* idx, float* conf,
                         cudaStream_t s, LaunchCounter& lc) {
  launch_k(sigmoid_topk_kernel, dim3(B), dim3(kTopkThreads), (size_t)n * sizeof(float), s, logits, n, sensitivity, k, idx, conf);
  BNB_LAUNCH_CHECK(lc);
}

}  // namespace bnb
