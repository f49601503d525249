// FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a: 8 independent chains per thread, 1024 threads x 148 x 2 CTAs.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/ffma2_bench.cu -o /tmp/ffma2_bench && /tmp/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, a2.x, b2.x); x[i].y = fmaf(x[i].y, a2.y, b2.y); }
      else x[i] = __ffma2_rn(x[i], a2, b2);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 296 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<296, 1024>>>(out, 0.999f, 0.001f, iters); else k<1><<<296, 1024>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("%s: %.3f ms  %.1f TFLOP/s fp32\n", mode ? "FFMA2" : "FFMA ", ms, 296.0 * 1024 * iters * 16 * 2 / ms * 1e-9);
    }
  }
  return 0;
}
