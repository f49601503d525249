// common.cuh — shared helpers for the sm_100a kernels of libbirdnet_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace bnb {

struct cuda_error : std::runtime_error {
  cudaError_t code;
  cuda_error(cudaError_t c, const char* what, const char* file, int line)
      : std::runtime_error(std::string(what) + ": " + cudaGetErrorString(c) + " (" + file + ":" + std::to_string(line) + ")"), code(c) {}
};

#define BNB_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) throw ::bnb::cuda_error(_e, #expr, __FILE__, __LINE__); \
  } while (0)

// launch bookkeeping: every kernel launch goes through this so bnb_kernel_launches() is exact
struct LaunchCounter { long long n = 0; };

#define BNB_LAUNCH_CHECK(counter)                                             \
  do {                                                                        \
    (counter).n++;                                                            \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) throw ::bnb::cuda_error(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace bnb
