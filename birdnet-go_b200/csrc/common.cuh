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

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------------
// The chain of one micro-batch is ~130 dependent kernels on one stream; with ordinary launches every boundary costs the drain
// of the previous grid plus the ramp-up of the next one (a few microseconds, and the persistent tensor-core kernels have long
// tails: 44 of 148 CTAs of a project GEMM run a second tile).  Launched with the programmatic-serialization attribute, the
// NEXT kernel's CTAs start on SMs as they free up, run their prologue (barrier init, TMEM allocation, weight loads: nothing that
// depends on the previous kernel), and block in pdl_wait() until the previous grid has completed and its writes are visible.
// Rules every kernel launched through launch_k() follows: (1) pdl_wait() before the first read of anything another kernel wrote
// AND before the first global write (the previous kernel may still be reading that buffer); (2) every CTA executes pdl_wait()
// (completion of grid N then implies completion of grid N-1); (3) pdl_trigger() as early as possible.
// BNB_PDL=0 launches everything fully serialized (pdl_wait / pdl_trigger are then no-ops).
bool pdl_enabled();
template <class... KArgs, class... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) throw ::bnb::cuda_error(e, "kernel launch", __FILE__, __LINE__);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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
