// How fast does one SM pull [128 rows x C fp32] boxes of a row-major matrix (pitch 1152 B, L2-resident) through TMA?
// Varies the box width C (bytes per row request) and the number of boxes in flight per CTA.  148 CTAs stream concurrently.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/tma_box_bench.cu -o /tmp/tma_box_bench -lcuda && /tmp/tma_box_bench
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int box_cols, int m_tiles, int k_boxes, int inflight, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  const uint32_t box_bytes = 128u * box_cols * 4u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  // sequence of boxes for this CTA: (m tile, k box) pairs, m tiles strided over the grid
  int issued = 0, done = 0, total = 0;
  for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) total += k_boxes;
  int mt_i = blockIdx.x, kb_i = 0;
  while (done < total) {
    while (issued < total && issued - done < inflight) {
      const int s = issued % inflight;
      mbar_expect(smem_u32(&bars[s]), box_bytes);
      tma_load_2d(smem_u32(smem) + s * box_bytes, &map, kb_i * box_cols, mt_i * 128, smem_u32(&bars[s]));
      if (++kb_i == k_boxes) { kb_i = 0; mt_i += gridDim.x; }
      ++issued;
    }
    const int s = done % inflight;
    mbar_wait(smem_u32(&bars[s]), (done / inflight) & 1);
    ++done;
  }
  cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const int M = 24576, K = 288;                       // block-5 project operand: 28 MB, fits L2
  float* A; cudaMalloc(&A, (size_t)M * K * 4); cudaMemset(A, 0, (size_t)M * K * 4);
  long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cols_list[] = {32, 64, 96, 144, 288}, infl_list[] = {1, 2, 3, 6};
  for (int cols : cols_list) for (int infl : infl_list) {
    if ((size_t)infl * 128 * cols * 4 > 196 * 1024) continue;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)cols, 128}; const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed cols=%d (%d)\n", cols, (int)r); continue; }
    const int k_boxes = (K + cols - 1) / cols;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k<<<148, 128, (size_t)infl * 128 * cols * 4 + 1024>>>(m, cols, M / 128, k_boxes, infl, cyc);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double boxes_per_cta = (double)(M / 128) * k_boxes / 148.0;
    printf("box 128 x %3d fp32 (%4d B rows, %3d KB)  in flight %d : %7.1f us  %6.2f TB/s  %7.0f cycles per box per CTA (%.0f cycles x in-flight)\n", cols, cols * 4,
           128 * cols * 4 / 1024, infl, ms * 1e3, (double)M * K * 4 / (ms * 1e-3) / 1e12, mx / boxes_per_cta, mx / boxes_per_cta * infl);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
