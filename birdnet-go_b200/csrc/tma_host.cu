// tma_host.cu — see tma_host.h
#include "tma_host.h"

#include <string.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace bnb {

namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::mutex g_mu;
EncodeFn g_fn = nullptr;
std::map<std::vector<uint64_t>, CUtensorMap> g_cache;
}  // namespace

CUtensorMap tma_encode(const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    BNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    g_fn = reinterpret_cast<EncodeFn>(p);
  }
  std::vector<uint64_t> key;
  key.push_back((uint64_t)(uintptr_t)base); key.push_back((uint64_t)elem_bytes * 1000 + (uint64_t)rank * 10000 + (uint64_t)swizzle);
  for (int i = 0; i < rank; ++i) { key.push_back(dims[i]); key.push_back(box[i]); if (i > 0) key.push_back(strides_bytes[i - 1]); }
  auto it = g_cache.find(key);
  if (it != g_cache.end()) return it->second;
  CUtensorMap m;
  cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; if (i > 0) st[i - 1] = strides_bytes[i - 1]; }
  const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = g_fn(&m, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  if (g_cache.size() > 8192) g_cache.clear();
  g_cache[key] = m;
  return m;
}

}  // namespace bnb
