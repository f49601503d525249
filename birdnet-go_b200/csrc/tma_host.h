// tma_host.h — host-side CUtensorMap encoding (cuTensorMapEncodeTiled fetched through the runtime: no link-time
// dependency on libcuda), cached per (pointer, geometry).  Thread-safe.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bnb {

// swizzle: 0 none, 32 / 64 / 128 bytes.  elem_bytes: 2 (fp16) or 4 (fp32).  dims / box innermost first; strides in
// bytes for dims 1..rank-1.  Out-of-range box elements are zero-filled.
CUtensorMap tma_encode(const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle);

}  // namespace bnb
