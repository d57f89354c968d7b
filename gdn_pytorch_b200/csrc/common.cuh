// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/gdn_b200.h"

namespace gdn {

void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);
int device_sm_count();

#define GDN_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return ::gdn::fail(GDN_CUDA_ERROR, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define GDN_LAUNCH_CHECK(name)                                                                \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) return ::gdn::fail(GDN_CUDA_ERROR, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in bytes for dims 1..rank-1.
int encode_tmap_bf16(CUtensorMap* out, void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

}  // namespace gdn
