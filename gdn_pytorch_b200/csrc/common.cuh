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

// Programmatic dependent launch (PDL).  Every kernel of the step chain starts with pdl_trigger() -- "the next kernel of
// this stream may be scheduled now": its launch latency and its prologue (barrier init, tensor-memory allocation,
// tensor-map prefetch) overlap THIS kernel -- and executes pdl_wait() before it touches global memory: the wait returns
// when the preceding kernel of the stream has completed and flushed.  Because every kernel launched through
// launch_pdl() waits before it finishes, completion stays transitive along the stream (kernel N+2 never overtakes N).
// Without the launch attribute (the default: GDN_PDL=1 opts in, see pdl_enabled(); or a launch that does not go through
// launch_pdl) both are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// BatchNorm finalisation of ONE channel from its batch sums -- shared by bn_finalize_kernel (elementwise.cu) and by the
// last CTA of a convolution (conv_igemm.cu, gdn_conv_desc.fin_*).  Every operation is an explicit round-to-nearest
// intrinsic, so the compiler cannot contract multiply-adds differently in the two kernels: both produce the same bits.
#ifdef __CUDACC__
struct BnFinOut {
  float scale, shift, mean, rstd, unbiased_var;
};
__device__ __forceinline__ BnFinOut bn_finalize_channel(double sum, double sq, double count, float gamma, float beta, float eps) {
  const double m = __ddiv_rn(sum, count);
  double var = __dsub_rn(__ddiv_rn(sq, count), __dmul_rn(m, m));
  if (var < 0) var = 0;
  BnFinOut o;
  o.rstd = (float)__ddiv_rn(1.0, sqrt(__dadd_rn(var, (double)eps)));
  o.mean = (float)m;
  o.scale = __fmul_rn(gamma, o.rstd);
  o.shift = __fsub_rn(beta, __fmul_rn(__fmul_rn(o.mean, gamma), o.rstd));
  o.unbiased_var = (float)(count > 1 ? __ddiv_rn(__dmul_rn(var, count), __dsub_rn(count, 1.0)) : var);
  return o;
}
__device__ __forceinline__ float bn_running_update(float running, float momentum, float value) {
  return __fadd_rn(__fmul_rn(__fsub_rn(1.f, momentum), running), __fmul_rn(momentum, value));
}
#endif
bool pdl_enabled();
// GDN_DETERMINISTIC=1: every fp32 reduction runs in a fixed order (see det_enabled() in common.cu)
bool det_enabled();

// Launch `kern` (which MUST execute pdl_wait()) on `st`, optionally as clusters of `cluster_x` CTAs.
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                              A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned na = 0;
  if (cluster_x > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster_x;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    na++;
  }
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in bytes for dims 1..rank-1.
int encode_tmap_bf16(CUtensorMap* out, void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);

}  // namespace gdn
