// Demo path: the image resizing around the B=1 network call of
// /root/reference/src/depth_extract.py:23-58,86,138 -- scipy.misc.imresize(arr, size, 'bilinear') =
// bytescale (min-max stretch to uint8) + PIL's BILINEAR resize of an 8-bit image -- on the device, bit-exact.
//
// PIL (src/libImaging/Resample.c): separable; horizontal pass first, then vertical, through an 8-bit intermediate;
// triangle filter with support max(in/out, 1) (antialiased down-scaling); per output the coefficients are normalised
// in double precision and quantised to PRECISION_BITS = 22 fixed point; output = clip8((2^21 + sum k*p) >> 22).
// The coefficient tables are built on the device with round-to-nearest double intrinsics (no FMA contraction), so
// they equal the host library's; the passes are integer arithmetic.  HBM/latency-bound: a 375x1242x3 image is 1.4 MB.
#include <cstdint>
#include "common.cuh"

namespace gdn {

constexpr int kPrecisionBits = 32 - 8 - 2;

// ---------------------------------------------------------------------------------------------- bytescale
__device__ __forceinline__ unsigned int f2ord(float f) {   // order-preserving float -> uint map
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(unsigned int* mm) {
  mm[0] = 0xffffffffu;   // running min (ordered encoding)
  mm[1] = 0u;            // running max
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ src, long long n, unsigned int* mm) {
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int o = f2ord((float)src[i]);
    lo = min(lo, o);
    hi = max(hi, o);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 0, lo);
    atomicMax(mm + 1, hi);
  }
}

// bytescale: uint8(clip((v - cmin) * scale, 0, 255) + 0.5).  F64 = false: float32 arithmetic with
// scale = float(255.0 / double(cmax - cmin)) (a float32 array times a scalar, the input image of the demo);
// F64 = true: the same in double precision (the demo copies the float32 depth map into a float64 array first,
// depth_extract.py:135-136, so NumPy evaluates bytescale in float64 there).
template <typename T, bool F64>
__global__ void __launch_bounds__(256) bytescale_kernel(const T* __restrict__ src, long long n, const unsigned int* mm,
                                                        uint8_t* __restrict__ dst) {
  const float cmin = ord2f(mm[0]), cmax = ord2f(mm[1]);
  if (F64) {
    double cscale = __dsub_rn((double)cmax, (double)cmin);
    if (cscale == 0.0) cscale = 1.0;
    const double scale = __ddiv_rn(255.0, cscale);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      double b = __dmul_rn(__dsub_rn((double)src[i], (double)cmin), scale);
      b = fmin(fmax(b, 0.0), 255.0);
      dst[i] = (uint8_t)(int)__dadd_rn(b, 0.5);
    }
  } else {
    float cscale = __fsub_rn(cmax, cmin);
    if (cscale == 0.f) cscale = 1.f;
    const float scale = (float)__ddiv_rn(255.0, (double)cscale);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      float b = __fmul_rn(__fsub_rn((float)src[i], cmin), scale);
      b = fminf(fmaxf(b, 0.f), 255.f);
      dst[i] = (uint8_t)(int)__fadd_rn(b, 0.5f);
    }
  }
}

// ------------------------------------------------------------------------------------- PIL bilinear resize
// one thread per output index xx: bounds[2*xx] = xmin, bounds[2*xx+1] = count, kk[xx*ksize + x] = fixed-point weight
__global__ void resize_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn((double)((float)in_size - 0.0f), (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // triangle filter: support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; x++) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (a < 0.0) a = -a;
    const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  int* k = kk + (size_t)xx * ksize;
  for (int x = 0; x < ksize; x++) {
    double w = 0.0;
    if (x < xmax) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
    }
    const double q = __dmul_rn(w, (double)(1 << kPrecisionBits));
    k[x] = w < 0.0 ? (int)__dadd_rn(-0.5, q) : (int)__dadd_rn(0.5, q);
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

// One separable pass.  The image is viewed as [outer][len][inner] bytes with the resized axis in the middle:
// horizontal: outer = n*h, len = w, inner = c;  vertical: outer = n, len = h, inner = w*c.
__global__ void __launch_bounds__(256) resize_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          long long outer, int in_len, int out_len, int inner, int ksize,
                                                          const int* __restrict__ bounds, const int* __restrict__ kk) {
  const long long total = outer * out_len * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int in = (int)(i % inner);
    const long long t = i / inner;
    const int xx = (int)(t % out_len);
    const long long o = t / out_len;
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const int* k = kk + (size_t)xx * ksize;
    const uint8_t* p = src + ((size_t)o * in_len + xmin) * inner + in;
    int acc = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < cnt; x++) acc += (int)p[(size_t)x * inner] * __ldg(k + x);
    acc >>= kPrecisionBits;
    dst[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

static int rs_ksize(int in_size, int out_size) {
  double scale = (double)((float)in_size - 0.0f) / out_size;
  if (scale < 1.0) scale = 1.0;
  return (int)ceil(scale) * 2 + 1;
}

static size_t rs_align(size_t v) { return (v + 255) / 256 * 256; }

static int rs_grid(long long work) {
  long long b = (work + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace gdn

using namespace gdn;
#define GDN_API extern "C" __attribute__((visibility("default")))

GDN_API int gdn_bytescale(const void* src, int src_is_u8, int64_t n, int f64_math, uint8_t* dst, void* scratch8,
                          gdn_stream stream) {
  if (!src || !dst || !scratch8 || n < 1) return fail(GDN_INVALID_DESC, "gdn_bytescale: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* mm = reinterpret_cast<unsigned int*>(scratch8);
  minmax_init_kernel<<<1, 1, 0, st>>>(mm);
  const int grid = rs_grid(n);
  if (src_is_u8) {
    const uint8_t* p = (const uint8_t*)src;
    minmax_kernel<uint8_t><<<grid, 256, 0, st>>>(p, n, mm);
    if (f64_math) bytescale_kernel<uint8_t, true><<<grid, 256, 0, st>>>(p, n, mm, dst);
    else bytescale_kernel<uint8_t, false><<<grid, 256, 0, st>>>(p, n, mm, dst);
  } else {
    const float* p = (const float*)src;
    minmax_kernel<float><<<grid, 256, 0, st>>>(p, n, mm);
    if (f64_math) bytescale_kernel<float, true><<<grid, 256, 0, st>>>(p, n, mm, dst);
    else bytescale_kernel<float, false><<<grid, 256, 0, st>>>(p, n, mm, dst);
  }
  GDN_LAUNCH_CHECK("bytescale kernels");
  return GDN_OK;
}

GDN_API size_t gdn_resize_u8_workspace(int n, int h, int w, int c, int oh, int ow) {
  if (n < 1 || h < 1 || w < 1 || c < 1 || oh < 1 || ow < 1) return 0;
  const size_t tab_h = rs_align((size_t)ow * (2 + rs_ksize(w, ow)) * sizeof(int));
  const size_t tab_v = rs_align((size_t)oh * (2 + rs_ksize(h, oh)) * sizeof(int));
  return tab_h + tab_v + rs_align((size_t)n * h * ow * c);
}

GDN_API int gdn_resize_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int c, int oh, int ow, void* workspace,
                          size_t ws_bytes, gdn_stream stream) {
  if (!src || !dst || n < 1 || h < 1 || w < 1 || c < 1 || c > 4 || oh < 1 || ow < 1)
    return fail(GDN_INVALID_DESC, "gdn_resize_u8: bad arguments (n=%d %dx%dx%d -> %dx%d)", n, h, w, c, oh, ow);
  const size_t need = gdn_resize_u8_workspace(n, h, w, c, oh, ow);
  if (!workspace || ws_bytes < need)
    return fail(GDN_WORKSPACE_TOO_SMALL, "gdn_resize_u8: workspace %zu < %zu bytes", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int kh_ = rs_ksize(w, ow), kv_ = rs_ksize(h, oh);
  char* ws = reinterpret_cast<char*>(workspace);
  int* bounds_h = reinterpret_cast<int*>(ws);
  int* kk_h = bounds_h + 2 * (size_t)ow;
  ws += rs_align((size_t)ow * (2 + kh_) * sizeof(int));
  int* bounds_v = reinterpret_cast<int*>(ws);
  int* kk_v = bounds_v + 2 * (size_t)oh;
  ws += rs_align((size_t)oh * (2 + kv_) * sizeof(int));
  uint8_t* tmp = reinterpret_cast<uint8_t*>(ws);
  const bool need_h = ow != w, need_v = oh != h;
  if (!need_h && !need_v) {
    GDN_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * h * w * c, cudaMemcpyDeviceToDevice, st));
    return GDN_OK;
  }
  const uint8_t* cur = src;
  if (need_h) {
    resize_coeffs_kernel<<<(ow + 127) / 128, 128, 0, st>>>(w, ow, kh_, bounds_h, kk_h);
    uint8_t* out = need_v ? tmp : dst;
    resize_pass_kernel<<<rs_grid((long long)n * h * ow * c), 256, 0, st>>>(cur, out, (long long)n * h, w, ow, c, kh_, bounds_h, kk_h);
    cur = out;
  }
  if (need_v) {
    resize_coeffs_kernel<<<(oh + 127) / 128, 128, 0, st>>>(h, oh, kv_, bounds_v, kk_v);
    resize_pass_kernel<<<rs_grid((long long)n * oh * ow * c), 256, 0, st>>>(cur, dst, (long long)n, h, oh, ow * c, kv_, bounds_v, kk_v);
  }
  GDN_LAUNCH_CHECK("resize kernels");
  return GDN_OK;
}
