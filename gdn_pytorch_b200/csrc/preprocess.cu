// Device-side input pipeline: the per-sample CPU transforms of the reference's data loader
//   RandomHorizontalFlip -> RandomScaleCrop -> ArrayToTensor (/255) -> Normalize(mean 0.5, std 0.5)
// (/root/reference/src/transform_list.py:84-113,161-203, composed in src/GDN_main.py:41-66) over a uint8 HWC batch:
// a B200 consumes > 500 images/s per GPU, which the Python / numpy / imresize loader (workers = 0 by default) cannot
// feed; a uint8 batch is also 4x less host->device traffic than fp32 tensors.
//
//   dst[n][c][y][x] = ((float)v / 255 - 0.5) / 0.5,   v = src'[n][y][x][c]
// where src' is src after the optional flip (x -> W-1-x) and the optional zoom-and-crop.  The zoom is the reference's
// scipy.misc.imresize(im, (scaled_h, scaled_w)) applied to the FLOAT32 image the loader produces
// (datasets_list.py load_as_float: imread(..).astype(np.float32)), i.e.
//   bytescale  : per-image min-max stretch of the whole array to 0..255 (float32 arithmetic, +0.5, truncate)
//   PIL resize : BILINEAR on the 8-bit image -- horizontal pass, 8-bit intermediate, vertical pass; triangle filter,
//                coefficients normalised in double and quantised to 22-bit fixed point (same arithmetic as
//                csrc/resize.cu, which the demo path pins bit-for-bit on Pillow)
// followed by the crop of the h x w window at (off_y, off_x).  Zooming only enlarges (factors in [1, 1.15]), so the
// filter support is 1 source pixel: at most 3 taps per axis, computed per thread.  HBM-bound: C bytes in (read ~4x
// through L1/L2), 4C bytes out per pixel.
#include "common.cuh"

namespace gdn {

constexpr int kPrepPrecisionBits = 32 - 8 - 2;

struct PrepK {
  const uint8_t* src;   // [N][H][W][C]
  float* dst;           // [N][C][H][W]
  int N, H, W, C;
  const int32_t* flip;  // [N] or NULL
  const float* crop;    // [N][4] = (scaled_h, scaled_w, off_y, off_x) as floats holding integers, or NULL
  const unsigned int* mm;  // [N][2] per-image (min, max) of the uint8 values (crop only)
};

__device__ __forceinline__ float prep_norm(float v) {
  // ArrayToTensor: float / 255 ; Normalize: sub_(0.5).div_(0.5) -- same operation order, IEEE division
  return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 0.5f);
}

// per-image min / max of the raw bytes (bytescale's cmin / cmax; a flip does not change them)
__global__ void prep_minmax_init_kernel(unsigned int* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = 255u;
    mm[2 * i + 1] = 0u;
  }
}

__global__ void __launch_bounds__(256) prep_minmax_kernel(const uint8_t* __restrict__ src, long long per_img, unsigned int* mm) {
  const int n = blockIdx.y;
  const uint8_t* p = src + (size_t)n * per_img;
  unsigned int lo = 255u, hi = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int v = p[i];
    lo = min(lo, v);
    hi = max(hi, v);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 2 * n, lo);
    atomicMax(mm + 2 * n + 1, hi);
  }
}

// Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for ONE output index of an enlarging
// BILINEAR resize (scale = in / out <= 1 -> support 1, ksize 3): first source index, tap count, fixed-point weights.
__device__ __forceinline__ void pil_coeffs3(int in_size, int out_size, int xx, int& xmin, int& cnt, int (&k)[3]) {
  const double scale = __ddiv_rn((double)((float)in_size - 0.0f), (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;       // == 1 here; kept for the formula's sake
  const double support = filterscale;
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
  xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  cnt = xmax - xmin;
  if (cnt > 3) cnt = 3;
  double w[3], ww = 0.0;
#pragma unroll
  for (int x = 0; x < 3; x++) {
    w[x] = 0.0;
    if (x < cnt) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      ww = __dadd_rn(ww, w[x]);
    }
  }
#pragma unroll
  for (int x = 0; x < 3; x++) {
    double v = w[x];
    if (x < cnt && ww != 0.0) v = __ddiv_rn(v, ww);
    const double q = __dmul_rn(v, (double)(1 << kPrepPrecisionBits));
    k[x] = v < 0.0 ? (int)__dadd_rn(-0.5, q) : (int)__dadd_rn(0.5, q);
  }
}

__device__ __forceinline__ int pil_clip8(int acc) {
  acc >>= kPrepPrecisionBits;
  return acc < 0 ? 0 : (acc > 255 ? 255 : acc);
}

__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PrepK k) {
  const int plane = k.H * k.W;
  const long long total = (long long)k.N * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / plane);
    const int rem = (int)(i - (long long)n * plane);
    const int y = rem / k.W, x = rem - y * k.W;
    const uint8_t* img = k.src + (size_t)n * plane * k.C;
    const bool flip = k.flip && k.flip[n];
    float v[4];
    if (k.crop) {
      const int sh = (int)k.crop[4 * n + 0], sw = (int)k.crop[4 * n + 1];
      const int Y = y + (int)k.crop[4 * n + 2], X = x + (int)k.crop[4 * n + 3];
      // bytescale constants (float32 array * float64 scalar rounded to float32: NumPy's legacy promotion)
      const float cmin = (float)k.mm[2 * n], cmax = (float)k.mm[2 * n + 1];
      float cscale = __fsub_rn(cmax, cmin);
      if (cscale == 0.f) cscale = 1.f;
      const float scale = (float)__ddiv_rn(255.0, (double)cscale);
      int x0 = X, nx = 1, kx[3] = {1 << kPrepPrecisionBits, 0, 0};
      int y0 = Y, ny = 1, ky[3] = {1 << kPrepPrecisionBits, 0, 0};
      const bool pass_h = sw != k.W, pass_v = sh != k.H;     // PIL skips a pass whose size does not change
      if (pass_h) pil_coeffs3(k.W, sw, X, x0, nx, kx);
      if (pass_v) pil_coeffs3(k.H, sh, Y, y0, ny, ky);
      for (int c = 0; c < k.C; c++) {
        int accv = 1 << (kPrepPrecisionBits - 1);
        int single = 0;
        for (int j = 0; j < ny; j++) {
          const uint8_t* row = img + (size_t)(y0 + j) * k.W * k.C;
          int acch = 1 << (kPrepPrecisionBits - 1);
          int t = 0;
          for (int a = 0; a < nx; a++) {
            const int xs = flip ? k.W - 1 - (x0 + a) : (x0 + a);       // the flip precedes the zoom: mirrored image
            float b = __fmul_rn(__fsub_rn((float)row[(size_t)xs * k.C + c], cmin), scale);
            b = fminf(fmaxf(b, 0.f), 255.f);
            const int byte = (int)__fadd_rn(b, 0.5f);
            acch += byte * kx[a];
            t = byte;
          }
          if (pass_h) t = pil_clip8(acch);       // 8-bit intermediate image of the horizontal pass
          accv += t * ky[j];
          single = t;
        }
        v[c] = (float)(pass_v ? pil_clip8(accv) : single);
      }
    } else {
      const int xs = flip ? k.W - 1 - x : x;
      for (int c = 0; c < k.C; c++) v[c] = (float)img[((size_t)y * k.W + xs) * k.C + c];
    }
    for (int c = 0; c < k.C; c++) k.dst[((size_t)n * k.C + c) * plane + rem] = prep_norm(v[c]);
  }
}

}  // namespace gdn

using namespace gdn;

extern "C" __attribute__((visibility("default"))) int gdn_preprocess_u8(const uint8_t* src, float* dst, int n, int h, int w,
                                                                        int c, const int32_t* flip, const float* crop,
                                                                        void* scratch, gdn_stream stream) {
  if (!src || !dst || n < 1 || h < 1 || w < 1 || c < 1 || c > 4)
    return fail(GDN_INVALID_DESC, "gdn_preprocess_u8: bad arguments (n=%d h=%d w=%d c=%d)", n, h, w, c);
  if (crop && !scratch) return fail(GDN_WORKSPACE_TOO_SMALL, "gdn_preprocess_u8: the zoom needs 8*n bytes of device scratch");
  cudaStream_t st = (cudaStream_t)stream;
  PrepK k{src, dst, n, h, w, c, flip, crop, reinterpret_cast<const unsigned int*>(scratch)};
  if (crop) {
    unsigned int* mm = reinterpret_cast<unsigned int*>(scratch);
    prep_minmax_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(mm, n);
    const long long per_img = (long long)h * w * c;
    int bx = (int)((per_img + 256 * 16 - 1) / (256 * 16));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    prep_minmax_kernel<<<dim3(bx, n), 256, 0, st>>>(src, per_img, mm);
  }
  const long long total = (long long)n * h * w;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  preprocess_u8_kernel<<<(int)blocks, 256, 0, st>>>(k);
  GDN_LAUNCH_CHECK("preprocess_u8_kernel");
  return GDN_OK;
}
