// Device-side input pipeline: the per-sample CPU transforms of the reference's data loader
//   RandomHorizontalFlip -> RandomScaleCrop -> ArrayToTensor (/255) -> Normalize(mean 0.5, std 0.5)
// (/root/reference/src/transform_list.py:84-113,161-203, composed in src/GDN_main.py:41-66) as ONE kernel over a
// uint8 HWC batch: a B200 consumes > 500 images/s per GPU, which the Python / numpy / imresize loader
// (workers = 0 by default) cannot feed; a uint8 batch is also 4x less host->device traffic than fp32 tensors.
//
//   dst[n][c][y][x] = ((float)v / 255 - 0.5) / 0.5,   v = src'[n][y][x][c]
// where src' is src after the optional flip (x -> W-1-x) and the optional zoom-and-crop: the image is resized to
// (round-down of H*sy, W*sx) with pixel-centre-aligned bilinear interpolation, rounded to uint8 like imresize does,
// and the H x W window at (off_y, off_x) is kept.  HBM-bound: C bytes in, 4C bytes out per pixel.
#include "common.cuh"

namespace gdn {

struct PrepK {
  const uint8_t* src;   // [N][H][W][C]
  float* dst;           // [N][C][H][W]
  int N, H, W, C;
  const int32_t* flip;  // [N] or NULL
  const float* crop;    // [N][4] = (scaled_h, scaled_w, off_y, off_x) as floats holding integers, or NULL
};

__device__ __forceinline__ float prep_norm(float v) {
  // ArrayToTensor: float / 255 ; Normalize: sub_(0.5).div_(0.5) -- same operation order, IEEE division
  return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 0.5f);
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
      const float sh = k.crop[4 * n + 0], sw = k.crop[4 * n + 1];
      const float Y = (float)y + k.crop[4 * n + 2], X = (float)x + k.crop[4 * n + 3];
      // pixel-centre alignment: source = (dst + 0.5) * in / out - 0.5, clamped to the image
      float fy = (Y + 0.5f) * ((float)k.H / sh) - 0.5f, fx = (X + 0.5f) * ((float)k.W / sw) - 0.5f;
      fy = fminf(fmaxf(fy, 0.f), (float)(k.H - 1));
      fx = fminf(fmaxf(fx, 0.f), (float)(k.W - 1));
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = min(y0 + 1, k.H - 1), x1 = min(x0 + 1, k.W - 1);
      const float wy = fy - (float)y0, wx = fx - (float)x0;
      // the flip is applied BEFORE the zoom in the reference: sample the mirrored image
      const int xa = flip ? k.W - 1 - x0 : x0, xb = flip ? k.W - 1 - x1 : x1;
      for (int c = 0; c < k.C; c++) {
        const float p00 = img[((size_t)y0 * k.W + xa) * k.C + c], p01 = img[((size_t)y0 * k.W + xb) * k.C + c];
        const float p10 = img[((size_t)y1 * k.W + xa) * k.C + c], p11 = img[((size_t)y1 * k.W + xb) * k.C + c];
        const float top = p00 + wx * (p01 - p00), bot = p10 + wx * (p11 - p10);
        v[c] = fminf(fmaxf(rintf(top + wy * (bot - top)), 0.f), 255.f);   // imresize returns uint8
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
                                                                        gdn_stream stream) {
  if (!src || !dst || n < 1 || h < 1 || w < 1 || c < 1 || c > 4)
    return fail(GDN_INVALID_DESC, "gdn_preprocess_u8: bad arguments (n=%d h=%d w=%d c=%d)", n, h, w, c);
  PrepK k{src, dst, n, h, w, c, flip, crop};
  const long long total = (long long)n * h * w;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  preprocess_u8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(k);
  GDN_LAUNCH_CHECK("preprocess_u8_kernel");
  return GDN_OK;
}
