// 64 -> 1 network heads (AE_model_unet.py:300,411,521,628: Conv2d / ConvTranspose2d(64, 1, k9, pad 4) + tanh).
//
// As an implicit GEMM the head is the worst shape there is: N = 1 output channel (a 16-wide MMA tile), so every one of
// the 81 taps re-reads its 128-pixel x 64-channel A window from shared memory for 1/16 of a tensor-core slot
// (profiles/r02f_profile_ops.log: 0.53 ms at 21 TFLOP/s).  Here the taps become the N dimension instead:
//
//     Z[p][t] = sum_c X[p][c] * W[t][c]          one 1x1 convolution, N = 81 (padded to 128), K = 64   (gdn_conv2d)
//     out[p]  = act( sum_t Z[p + off(t)][t] )     shifted sum over the 9x9 neighbourhood                (this file)
//
// With ZERO padding Z of an out-of-image pixel is 0, so the sum just skips the taps that fall outside.
// One CTA = an 8 x 32 tile of outputs: the 16 x 40 neighbourhood of Z rows (only the first 88 of the 128 columns: 11
// 16-byte vectors per pixel) is staged in shared memory with coalesced vector loads, then every thread sums its 81 taps.
// HBM-bound: 2 * zc bytes per pixel read once from HBM (tile overlaps are L2 hits) + 4 bytes written.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace gdn {

constexpr int kHgTH = 8, kHgTW = 32, kHgThreads = kHgTH * kHgTW;

template <int K>
__global__ void __launch_bounds__(kHgThreads) head_gather_kernel(const uint16_t* __restrict__ z, const int z_half, const int zc,
                                                                 const int N, const int H, const int W, const int pad,
                                                                 const int tanh_out, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  constexpr int KV = (K * K + 7) / 8;          // 16-byte vectors of Z kept per pixel
  constexpr int HH = kHgTH + K - 1, HW = kHgTW + K - 1;
  extern __shared__ uint4 s_z[];               // [HH][HW][KV]
  const int tiles_x = (W + kHgTW - 1) / kHgTW, tiles_y = (H + kHgTH - 1) / kHgTH;
  int t = blockIdx.x;
  const int bx = t % tiles_x;
  t /= tiles_x;
  const int by = t % tiles_y;
  const int n = t / tiles_y;
  const int y0 = by * kHgTH, x0 = bx * kHgTW;
  // asynchronous 16-byte copies (LDGSTS): all ~28 copies of a thread are in flight at once -- with plain loads the loop
  // runs one memory latency per iteration (measured 0.21 ms for the 272 MB of the bench shape, profiles/r02l_profile_ops.log);
  // out-of-image pixels copy 0 source bytes = zero fill
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_z);
  for (int i = threadIdx.x; i < HH * HW * KV; i += kHgThreads) {
    const int v = i % KV, px = i / KV;
    const int hy = px / HW, hx = px - hy * HW;
    const int gy = y0 - pad + hy, gx = x0 - pad + hx;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const uint16_t* src = in ? z + (((size_t)n * H + gy) * W + gx) * zc + v * 8 : z;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_base + (uint32_t)i * 16u), "l"(src), "r"(in ? 16 : 0)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int ty = threadIdx.x / kHgTW, tx = threadIdx.x % kHgTW;
  const uint16_t* sz = reinterpret_cast<const uint16_t*>(s_z);
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < K; r++) {
    const uint16_t* row = sz + (size_t)((ty + r) * HW + tx) * (KV * 8) + r * K;
#pragma unroll
    for (int s = 0; s < K; s++) {
      const uint16_t u = row[s * (KV * 8) + s];
      acc += z_half ? __half2float(__ushort_as_half(u)) : __bfloat162float(__ushort_as_bfloat16(u));
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < W) out[((size_t)n * H + y) * W + x] = tanh_out ? tanhf(acc) : acc;
}

}  // namespace gdn

using namespace gdn;

#define GDN_API __attribute__((visibility("default")))
extern "C" GDN_API int gdn_head_gather(const void* z, int z_is_half, int zc, int n, int h, int w, int k, int pad,
                                       int tanh_out, float* out, gdn_stream stream) {
  if (!z || !out || n < 1 || h < 1 || w < 1) return fail(GDN_INVALID_DESC, "gdn_head_gather: null pointer / empty extent");
  if (k != 9) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_head_gather: kernel size %d (the heads of the path are 9 x 9)", k);
  if (zc % 8 || zc < (k * k + 7) / 8 * 8)
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_head_gather: %d columns per pixel (need a multiple of 8, >= %d)", zc, (k * k + 7) / 8 * 8);
  if (pad < 0 || pad > k - 1) return fail(GDN_INVALID_DESC, "gdn_head_gather: pad %d", pad);
  const long long tiles = (long long)n * ((h + kHgTH - 1) / kHgTH) * ((w + kHgTW - 1) / kHgTW);
  if (tiles > 0x7fffffffll) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_head_gather: too many tiles");
  constexpr int KV = (9 * 9 + 7) / 8;
  const size_t smem = (size_t)(kHgTH + 8) * (kHgTW + 8) * KV * sizeof(uint4);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    GDN_CUDA_CHECK(cudaFuncSetAttribute(head_gather_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  GDN_CUDA_CHECK(launch_pdl(head_gather_kernel<9>, dim3((unsigned)tiles), dim3(kHgThreads), smem, (cudaStream_t)stream, 1,
                            reinterpret_cast<const uint16_t*>(z), z_is_half, zc, n, h, w, pad, tanh_out, out));
  GDN_LAUNCH_CHECK("head_gather_kernel");
  return GDN_OK;
}
