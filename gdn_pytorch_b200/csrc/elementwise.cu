// HBM-bound helper kernels around the tensor-core convolutions (sm_100a):
//   im2col for thin-channel layers, BatchNorm finalize / apply (+ReLU, residual, reflection halo, x2 bilinear
//   upsample, zero dilation), BatchNorm backward (reduce + apply), gradient fold (adjoint of reflection padding /
//   upsampling / dilation / channel slicing), weight pack / gradient unpack.
// All tensors are NHWC; 8 bf16 (16 B) or 4 fp32 (16 B) per thread access, grid-stride loops.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "fold_rows.cuh"
#include "pack_tile.cuh"

namespace gdn {

constexpr int kEwThreads = 256;

static inline int ew_grid(long long work) {
  long long b = (work + kEwThreads - 1) / kEwThreads;
  const long long cap = (long long)device_sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static inline int ew_lg2(int v) {  // log2 of a power of two, -1 otherwise
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

static inline int ew_row_mult() {   // CTAs per SM of the row kernels (GDN_EW_ROW_MULT, experiment knob; default 8)
  static int m = 0;
  if (!m) {
    const char* e = getenv("GDN_EW_ROW_MULT");
    m = e ? atoi(e) : 8;
    if (m < 1) m = 8;
  }
  return m;
}

static inline int ew_row_grid(long long rows) {
  const long long cap = (long long)device_sm_count() * ew_row_mult();
  return (int)(rows < cap ? (rows < 1 ? 1 : rows) : cap);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // nn.ReflectionPad2d index map (no edge repeat); valid for |overshoot| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct bf16x8 {
  uint4 u;
};
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack8h(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ------------------------------------------------------------------------------------------------ im2col
// src: fp32 planar [N][C][H][W] (C <= 4).  dst: bf16 [N*H*W][kpad], k = (r*kw + s)*C + c, zero beyond kh*kw*C.
__global__ void im2col_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int C, int H,
                              int W, int kh, int kw, int pad, int reflect, int kpad) {
  pdl_trigger();
  pdl_wait();
  const int kreal = kh * kw * C;
  const int groups = kpad / 8;
  const long long total = (long long)N * H * W * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long pix = i / groups;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int k = g * 8 + j;
      float v = 0.f;
      if (k < kreal) {
        const int c = k % C, t = k / C;
        const int s = t % kw, r = t / kw;
        int yy = y + r - pad, xx = x + s - pad;
        bool ok = true;
        if (reflect) {
          yy = reflect_idx(yy, H);
          xx = reflect_idx(xx, W);
        } else {
          ok = (yy >= 0 && yy < H && xx >= 0 && xx < W);
        }
        if (ok) v = __ldg(src + (((long long)n * C + c) * H + yy) * W + xx);
      }
      f[j] = v;
    }
    *reinterpret_cast<uint4*>(dst + pix * kpad + g * 8) = pack8(f);
  }
}



// shared-memory path (the one the networks use): a CTA iteration produces a BAND of up to kIm2colBand output rows; the
// band's kh - 1 + band input rows are staged once in shared memory, pixel-interleaved ([row][x + pad][c], borders already
// reflected / zeroed), so that the kw*C columns of one tap row are a contiguous run: column k of pixel x is
// s_rows[off(k) + x*C] with off(k) from a small per-lane table.  A warp writes one pixel's kpad columns, each lane two
// adjacent columns per store (128 contiguous bytes per warp instruction).  The first version staged kh rows per OUTPUT
// row (9x re-staging with two integer divisions per element) and stored 2 bytes per lane: 1.7 TB/s on the column matrix
// (profiles/r02e_ncu_elem.summary.txt).
constexpr int kIm2colBand = 8;

__global__ void __launch_bounds__(kEwThreads) im2col_smem_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                 int N, int C, int H, int W, int kh, int kw, int pad,
                                                                 int reflect, int kpad, int band_rows) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_rows[];   // [kh - 1 + band_rows][(W + 2*pad) * C]
  const int Wp = W + 2 * pad;
  const int rowstride = Wp * C;
  const int kreal = kh * kw * C, kwc = kw * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int off[4][2];                      // kpad <= 256: at most 4 column pairs per lane
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int k = j * 64 + 2 * lane + e;
      off[j][e] = (k < kreal && k < kpad) ? (k / kwc) * rowstride + (k % kwc) : -1;
    }
  const int nj = kpad >> 6;
  const int bands_img = (H + band_rows - 1) / band_rows;
  const int bands = N * bands_img, plane = H * W;
  for (int band = blockIdx.x; band < bands; band += gridDim.x) {
    const int n = band / bands_img, y0 = (band - n * bands_img) * band_rows;
    const int nr = min(band_rows, H - y0);
    const float* img = src + (size_t)n * C * plane;
    __syncthreads();
    for (int r = 0; r < nr + kh - 1; r++) {
      int yy = y0 + r - pad;
      const bool yok = reflect || (yy >= 0 && yy < H);
      if (reflect) yy = reflect_idx(yy, H);
      for (int c = 0; c < C; c++) {
        const float* line = img + (size_t)c * plane + (size_t)(yok ? yy : 0) * W;
        for (int xp = threadIdx.x; xp < Wp; xp += kEwThreads) {        // global reads run along x
          int xx = xp - pad;
          bool ok = yok;
          if (reflect) xx = reflect_idx(xx, W);
          else ok = ok && (xx >= 0 && xx < W);
          s_rows[r * rowstride + xp * C + c] = ok ? __ldg(line + xx) : 0.f;
        }
      }
    }
    __syncthreads();
    for (int ry = 0; ry < nr; ry++) {
      __nv_bfloat16* drow = dst + ((size_t)(n * H + y0 + ry) * W) * kpad;
      const float* srow = s_rows + ry * rowstride;
      for (int x = warp; x < W; x += kEwThreads / 32) {
        const float* base = srow + x * C;
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(drow + (size_t)x * kpad) + lane;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (j < nj)
            o[j * 32] = __floats2bfloat162_rn(off[j][0] >= 0 ? base[off[j][0]] : 0.f, off[j][1] >= 0 ? base[off[j][1]] : 0.f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- BN finalize
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float4* __restrict__ coef4, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const BnFinOut o = bn_finalize_channel(sum[c], sq[c], count, gamma[c], beta[c], eps);
  scale[c] = o.scale;
  shift[c] = o.shift;
  mean_out[c] = o.mean;
  rstd_out[c] = o.rstd;
  // what a dgrad epilogue needs per channel for the fused BatchNorm-backward statistics: y = x*scale + shift (ReLU
  // mask), xhat = (x - mean)*rstd -- the same expressions gdn_bn_bwd_reduce evaluates
  if (coef4) coef4[c] = make_float4(o.scale, o.shift, o.mean, o.rstd);
  if (running_mean) {
    running_mean[c] = bn_running_update(running_mean[c], momentum, o.mean);
    running_var[c] = bn_running_update(running_var[c], momentum, o.unbiased_var);
  }
}

// ------------------------------------------------------------------------------------------- act forward
struct ActFwd {
  const __nv_bfloat16* src_bf16;  // [N][H][W][C]  (exactly one of src_bf16 / src_f32)
  const float* src_f32;
  const float* scale;             // per channel or NULL (identity)
  const float* shift;
  const float* resid;             // fp32 [N][H][W][C] or NULL
  int relu;
  int N, H, W, C;
  float* out_f32;                 // [N][H][W][C] or NULL
  __nv_bfloat16* out_bf16;        // [N][OH+2P][OW+2P][C] or NULL, OH = H*up (or 2H when dilate)
  int P, reflect, up, dilate;     // up: 0 none, 1 bilinear x2 align_corners=False, 2 align_corners=True
  int src_half;
};

__device__ __forceinline__ void act_value8(const ActFwd& a, long long pix, int c8, float* v, const float* sc, const float* sh) {
  if (a.src_bf16) {
    const uint4 u = *reinterpret_cast<const uint4*>(a.src_bf16 + pix * a.C + c8);
    if (a.src_half) unpack8h(u, v); else unpack8(u, v);
  } else {
    const float4 p0 = *reinterpret_cast<const float4*>(a.src_f32 + pix * a.C + c8);
    const float4 p1 = *reinterpret_cast<const float4*>(a.src_f32 + pix * a.C + c8 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
  }
  if (a.scale) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = fmaf(v[j], sc[j], sh[j]);
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
  }
  if (a.resid) {
    const float4 p0 = *reinterpret_cast<const float4*>(a.resid + pix * a.C + c8);
    const float4 p1 = *reinterpret_cast<const float4*>(a.resid + pix * a.C + c8 + 4);
    v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
  }
}

// source coordinate + weight of F.interpolate(scale_factor=2, mode='bilinear')
__global__ void act_forward_kernel(const ActFwd a) {
  pdl_trigger();
  pdl_wait();
  const int cg = a.C / 8;
  const long long n_f32 = a.out_f32 ? (long long)a.N * a.H * a.W * cg : 0;
  const int OH = (a.up || a.dilate) ? 2 * a.H : a.H, OW = (a.up || a.dilate) ? 2 * a.W : a.W;
  const int Hp = OH + 2 * a.P, Wp = OW + 2 * a.P;
  const long long n_b16 = a.out_bf16 ? (long long)a.N * Hp * Wp * cg : 0;
  const long long total = n_f32 > n_b16 ? n_f32 : n_b16;
  // blockDim.x is a multiple of cg (host-checked): a thread keeps its 8 channels over the grid-stride loop
  float sc[8], sh[8];
  {
    const int c8 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % cg) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sc[j] = a.scale ? a.scale[c8 + j] : 1.f;
      sh[j] = a.scale ? a.shift[c8 + j] : 0.f;
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_f32) {
      const int c8 = (int)(i % cg) * 8;
      const long long pix = i / cg;
      float v[8];
      act_value8(a, pix, c8, v, sc, sh);
      float* o = a.out_f32 + pix * a.C + c8;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (i < n_b16) {
      const int c8 = (int)(i % cg) * 8;
      long long t = i / cg;
      const int xp = (int)(t % Wp);
      t /= Wp;
      const int yp = (int)(t % Hp);
      const int n = (int)(t / Hp);
      int Y = yp - a.P, X = xp - a.P;
      bool inside = (Y >= 0 && Y < OH && X >= 0 && X < OW);
      if (!inside && !a.reflect) continue;  // non-reflect borders are never read as data (TMA zero fill is used instead)
      Y = reflect_idx(Y, OH);
      X = reflect_idx(X, OW);
      float v[8];
      if (a.dilate) {
        if ((Y & 1) || (X & 1)) {
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = 0.f;
        } else {
          act_value8(a, ((long long)n * a.H + (Y >> 1)) * a.W + (X >> 1), c8, v, sc, sh);
        }
      } else if (a.up) {
        int y0, y1, x0, x1;
        float wy, wx;
        up_coord(Y, a.H, a.up, y0, y1, wy);
        up_coord(X, a.W, a.up, x0, x1, wx);
        float v00[8], v01[8], v10[8], v11[8];
        const long long base = (long long)n * a.H;
        act_value8(a, (base + y0) * a.W + x0, c8, v00, sc, sh);
        act_value8(a, (base + y0) * a.W + x1, c8, v01, sc, sh);
        act_value8(a, (base + y1) * a.W + x0, c8, v10, sc, sh);
        act_value8(a, (base + y1) * a.W + x1, c8, v11, sc, sh);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          // same association as ATen's upsample_bilinear2d: w(1-wy)*(row0 blend) + wy*(row1 blend)
          const float top = (1.f - wx) * v00[j] + wx * v01[j];
          const float bot = (1.f - wx) * v10[j] + wx * v11[j];
          v[j] = (1.f - wy) * top + wy * bot;
        }
      } else {
        act_value8(a, ((long long)n * a.H + Y) * a.W + X, c8, v, sc, sh);
      }
      *reinterpret_cast<uint4*>(a.out_bf16 + (((long long)n * Hp + yp) * Wp + xp) * a.C + c8) = pack8(v);
    }
  }
}


// ---- fast paths (the ones the networks use): one image row per CTA iteration, 32-bit index math (channel-group
// counts are powers of two), two independent items in flight per thread.  The generic kernel above stays as the
// path for everything else (align_corners=True upsampling, odd channel counts).
__device__ __forceinline__ void act_finish8(const ActFwd& a, float* v, const float* sc, const float* sh,
                                            const float4& r0, const float4& r1) {
  if (a.scale) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = fmaf(v[j], sc[j], sh[j]);
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
  }
  if (a.resid) {
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
}

// source-driven: every source pixel is read once and written to out_f32 and to the interior + reflection images of
// out_bf16 (no upsampling / dilation).  U independent items are in flight per thread (4 without a residual stream,
// 2 with one) so that ~64 KB of loads are outstanding per SM.
template <bool RESID, int U>
__global__ void __launch_bounds__(kEwThreads, 3) act_rows_kernel(const ActFwd a, const int lg_cg, const int rows) {
  pdl_trigger();
  pdl_wait();
  const int cgm = (1 << lg_cg) - 1;
  const int items = a.W << lg_cg;
  const int c8 = (threadIdx.x & cgm) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = a.scale ? a.scale[c8 + j] : 1.f;
    sh[j] = a.scale ? a.shift[c8 + j] : 0.f;
  }
  const int P = a.P, Hp = a.H + 2 * P, Wp = a.W + 2 * P;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / a.H, y = row - n * a.H;
    int ys[3], ny = 0;
    ys[ny++] = y + P;
    if (a.reflect && a.out_bf16) {
      if (y >= 1 && y <= P) ys[ny++] = P - y;
      if (y >= a.H - 1 - P && y <= a.H - 2) ys[ny++] = P + 2 * (a.H - 1) - y;
    }
    const size_t row_off = (size_t)row * a.W * a.C;
    for (int it0 = threadIdx.x; it0 < items; it0 += U * kEwThreads) {
      uint4 q[U];        // 16-bit sources: the raw 16 B; fp32 sources: first float4
      float4 q1[U];      // fp32 sources: second float4
      float4 r0[RESID ? U : 1], r1[RESID ? U : 1];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int it = it0 + u * kEwThreads;
        if (it >= items) continue;
        const size_t off = row_off + (size_t)(it >> lg_cg) * a.C + c8;
        if (a.src_bf16) {
          q[u] = __ldg(reinterpret_cast<const uint4*>(a.src_bf16 + off));
        } else {
          q[u] = __ldg(reinterpret_cast<const uint4*>(a.src_f32 + off));
          q1[u] = __ldg(reinterpret_cast<const float4*>(a.src_f32 + off + 4));
        }
        if (RESID) {
          r0[u] = __ldg(reinterpret_cast<const float4*>(a.resid + off));
          r1[u] = __ldg(reinterpret_cast<const float4*>(a.resid + off + 4));
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int it = it0 + u * kEwThreads;
        if (it >= items) continue;
        const int x = it >> lg_cg;
        float v[8];
        if (a.src_bf16) {
          if (a.src_half) unpack8h(q[u], v); else unpack8(q[u], v);
        } else {
          v[0] = __uint_as_float(q[u].x); v[1] = __uint_as_float(q[u].y); v[2] = __uint_as_float(q[u].z); v[3] = __uint_as_float(q[u].w);
          v[4] = q1[u].x; v[5] = q1[u].y; v[6] = q1[u].z; v[7] = q1[u].w;
        }
        act_finish8(a, v, sc, sh, RESID ? r0[u] : z4, RESID ? r1[u] : z4);
        if (a.out_f32) {
          float* o = a.out_f32 + row_off + (size_t)x * a.C + c8;
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (a.out_bf16) {
          const uint4 w = pack8(v);
          int xs[3], nx = 0;
          xs[nx++] = x + P;
          if (a.reflect) {
            if (x >= 1 && x <= P) xs[nx++] = P - x;
            if (x >= a.W - 1 - P && x <= a.W - 2) xs[nx++] = P + 2 * (a.W - 1) - x;
          }
          for (int p = 0; p < ny; p++)
            for (int qq = 0; qq < nx; qq++)
              *reinterpret_cast<uint4*>(a.out_bf16 + (((size_t)n * Hp + ys[p]) * Wp + xs[qq]) * a.C + c8) = w;
        }
      }
    }
  }
}

// output-driven x2 bilinear upsampling (align_corners=False) with an optional reflection border: one padded output
// row per CTA iteration; the row's two source rows and the vertical weight are computed once per row.
__global__ void __launch_bounds__(kEwThreads, 3) act_up_rows_kernel(const ActFwd a, const int lg_cg, const int rows) {
  pdl_trigger();
  pdl_wait();
  const int cgm = (1 << lg_cg) - 1;
  const int OH = 2 * a.H, OW = 2 * a.W;
  const int P = a.P, Hp = OH + 2 * P, Wp = OW + 2 * P;
  const int items = Wp << lg_cg;
  const int c8 = (threadIdx.x & cgm) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = a.scale ? a.scale[c8 + j] : 1.f;
    sh[j] = a.scale ? a.shift[c8 + j] : 0.f;
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Hp, yp = row - n * Hp;
    int Y = yp - P;
    if ((Y < 0 || Y >= OH) && !a.reflect) continue;
    Y = reflect_idx(Y, OH);
    int y0, y1;
    float wy;
    up_coord(Y, a.H, 1, y0, y1, wy);
    const size_t base0 = ((size_t)n * a.H + y0) * a.W, base1 = ((size_t)n * a.H + y1) * a.W;
    for (int it = threadIdx.x; it < items; it += kEwThreads) {
      const int xp = it >> lg_cg;
      int X = xp - P;
      if ((X < 0 || X >= OW) && !a.reflect) continue;
      X = reflect_idx(X, OW);
      int x0, x1;
      float wx;
      up_coord(X, a.W, 1, x0, x1, wx);
      float v00[8], v01[8], v10[8], v11[8];
      const size_t o00 = (base0 + x0) * a.C + c8, o01 = (base0 + x1) * a.C + c8;
      const size_t o10 = (base1 + x0) * a.C + c8, o11 = (base1 + x1) * a.C + c8;
      if (a.src_bf16) {
        const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(a.src_bf16 + o00));
        const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(a.src_bf16 + o01));
        const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(a.src_bf16 + o10));
        const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(a.src_bf16 + o11));
        if (a.src_half) { unpack8h(q00, v00); unpack8h(q01, v01); unpack8h(q10, v10); unpack8h(q11, v11); }
        else { unpack8(q00, v00); unpack8(q01, v01); unpack8(q10, v10); unpack8(q11, v11); }
      } else {
        const size_t os[4] = {o00, o01, o10, o11};
        float* vs[4] = {v00, v01, v10, v11};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(a.src_f32 + os[k]));
          const float4 p1 = __ldg(reinterpret_cast<const float4*>(a.src_f32 + os[k] + 4));
          vs[k][0] = p0.x; vs[k][1] = p0.y; vs[k][2] = p0.z; vs[k][3] = p0.w;
          vs[k][4] = p1.x; vs[k][5] = p1.y; vs[k][6] = p1.z; vs[k][7] = p1.w;
        }
      }
      float4 r00[2] = {z4, z4}, r01[2] = {z4, z4}, r10[2] = {z4, z4}, r11[2] = {z4, z4};
      if (a.resid) {
        r00[0] = __ldg(reinterpret_cast<const float4*>(a.resid + o00)); r00[1] = __ldg(reinterpret_cast<const float4*>(a.resid + o00 + 4));
        r01[0] = __ldg(reinterpret_cast<const float4*>(a.resid + o01)); r01[1] = __ldg(reinterpret_cast<const float4*>(a.resid + o01 + 4));
        r10[0] = __ldg(reinterpret_cast<const float4*>(a.resid + o10)); r10[1] = __ldg(reinterpret_cast<const float4*>(a.resid + o10 + 4));
        r11[0] = __ldg(reinterpret_cast<const float4*>(a.resid + o11)); r11[1] = __ldg(reinterpret_cast<const float4*>(a.resid + o11 + 4));
      }
      act_finish8(a, v00, sc, sh, r00[0], r00[1]);
      act_finish8(a, v01, sc, sh, r01[0], r01[1]);
      act_finish8(a, v10, sc, sh, r10[0], r10[1]);
      act_finish8(a, v11, sc, sh, r11[0], r11[1]);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float top = (1.f - wx) * v00[j] + wx * v01[j];
        const float bot = (1.f - wx) * v10[j] + wx * v11[j];
        v[j] = (1.f - wy) * top + wy * bot;
      }
      *reinterpret_cast<uint4*>(a.out_bf16 + ((size_t)row * Wp + xp) * a.C + c8) = pack8(v);
    }
  }
}

// Pure x2 bilinear upsampling (align_corners=False) of a plain bf16 tensor into a (reflection-)bordered bf16 buffer --
// the operand of the decoder's up-convolutions (F.interpolate + nn.ReflectionPad2d, AE_model_unet.py:336-355, 66).
// SOURCE-block driven: a thread owns the 2 x 2 source block (yb, yb+1) x (xb, xb+1) of 8 channels and produces the
// 2 x 2 outputs (2yb+1, 2yb+2) x (2xb+1, 2xb+2) it alone determines (fixed weights 0.75 / 0.25; yb = -1 and H-1 give the
// clamped edge rows), plus their reflection images: four 16-byte loads for four 16-byte stores.  The output-driven
// act_up_rows_kernel re-reads four sources and redoes the coordinate arithmetic per OUTPUT (1.4 TB/s at the bench
// shapes, profiles/r02e_ncu_elem.summary.txt); it stays for the fused BatchNorm / residual / fp32-source forms.
__global__ void __launch_bounds__(kEwThreads, 3) up2x_blocks_kernel(const ActFwd a, const int lg_cg, const int rows) {
  pdl_trigger();
  pdl_wait();
  const int cgm = (1 << lg_cg) - 1;
  const int OH = 2 * a.H, OW = 2 * a.W;
  const int P = a.P, Hp = OH + 2 * P, Wp = OW + 2 * P;
  const int items = (a.W + 1) << lg_cg;
  const int c8 = (threadIdx.x & cgm) * 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {      // row = (image, source-block row yb + 1)
    const int n = row / (a.H + 1), yb = row - n * (a.H + 1) - 1;
    const int y0 = max(yb, 0), y1 = min(yb + 1, a.H - 1);
    const size_t base0 = ((size_t)n * a.H + y0) * a.W, base1 = ((size_t)n * a.H + y1) * a.W;
    // the two output rows of this block row and the buffer rows holding them (interior + reflection images)
    int ys[2][3], ny[2] = {0, 0};
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int Y = 2 * yb + 1 + r;
      if (Y < 0 || Y >= OH) continue;
      ys[r][ny[r]++] = Y + P;
      if (a.reflect) {
        if (Y >= 1 && Y <= P) ys[r][ny[r]++] = P - Y;
        if (Y >= OH - 1 - P && Y <= OH - 2) ys[r][ny[r]++] = P + 2 * (OH - 1) - Y;
      }
    }
    for (int it = threadIdx.x; it < items; it += kEwThreads) {
      const int xb = (it >> lg_cg) - 1;
      const int x0 = max(xb, 0), x1 = min(xb + 1, a.W - 1);
      float v00[8], v01[8], v10[8], v11[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.src_bf16 + (base0 + x0) * a.C + c8)), v00);
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.src_bf16 + (base0 + x1) * a.C + c8)), v01);
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.src_bf16 + (base1 + x0) * a.C + c8)), v10);
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.src_bf16 + (base1 + x1) * a.C + c8)), v11);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int X = 2 * xb + 1 + c;
        if (X < 0 || X >= OW) continue;
        // same expression as the output-driven kernel: weight of the second source = 0.25 for odd outputs, 0.75 for even ones
        const float wx = (x0 == x1) ? 0.f : (c ? 0.75f : 0.25f);     // clamped edge columns copy their source exactly
        int xs[3], nx = 0;
        xs[nx++] = X + P;
        if (a.reflect) {
          if (X >= 1 && X <= P) xs[nx++] = P - X;
          if (X >= OW - 1 - P && X <= OW - 2) xs[nx++] = P + 2 * (OW - 1) - X;
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
          if (ny[r] == 0) continue;
          const float wy = (y0 == y1) ? 0.f : (r ? 0.75f : 0.25f);
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float top = (1.f - wx) * v00[j] + wx * v01[j];
            const float bot = (1.f - wx) * v10[j] + wx * v11[j];
            v[j] = (1.f - wy) * top + wy * bot;
          }
          const uint4 o = pack8(v);
          for (int p = 0; p < ny[r]; p++)
            for (int q = 0; q < nx; q++)
              *reinterpret_cast<uint4*>(a.out_bf16 + (((size_t)n * Hp + ys[r][p]) * Wp + xs[q]) * a.C + c8) = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ BN backward
struct BnBwd {
  const float* dact;              // fp32 [N][H][W][C] gradient w.r.t. the activation output
  const __nv_bfloat16* raw;       // bf16 [N][H][W][C] stored conv output (pre-BN)
  const float* scale;             // gamma * rstd
  const float* shift;
  const float* mean;
  const float* rstd;
  int relu;
  long long npix;
  int C;
  double* sum_g;                  // [C]
  double* sum_gx;                 // [C]
  // apply
  __nv_bfloat16* dy;              // bf16 [N][H][W][C] (or dilated [N][2H][2W][C]) gradient w.r.t. the conv output
  int H, W, dilate;
  float* dgamma;                  // += sum_gx   (may be NULL)
  float* dbeta;                   // += sum_g
  int raw_half;
  int dact_bf16;                  // dact holds bf16 (written by a convolution epilogue with the ReLU mask already applied)
  int det;                        // GDN_DETERMINISTIC: thread replicas add their partials in a fixed order
};

// 8 consecutive gradient values of element offset `off` (fp32 stream, or the bf16 buffer a dgrad epilogue wrote)
__device__ __forceinline__ void bn_grad8(const BnBwd& b, size_t off, float* d) {
  if (b.dact_bf16) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(b.dact) + off)), d);
  } else {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(b.dact + off));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(b.dact + off + 4));
    d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
  }
}


// Per-channel reductions: threads are (pixel lane, 8-channel group) with the channel group fastest, so a warp
// reads contiguous NHWC bytes.  Partials go block-level through shared-memory atomics, then one fp64 atomic per
// channel per block.
__global__ void bn_bwd_reduce_kernel(const BnBwd b) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_red[];  // [2][C]
  const int cgs = b.C / 8;
  const int lanes = blockDim.x / cgs;
  const int g = threadIdx.x % cgs, pl = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < 2 * b.C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float sg[8], sx[8];
#pragma unroll
  for (int j = 0; j < 8; j++) sg[j] = sx[j] = 0.f;
  if (pl < lanes) {
    const int c8 = g * 8;
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      sc[j] = b.scale[c8 + j];
      sh[j] = b.shift[c8 + j];
      mu[j] = b.mean[c8 + j];
      rs[j] = b.rstd[c8 + j];
    }
    for (long long pix = (long long)blockIdx.x * lanes + pl; pix < b.npix; pix += (long long)gridDim.x * lanes) {
      float x[8];
      if (b.raw_half) unpack8h(*reinterpret_cast<const uint4*>(b.raw + pix * b.C + c8), x);
      else unpack8(*reinterpret_cast<const uint4*>(b.raw + pix * b.C + c8), x);
      float d[8];
      bn_grad8(b, (size_t)(pix * b.C + c8), d);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float gq = d[j];
        if (b.relu && fmaf(x[j], sc[j], sh[j]) <= 0.f) gq = 0.f;
        sg[j] += gq;
        sx[j] += gq * (x[j] - mu[j]) * rs[j];
      }
    }
    if (!b.det) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        atomicAdd(&s_red[c8 + j], sg[j]);
        atomicAdd(&s_red[b.C + c8 + j], sx[j]);
      }
    }
  }
  if (b.det) {
    // fixed order: the pixel lanes that hold the same 8 channels take turns
    for (int r = 0; r < lanes; r++) {
      if (pl == r) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s_red[g * 8 + j] += sg[j];
          s_red[b.C + g * 8 + j] += sx[j];
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < b.C; i += blockDim.x) {
    atomicAdd(b.sum_g + i, (double)s_red[i]);
    atomicAdd(b.sum_gx + i, (double)s_red[b.C + i]);
  }
}

__global__ void bn_bwd_apply_kernel(const BnBwd b) {
  pdl_trigger();
  pdl_wait();
  const int cg = b.C / 8;
  const int OH = b.dilate ? 2 * b.H : b.H, OW = b.dilate ? 2 * b.W : b.W;
  const long long N = b.npix / ((long long)b.H * b.W);
  const long long total = N * OH * OW * cg;
  const double inv_n = 1.0 / (double)b.npix;
  if (b.dgamma && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
      b.dgamma[c] += (float)(b.sum_gx[c]);
      b.dbeta[c] += (float)(b.sum_g[c]);
    }
  }
  // blockDim.x (256) is a multiple of cg, so every thread keeps the same 8 channels over the grid-stride loop
  const int c8 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % cg) * 8;
  float sc[8], sh[8], mu[8], rs[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = b.scale[c8 + j];
    sh[j] = b.shift[c8 + j];
    mu[j] = b.mean[c8 + j];
    rs[j] = b.rstd[c8 + j];
    m1[j] = (float)(b.sum_g[c8 + j] * inv_n);
    m2[j] = (float)(b.sum_gx[c8 + j] * inv_n);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i / cg;
    const int X = (int)(t % OW);
    t /= OW;
    const int Y = (int)(t % OH);
    const long long n = t / OH;
    float o[8];
    if (b.dilate && ((Y & 1) || (X & 1))) {
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = 0.f;
    } else {
      const int y = b.dilate ? (Y >> 1) : Y, x = b.dilate ? (X >> 1) : X;
      const long long pix = (n * b.H + y) * b.W + x;
      float xr[8];
      if (b.raw_half) unpack8h(*reinterpret_cast<const uint4*>(b.raw + pix * b.C + c8), xr);
      else unpack8(*reinterpret_cast<const uint4*>(b.raw + pix * b.C + c8), xr);
      float d[8];
      bn_grad8(b, (size_t)(pix * b.C + c8), d);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float gq = d[j];
        if (b.relu && fmaf(xr[j], sc[j], sh[j]) <= 0.f) gq = 0.f;
        const float xh = (xr[j] - mu[j]) * rs[j];
        o[j] = sc[j] * (gq - m1[j] - xh * m2[j]);
      }
    }
    *reinterpret_cast<uint4*>(b.dy + i * 8) = pack8(o);
  }
}


// ---- fast paths: contiguous pixel ranges per CTA, 32-bit index math, two items in flight per thread
__device__ __forceinline__ void bn_load8(const BnBwd& b, size_t off, float* x, float* d) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(b.raw + off));
  bn_grad8(b, off, d);
  if (b.raw_half) unpack8h(q, x); else unpack8(q, x);
}

// per-channel sums of g and g*(x - mean) (scaled by rstd at the end); items = npix * cg, block-contiguous chunks
__global__ void __launch_bounds__(kEwThreads, 3) bn_bwd_reduce_fast_kernel(const BnBwd b, const int lg_cg, const int chunk) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_red[];  // [2][C]
  const int cgm = (1 << lg_cg) - 1;
  const int c8 = (threadIdx.x & cgm) * 8;
  for (int i = threadIdx.x; i < 2 * b.C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  // sx accumulates g*x; the mean is taken out once per thread at the end (sum g*(x - mu) = sum g*x - mu * sum g)
  float sc[8], sh[8], sg[8], sx[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = b.scale[c8 + j];
    sh[j] = b.shift[c8 + j];
    sg[j] = sx[j] = 0.f;
  }
  const long long total = b.npix << lg_cg;
  const long long i0 = (long long)blockIdx.x * chunk;
  const long long i1 = (i0 + chunk < total) ? i0 + chunk : total;
  for (long long it0 = i0 + threadIdx.x; it0 < i1; it0 += 2 * kEwThreads) {
    float x[2][8], d[2][8];
    const bool two = it0 + kEwThreads < i1;
    bn_load8(b, (size_t)it0 * 8, x[0], d[0]);
    if (two) bn_load8(b, (size_t)(it0 + kEwThreads) * 8, x[1], d[1]);
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (u == 1 && !two) break;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float gq = d[u][j];
        if (b.relu && fmaf(x[u][j], sc[j], sh[j]) <= 0.f) gq = 0.f;
        sg[j] += gq;
        sx[j] = fmaf(gq, x[u][j], sx[j]);
      }
    }
  }
  if (!b.det) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      atomicAdd(&s_red[c8 + j], sg[j]);
      atomicAdd(&s_red[b.C + c8 + j], (sx[j] - b.mean[c8 + j] * sg[j]) * b.rstd[c8 + j]);
    }
  } else {
    // fixed order: the kEwThreads >> lg_cg threads that hold the same 8 channels take turns
    for (int r = 0; r < (kEwThreads >> lg_cg); r++) {
      if ((int)(threadIdx.x >> lg_cg) == r) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          s_red[c8 + j] += sg[j];
          s_red[b.C + c8 + j] += (sx[j] - b.mean[c8 + j] * sg[j]) * b.rstd[c8 + j];
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < b.C; i += blockDim.x) {
    atomicAdd(b.sum_g + i, (double)s_red[i]);
    atomicAdd(b.sum_gx + i, (double)s_red[b.C + i]);
  }
}

// dy = scale * (g - mean(g) - xhat * mean(g xhat)), no dilation: flat item index == flat element index / 8.
// U items in flight per thread: 2 with the fp32 gradient stream (96 B of loads per thread), 4 with the bf16 one (its
// 32 B per item leave the kernel latency-bound at 2: 3.2 TB/s measured, profiles/r02c_profile_ops.log).
template <int U>
__global__ void __launch_bounds__(kEwThreads, 3) bn_bwd_apply_fast_kernel(const BnBwd b, const int lg_cg, const int chunk) {
  pdl_trigger();
  pdl_wait();
  const int cgm = (1 << lg_cg) - 1;
  const int c8 = (threadIdx.x & cgm) * 8;
  const double inv_n = 1.0 / (double)b.npix;
  if (b.dgamma && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
      b.dgamma[c] += (float)(b.sum_gx[c]);
      b.dbeta[c] += (float)(b.sum_g[c]);
    }
  }
  // dy = sc*(g - m1 - (x - mu)*rs*m2) = sc*g + k1*x + k0
  float sc[8], sh[8], k1[8], k0[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = b.scale[c8 + j];
    sh[j] = b.shift[c8 + j];
    const float m1 = (float)(b.sum_g[c8 + j] * inv_n);
    const float m2 = (float)(b.sum_gx[c8 + j] * inv_n);
    const float t = sc[j] * b.rstd[c8 + j] * m2;
    k1[j] = -t;
    k0[j] = t * b.mean[c8 + j] - sc[j] * m1;
  }
  const long long total = b.npix << lg_cg;
  const long long i0 = (long long)blockIdx.x * chunk;
  const long long i1 = (i0 + chunk < total) ? i0 + chunk : total;
  for (long long it0 = i0 + threadIdx.x; it0 < i1; it0 += U * kEwThreads) {
    float x[U][8], d[U][8];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (it0 + u * kEwThreads < i1) bn_load8(b, (size_t)(it0 + u * kEwThreads) * 8, x[u], d[u]);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (it0 + u * kEwThreads >= i1) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float gq = d[u][j];
        if (b.relu && fmaf(x[u][j], sc[j], sh[j]) <= 0.f) gq = 0.f;
        o[j] = fmaf(sc[j], gq, fmaf(k1[j], x[u][j], k0[j]));
      }
      *reinterpret_cast<uint4*>(b.dy + (size_t)(it0 + u * kEwThreads) * 8) = pack8(o);
    }
  }
}

// -------------------------------------------------------------------------------------------- fold grad
__global__ void fold_grad_kernel(const FoldK f) {
  pdl_trigger();
  pdl_wait();
  const int cg = f.C / 4;
  const int OH = (f.up || f.dilate) ? 2 * f.H : f.H, OW = (f.up || f.dilate) ? 2 * f.W : f.W;
  const int Hq = OH + 2 * f.P, Wq = OW + 2 * f.P;
  const long long total = (long long)f.N * f.H * f.W * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cg) * 4;
    long long t = i / cg;
    const int x = (int)(t % f.W);
    t /= f.W;
    const int y = (int)(t % f.H);
    const int n = (int)(t / f.H);
    // hi-res rows / cols fed by (y, x) and their weights
    int Ys[4], Xs[4], ny = 0, nx = 0;
    float wys[4], wxs[4];
    if (f.up) {
      for (int Y = 2 * y - 1; Y <= 2 * y + 2; Y++) {
        if (Y < 0 || Y >= OH) continue;
        int a0, a1;
        float w1;
        up_coord(Y, f.H, f.up, a0, a1, w1);
        const float w = (a0 == y ? 1.f - w1 : 0.f) + (a1 == y ? w1 : 0.f);
        if (w != 0.f) { Ys[ny] = Y; wys[ny++] = w; }
      }
      for (int X = 2 * x - 1; X <= 2 * x + 2; X++) {
        if (X < 0 || X >= OW) continue;
        int a0, a1;
        float w1;
        up_coord(X, f.W, f.up, a0, a1, w1);
        const float w = (a0 == x ? 1.f - w1 : 0.f) + (a1 == x ? w1 : 0.f);
        if (w != 0.f) { Xs[nx] = X; wxs[nx++] = w; }
      }
    } else {
      Ys[0] = f.dilate ? 2 * y : y; wys[0] = 1.f; ny = 1;
      Xs[0] = f.dilate ? 2 * x : x; wxs[0] = 1.f; nx = 1;
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < ny; a++) {
      int my[3];
      const int cy = mirror_set(Ys[a], OH, f.P, f.reflect, my);
      for (int b = 0; b < nx; b++) {
        int mx[3];
        const int cx = mirror_set(Xs[b], OW, f.P, f.reflect, mx);
        const float w = wys[a] * wxs[b];
        for (int p = 0; p < cy; p++)
          for (int q = 0; q < cx; q++) {
            const FoldV4 v = fold_load4(f, (size_t)((((long long)n * Hq + my[p]) * Wq + mx[q]) * f.ctot + f.c_off + c4));
            acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
          }
      }
    }
    float* o = f.dact + (((long long)n * f.H + y) * f.W + x) * f.C + c4;
    if (f.accumulate) {
      const float4 v = *reinterpret_cast<const float4*>(o);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}


// fast path: one source row (n, y) per CTA iteration.  The column candidates / weights of every source column are
// tabulated ONCE per CTA in shared memory (they do not depend on the row or the channel), the hi-res rows feeding a
// source row once per row; the per-item loop is loads + FMAs only (a first version that recomputed the candidates per
// item was ALU-bound at ~1 TB/s on the x2-bilinear adjoints; same-box A/B in profiles/r02a_*).  Dynamic shared memory:
// W * 64 bytes.
// (A version with 8 channels per thread and the up-to-12 candidate loads of a buffer row issued together was measured and
// rejected: 1.21 -> 1.71 ms per step over the eight folds, profiles/r02u_profile_ops.log -- 128 registers, two CTAs per SM.)
__global__ void __launch_bounds__(kEwThreads) fold_rows2_kernel(const FoldK f, const int lg_cg, const int rows) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char s_fold[];
  int* s_pc = reinterpret_cast<int*>(s_fold);                                  // [W][12]
  float* s_qw = reinterpret_cast<float*>(s_fold + (size_t)f.W * kFoldColInts * sizeof(int));   // [W][4]
  __shared__ int s_prow[12];
  __shared__ float s_pw[12];
  __shared__ int s_np;
  const int OH = (f.up || f.dilate) ? 2 * f.H : f.H, OW = (f.up || f.dilate) ? 2 * f.W : f.W;
  const int Hq = OH + 2 * f.P, Wq = OW + 2 * f.P;
  for (int x = threadIdx.x; x < f.W; x += kEwThreads) fold_col_entry(f, OW, x, s_pc + x * kFoldColInts, s_qw + x * 4);
  const int items = f.W << lg_cg;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) s_np = fold_row_entry(f, OH, row % f.H, s_prow, s_pw);
    __syncthreads();
    const int np = s_np;
    for (int it = threadIdx.x; it < items; it += kEwThreads) fold_item(f, lg_cg, Hq, Wq, row, it, np, s_prow, s_pw, s_pc, s_qw);
  }
}

// ---- backward through a FROZEN (eval-mode, BatchNorm folded) unit: dy = g * [y > 0] * scale[c] as bf16.
// y is the unit's own post-ReLU output (fp32 stream or plain bf16 buffer); the folded BatchNorm scale is applied to the
// gradient here so that the input-gradient convolution can use the un-folded weight pack.
struct FrozenBwd {
  const float* dact;
  const float* y_f32;
  const __nv_bfloat16* y_bf16;
  const float* scale;
  int relu, C;
  long long items;   // n*h*w*C/8
  __nv_bfloat16* dy;
};

__global__ void __launch_bounds__(kEwThreads) frozen_bwd_kernel(const FrozenBwd b) {
  pdl_trigger();
  pdl_wait();
  const int cg = b.C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b.items; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cg) * 8;
    const size_t off = (size_t)i * 8;
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(b.dact + off));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(b.dact + off + 4));
    float g[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    if (b.relu) {
      float y[8];
      if (b.y_f32) {
        const float4 y0 = __ldg(reinterpret_cast<const float4*>(b.y_f32 + off));
        const float4 y1 = __ldg(reinterpret_cast<const float4*>(b.y_f32 + off + 4));
        y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w; y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
      } else {
        unpack8(__ldg(reinterpret_cast<const uint4*>(b.y_bf16 + off)), y);
      }
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (!(y[j] > 0.f)) g[j] = 0.f;
    }
    if (b.scale) {
#pragma unroll
      for (int j = 0; j < 8; j++) g[j] *= __ldg(b.scale + c8 + j);
    }
    *reinterpret_cast<uint4*>(b.dy + off) = pack8(g);
  }
}

// adjoint of reflection padding for thin (C not a multiple of 4) tensors: the input gradient of a first layer
__global__ void fold_thin_kernel(const FoldK f) {
  pdl_trigger();
  pdl_wait();
  const int Hq = f.H + 2 * f.P, Wq = f.W + 2 * f.P;
  const long long total = (long long)f.N * f.H * f.W * f.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % f.C);
    long long t = i / f.C;
    const int x = (int)(t % f.W);
    t /= f.W;
    const int y = (int)(t % f.H);
    const long long n = t / f.H;
    int my[3], mx[3];
    const int cy = mirror_set(y, f.H, f.P, f.reflect, my);
    const int cx = mirror_set(x, f.W, f.P, f.reflect, mx);
    float acc = f.accumulate ? f.dact[i] : 0.f;
    for (int p = 0; p < cy; p++)
      for (int q = 0; q < cx; q++) acc += f.dpad[((n * Hq + my[p]) * Wq + mx[q]) * f.ctot + f.c_off + c];
    f.dact[i] = acc;
  }
}

// -------------------------------------------------------------------------------- weight pack / unpack
// packed[t][a][b] (t = r*kw + s, a < A, b < B) <-> w[a*sa + b*sb + r'*sr + s'*ss], (r', s') flipped when flip.
// For im2col'd layers (col_c > 0): packed[0][a][k], k = (r*kw + s)*col_c + c  <->  w[a*sa + c*sb + r*sr + s*ss].
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ scale_a,
                                    __nv_bfloat16* __restrict__ out, const PackK k) {
  pdl_trigger();
  pdl_wait();
  const int T = k.col_c ? 1 : k.kh * k.kw;
  const long long total = (long long)T * k.Apad * k.Bpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % k.Bpad);
    const int a = (int)((i / k.Bpad) % k.Apad);
    const int t = (int)(i / ((long long)k.Bpad * k.Apad));
    float v = 0.f;
    if (a < k.A) {
      if (k.col_c) {
        const int c = b % k.col_c, tt = b / k.col_c;
        if (tt < k.kh * k.kw) {
          int r = tt / k.kw, s2 = tt % k.kw;
          if (k.flip) { r = k.kh - 1 - r; s2 = k.kw - 1 - s2; }
          v = w[a * k.sa + c * k.sb + r * k.sr + s2 * k.ss];
        }
      } else if (b < k.B) {
        int r = t / k.kw, s = t % k.kw;
        if (k.flip) { r = k.kh - 1 - r; s = k.kw - 1 - s; }
        v = w[a * k.sa + b * k.sb + r * k.sr + s * k.ss];
      }
      if (scale_a) v *= scale_a[a];
    }
    out[i] = __float2bfloat16(v);
  }
}

// grad[w index] (+)= packed_dw[t][b][a]   (wgrad layout: [tap][ci = b][co = a], co padded to Apad)
// (both unpack kernels: `slabs` > 1 = the deterministic split-K layout -- slab s of the weight gradient lies slab_elems
// floats after slab s - 1; the slabs are summed in index order)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ grad, const PackK k, int accumulate,
                                    const int slabs, const long long slab_elems) {
  pdl_trigger();
  pdl_wait();
  const int T = k.col_c ? 1 : k.kh * k.kw;
  const long long total = (long long)T * k.B * k.A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % k.A);
    const int b = (int)((i / k.A) % k.B);
    const int t = (int)(i / ((long long)k.A * k.B));
    long long widx;
    if (k.col_c) {
      const int c = b % k.col_c, tt = b / k.col_c;
      if (tt >= k.kh * k.kw) continue;
      int r = tt / k.kw, s2 = tt % k.kw;
      if (k.flip) { r = k.kh - 1 - r; s2 = k.kw - 1 - s2; }
      widx = a * k.sa + c * k.sb + r * k.sr + s2 * k.ss;
    } else {
      int r = t / k.kw, s = t % k.kw;
      if (k.flip) { r = k.kh - 1 - r; s = k.kw - 1 - s; }
      widx = a * k.sa + b * k.sb + r * k.sr + s * k.ss;
    }
    const float* src = dw + ((long long)t * k.Bpad + b) * k.Apad + a;
    float v = src[0];
    for (int sl = 1; sl < slabs; sl++) v += src[(long long)sl * slab_elems];
    if (accumulate) grad[widx] += v; else grad[widx] = v;
  }
}


// fast paths for ordinary (non-im2col) layers whose taps are contiguous in the parameter tensor (stride_s == 1,
// stride_r == kw): a CTA moves a 16(a) x 16(b) tile for ALL taps through shared memory so that both the fp32
// parameter side (runs of kh*kw floats per (a, b)) and the packed side (b or a fastest) are accessed in full sectors.
constexpr int kPackTile = 16;
constexpr int kPackMaxTaps = 81;

// All weight packs of a network in ONE launch: a device table of jobs (one per packed tensor), CTA c works on tile
// (c - cta0) of the job whose [cta0, cta0 + tiles) range contains it.  A training step re-packs ~90 tensors; as
// separate launches they are latency-bound (~16 us each), as one launch the whole re-pack is a single HBM-bound pass.
struct PackJob {
  PackK k;
  const float* w;
  const float* scale_a;
  __nv_bfloat16* out;
  int cta0, tiles_x;
  int tb;            // b-tile width (pack_tile.cuh: 16 x 64 / 32 / 16 tiles by tap count)
};

__global__ void __launch_bounds__(256) pack_tile2_kernel(const float* __restrict__ w, const float* __restrict__ scale_a,
                                                        __nv_bfloat16* __restrict__ out, const PackK k, const int tb) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_tile[];
  const int a0 = blockIdx.y * kPackTA, b0 = blockIdx.x * tb;
  pack_v2_phase1(w, k, a0, b0, tb, s_tile, threadIdx.x, blockDim.x);
  __syncthreads();
  pack_v2_phase2(scale_a, out, k, a0, b0, tb, s_tile, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) pack_table_kernel(const PackJob* __restrict__ jobs, const int njobs) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_tile[];
  __shared__ PackJob job;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {  // last job with cta0 <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].cta0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    job = jobs[lo];
  }
  __syncthreads();
  const int tile = blockIdx.x - job.cta0;
  const int a0 = (tile / job.tiles_x) * kPackTA, b0 = (tile % job.tiles_x) * job.tb;
  pack_v2_phase1(job.w, job.k, a0, b0, job.tb, s_tile, threadIdx.x, blockDim.x);
  __syncthreads();
  pack_v2_phase2(job.scale_a, job.out, job.k, a0, b0, job.tb, s_tile, threadIdx.x, blockDim.x);
}

// grad[a*sa + b*sb + tap] (+)= dw[t][b][a]
__global__ void __launch_bounds__(256) unpack_tile_kernel(const float* __restrict__ dw, float* __restrict__ grad, const PackK k,
                                                         int accumulate, const int slabs, const long long slab_elems) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_tile[];
  const int T = k.kh * k.kw;
  const int a0 = blockIdx.y * kPackTile, b0 = blockIdx.x * kPackTile;
  for (int i = threadIdx.x; i < T * kPackTile * kPackTile; i += blockDim.x) {
    const int al = i & (kPackTile - 1), bl = (i >> 4) & (kPackTile - 1), t = i >> 8;
    const int a = a0 + al, b = b0 + bl;
    float v = 0.f;
    if (a < k.A && b < k.B) {
      const float* src = dw + ((long long)t * k.Bpad + b) * k.Apad + a;
      v = __ldg(src);
      for (int sl = 1; sl < slabs; sl++) v += __ldg(src + (long long)sl * slab_elems);
    }
    const int tap = k.flip ? T - 1 - t : t;
    s_tile[(al * kPackTile + bl) * (T + 1) + tap] = v;
  }
  __syncthreads();
  const bool b_inner = (k.sb == T);
  const int run = kPackTile * T;
  for (int i = threadIdx.x; i < kPackTile * run; i += blockDim.x) {
    const int outer = i / run, rem = i - outer * run;
    const int inner = rem / T, tap = rem - inner * T;
    const int al = b_inner ? outer : inner, bl = b_inner ? inner : outer;
    const int a = a0 + al, b = b0 + bl;
    if (a >= k.A || b >= k.B) continue;
    float* g = grad + (long long)a * k.sa + (long long)b * k.sb + tap;
    const float v = s_tile[(al * kPackTile + bl) * (T + 1) + tap];
    if (accumulate) *g += v; else *g = v;
  }
}

static bool pack_tileable(const PackK& k) {
  const int T = k.kh * k.kw;
  if (k.col_c || T > kPackMaxTaps) return false;
  if (!(k.ss == 1 && (k.sr == k.kw || k.kh == 1))) return false;
  // the inner index must be contiguous runs of T floats: [a][b][T] with sa >= B*T, or [b][a][T] with sb >= A*T
  return (k.sb == T && k.sa >= (long long)k.B * T) || (k.sa == T && k.sb >= (long long)k.A * T);
}

// eval-mode BatchNorm folded into a per-channel scale / bias
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rmean, const float* __restrict__ rvar, float eps,
                               float* __restrict__ scale, float* __restrict__ bias, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] / sqrtf(rvar[c] + eps);
  scale[c] = s;
  bias[c] = beta[c] - rmean[c] * s;
}

}  // namespace gdn

using namespace gdn;
#define GDN_API extern "C" __attribute__((visibility("default")))

GDN_API int gdn_im2col(const float* src, void* dst, int n, int c, int h, int w, int kh, int kw, int pad, int reflect,
                       int kpad, gdn_stream stream) {
  if (!src || !dst || c < 1 || c > 4 || kpad % 64 || kh * kw * c > kpad)
    return fail(GDN_INVALID_DESC, "gdn_im2col: bad arguments (c=%d kpad=%d)", c, kpad);
  if (reflect && (pad >= h || pad >= w)) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_im2col: reflection pad %d >= extent", pad);
  {
    // band of output rows per CTA iteration: as many as fit 200 KB of shared memory next to the kh - 1 halo rows (<= 8)
    const size_t row_bytes = (size_t)(w + 2 * pad) * c * sizeof(float);
    int band = (int)((200 * 1024) / row_bytes) - (kh - 1);
    if (band > kIm2colBand) band = kIm2colBand;
    if (kpad <= 256 && kpad % 64 == 0 && band >= 1 && (long long)h * w * c < (1ll << 31)) {
      const size_t smem = (size_t)(kh - 1 + band) * row_bytes;
      static bool configured[64] = {false};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < 64 && !configured[dev]) {
        GDN_CUDA_CHECK(cudaFuncSetAttribute(im2col_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev] = true;
      }
      GDN_CUDA_CHECK(launch_pdl(im2col_smem_kernel, dim3(ew_row_grid((long long)n * ((h + band - 1) / band))), dim3(kEwThreads), smem, (cudaStream_t)stream, 1, src, (__nv_bfloat16*)dst, n, c, h, w, kh, kw, pad, reflect, kpad, band));
      GDN_LAUNCH_CHECK("im2col_smem_kernel");
      return GDN_OK;
    }
  }
  const long long work = (long long)n * h * w * (kpad / 8);
  GDN_CUDA_CHECK(launch_pdl(im2col_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, src, (__nv_bfloat16*)dst, n, c, h, w, kh, kw, pad, reflect, kpad));
  GDN_LAUNCH_CHECK("im2col_kernel");
  return GDN_OK;
}

GDN_API int gdn_bn_finalize(const double* sum, const double* sqsum, double count, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, float* scale,
                            float* shift, float* mean, float* rstd, float* coef4, int c, gdn_stream stream) {
  if (!sum || !sqsum || !gamma || !beta || !scale || !shift || !mean || !rstd || c < 1)
    return fail(GDN_INVALID_DESC, "gdn_bn_finalize: null pointer");
  GDN_CUDA_CHECK(launch_pdl(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, sum, sqsum, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, rstd, reinterpret_cast<float4*>(coef4), c));
  GDN_LAUNCH_CHECK("bn_finalize_kernel");
  return GDN_OK;
}

GDN_API int gdn_bn_fold(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps,
                        float* scale, float* bias, int c, gdn_stream stream) {
  GDN_CUDA_CHECK(launch_pdl(bn_fold_kernel, dim3((c + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1, gamma, beta, rmean, rvar, eps, scale, bias, c));
  GDN_LAUNCH_CHECK("bn_fold_kernel");
  return GDN_OK;
}

GDN_API int gdn_act_forward(const gdn_act_fwd_desc* d, gdn_stream stream) {
  if (!d || (!d->src_bf16 == !d->src_f32)) return fail(GDN_INVALID_DESC, "gdn_act_forward: exactly one source required");
  if (d->c % 8 || (kEwThreads % (d->c / 8)))
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_act_forward: channels %d (need a multiple of 8 whose /8 divides %d)", d->c, kEwThreads);
  if (d->up && d->dilate) return fail(GDN_INVALID_DESC, "gdn_act_forward: up and dilate are exclusive");
  const int OH = (d->up || d->dilate) ? 2 * d->h : d->h, OW = (d->up || d->dilate) ? 2 * d->w : d->w;
  if (d->reflect && (d->pad >= OH || d->pad >= OW)) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_act_forward: reflection pad too large");
  ActFwd a{};
  a.src_bf16 = (const __nv_bfloat16*)d->src_bf16;
  a.src_f32 = d->src_f32;
  a.scale = d->scale;
  a.shift = d->shift;
  a.resid = d->resid;
  a.relu = d->relu;
  a.N = d->n; a.H = d->h; a.W = d->w; a.C = d->c;
  a.out_f32 = d->out_f32;
  a.out_bf16 = (__nv_bfloat16*)d->out_bf16;
  a.P = d->pad; a.reflect = d->reflect; a.up = d->up; a.dilate = d->dilate;
  a.src_half = d->src16_is_half;
  const long long w1 = a.out_f32 ? (long long)a.N * a.H * a.W * (a.C / 8) : 0;
  const long long w2 = a.out_bf16 ? (long long)a.N * (OH + 2 * a.P) * (OW + 2 * a.P) * (a.C / 8) : 0;
  const long long work = w1 > w2 ? w1 : w2;
  if (work == 0) return GDN_OK;
  const int lg = ew_lg2(a.C / 8);
  const bool small_idx = (long long)a.W * (a.C / 8) * 4 < (1ll << 30) && (long long)a.N * (OH + 2 * a.P) < (1ll << 30);
  if (lg >= 0 && small_idx && !a.dilate && (a.up == 0 || a.up == 1)) {
    ActFwd lo = a;
    if (a.up) lo.out_bf16 = nullptr;
    if (lo.out_f32 || lo.out_bf16) {
      const int rows = a.N * a.H;
      if (lo.resid) GDN_CUDA_CHECK(launch_pdl(act_rows_kernel<true, 2>, dim3(ew_row_grid(rows)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, lo, lg, rows));
      else GDN_CUDA_CHECK(launch_pdl(act_rows_kernel<false, 4>, dim3(ew_row_grid(rows)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, lo, lg, rows));
      GDN_LAUNCH_CHECK("act_rows_kernel");
    }
    if (a.up && a.out_bf16) {
      const bool pure = a.src_bf16 && !a.src_half && !a.scale && !a.resid && !a.relu && !a.out_f32 && a.H >= 2 && a.W >= 2 &&
                        a.P < 2 * a.H && a.P < 2 * a.W;
      if (pure) {
        const int rows = a.N * (a.H + 1);
        GDN_CUDA_CHECK(launch_pdl(up2x_blocks_kernel, dim3(ew_row_grid(rows)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, a, lg, rows));
        GDN_LAUNCH_CHECK("up2x_blocks_kernel");
      } else {
        const int rows = a.N * (OH + 2 * a.P);
        GDN_CUDA_CHECK(launch_pdl(act_up_rows_kernel, dim3(ew_row_grid(rows)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, a, lg, rows));
        GDN_LAUNCH_CHECK("act_up_rows_kernel");
      }
    }
    return GDN_OK;
  }
  GDN_CUDA_CHECK(launch_pdl(act_forward_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, a));
  GDN_LAUNCH_CHECK("act_forward_kernel");
  return GDN_OK;
}

static int fill_bnbwd(const gdn_bn_bwd_desc* d, BnBwd& b) {
  if (!d || !d->dact || !d->raw || !d->scale || !d->shift || !d->mean || !d->rstd || !d->sum_g || !d->sum_gx)
    return fail(GDN_INVALID_DESC, "gdn_bn_bwd: null pointer");
  if (d->dact_is_bf16 && d->dilate) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_bn_bwd: bf16 gradient input with dilation");
  if (d->c % 8 || d->c > 512 || (256 % (d->c / 8))) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_bn_bwd: channels %d", d->c);
  b.dact = reinterpret_cast<const float*>(d->dact);
  b.raw = (const __nv_bfloat16*)d->raw;
  b.scale = d->scale; b.shift = d->shift; b.mean = d->mean; b.rstd = d->rstd;
  b.relu = d->relu;
  b.npix = (long long)d->n * d->h * d->w;
  b.C = d->c;
  b.sum_g = d->sum_g; b.sum_gx = d->sum_gx;
  b.dy = (__nv_bfloat16*)d->dy;
  b.H = d->h; b.W = d->w; b.dilate = d->dilate;
  b.dgamma = d->dgamma; b.dbeta = d->dbeta;
  b.raw_half = d->raw_is_half;
  b.dact_bf16 = d->dact_is_bf16;
  b.det = det_enabled() ? 1 : 0;
  return GDN_OK;
}

GDN_API int gdn_bn_bwd_reduce(const gdn_bn_bwd_desc* d, gdn_stream stream) {
  BnBwd b{};
  int rc = fill_bnbwd(d, b);
  if (rc) return rc;
  {
    const int lg = ew_lg2(b.C / 8);
    const long long total = b.npix * (b.C / 8);
    if (lg >= 0 && total < (1ll << 40)) {
      const long long want = (long long)device_sm_count() * 6;
      long long chunk = (total + want - 1) / want;
      chunk = (chunk + 2 * kEwThreads - 1) / (2 * kEwThreads) * (2 * kEwThreads);
      if (chunk < (1ll << 30)) {
        const int grid = (int)((total + chunk - 1) / chunk);
        GDN_CUDA_CHECK(launch_pdl(bn_bwd_reduce_fast_kernel, dim3(grid), dim3(kEwThreads), 2 * b.C * sizeof(float), (cudaStream_t)stream, 1, b, lg, (int)chunk));
        GDN_LAUNCH_CHECK("bn_bwd_reduce_fast_kernel");
        return GDN_OK;
      }
    }
  }
  const int lanes = kEwThreads / (b.C / 8);
  long long blocks = (b.npix + lanes - 1) / lanes;
  const long long cap = (long long)device_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  GDN_CUDA_CHECK(launch_pdl(bn_bwd_reduce_kernel, dim3((int)blocks), dim3(kEwThreads), 2 * b.C * sizeof(float), (cudaStream_t)stream, 1, b));
  GDN_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return GDN_OK;
}

GDN_API int gdn_act_backward(const gdn_bn_bwd_desc* d, gdn_stream stream) {
  BnBwd b{};
  int rc = fill_bnbwd(d, b);
  if (rc) return rc;
  if (!b.dy) return fail(GDN_INVALID_DESC, "gdn_act_backward: dy is NULL");
  if (!b.dilate) {
    const int lg = ew_lg2(b.C / 8);
    const long long total = b.npix * (b.C / 8);
    if (lg >= 0) {
      const long long want = (long long)device_sm_count() * 8;
      long long chunk = (total + want - 1) / want;
      chunk = (chunk + 4 * kEwThreads - 1) / (4 * kEwThreads) * (4 * kEwThreads);
      if (chunk < (1ll << 30)) {
        const int grid = (int)((total + chunk - 1) / chunk);
        if (b.dact_bf16) GDN_CUDA_CHECK(launch_pdl(bn_bwd_apply_fast_kernel<4>, dim3(grid), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, b, lg, (int)chunk));
        else GDN_CUDA_CHECK(launch_pdl(bn_bwd_apply_fast_kernel<2>, dim3(grid), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, b, lg, (int)chunk));
        GDN_LAUNCH_CHECK("bn_bwd_apply_fast_kernel");
        return GDN_OK;
      }
    }
  }
  const long long work = b.npix * (b.dilate ? 4 : 1) * (b.C / 8);
  GDN_CUDA_CHECK(launch_pdl(bn_bwd_apply_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, b));
  GDN_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return GDN_OK;
}

GDN_API int gdn_fold_grad(const gdn_fold_desc* d, gdn_stream stream) {
  if (!d || !d->dpad || !d->dact) return fail(GDN_INVALID_DESC, "gdn_fold_grad: null pointer");
  const bool thin = (d->c % 4 || d->ctot % 4 || d->c_off % 4);
  if (thin && (d->up || d->dilate))
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_fold_grad: channel alignment (thin tensors: reflection / zero padding only)");
  if (d->up > 1) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_fold_grad: only the align_corners=False adjoint is implemented");
  if (d->reflect && (d->pad >= d->h || d->pad >= d->w)) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_fold_grad: reflection pad too large");
  FoldK f{};
  f.dpad = reinterpret_cast<const float*>(d->dpad); f.ctot = d->ctot; f.c_off = d->c_off;
  f.N = d->n; f.H = d->h; f.W = d->w; f.C = d->c;
  f.P = d->pad; f.reflect = d->reflect; f.up = d->up; f.dilate = d->dilate;
  f.dact = d->dact; f.accumulate = d->accumulate;
  f.dpad_bf16 = d->dpad_is_bf16;
  if (thin && f.dpad_bf16) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_fold_grad: bf16 input needs channel counts that are multiples of 4");
  if (thin) {
    GDN_CUDA_CHECK(launch_pdl(fold_thin_kernel, dim3(ew_grid((long long)f.N * f.H * f.W * f.C)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, f));
    GDN_LAUNCH_CHECK("fold_thin_kernel");
    return GDN_OK;
  }
  {
    const int lg = ew_lg2(f.C / 4);
    if (lg >= 0 && (long long)f.W * (f.C / 4) < (1ll << 30) && (long long)f.N * f.H < (1ll << 30)) {
      const int rows = f.N * f.H;
      const size_t smem = (size_t)f.W * (kFoldColInts * sizeof(int) + 4 * sizeof(float));
      if (smem <= 40 * 1024) {
        GDN_CUDA_CHECK(launch_pdl(fold_rows2_kernel, dim3(ew_row_grid(rows)), dim3(kEwThreads), smem, (cudaStream_t)stream, 1, f, lg, rows));
        GDN_LAUNCH_CHECK("fold_rows2_kernel");
        return GDN_OK;
      }
    }
  }
  const long long work = (long long)f.N * f.H * f.W * (f.C / 4);
  GDN_CUDA_CHECK(launch_pdl(fold_grad_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, f));
  GDN_LAUNCH_CHECK("fold_grad_kernel");
  return GDN_OK;
}

GDN_API int gdn_act_backward_frozen(const gdn_frozen_bwd_desc* d, gdn_stream stream) {
  if (!d || !d->dact || !d->dy) return fail(GDN_INVALID_DESC, "gdn_act_backward_frozen: null pointer");
  if (d->relu && !d->y_f32 && !d->y_bf16) return fail(GDN_INVALID_DESC, "gdn_act_backward_frozen: ReLU needs the unit's output");
  if (d->c % 8) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_act_backward_frozen: channels %d (need a multiple of 8)", d->c);
  FrozenBwd b{};
  b.dact = d->dact; b.y_f32 = d->y_f32; b.y_bf16 = (const __nv_bfloat16*)d->y_bf16; b.scale = d->scale;
  b.relu = d->relu; b.C = d->c;
  b.items = (long long)d->n * d->h * d->w * (d->c / 8);
  b.dy = (__nv_bfloat16*)d->dy;
  if (b.items == 0) return GDN_OK;
  GDN_CUDA_CHECK(launch_pdl(frozen_bwd_kernel, dim3(ew_grid(b.items)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, b));
  GDN_LAUNCH_CHECK("frozen_bwd_kernel");
  return GDN_OK;
}

static void fill_pack(const gdn_pack_desc* d, PackK& k) {
  k.kh = d->kh; k.kw = d->kw; k.A = d->a; k.B = d->b; k.Apad = d->a_pad; k.Bpad = d->b_pad;
  k.sa = d->stride_a; k.sb = d->stride_b; k.sr = d->stride_r; k.ss = d->stride_s;
  k.flip = d->flip; k.col_c = d->col_c;
}

GDN_API int gdn_pack_weights(const gdn_pack_desc* d, const float* w, const float* scale_a, void* out, gdn_stream stream) {
  if (!d || !w || !out || d->a_pad < d->a || d->b_pad < d->b) return fail(GDN_INVALID_DESC, "gdn_pack_weights: bad arguments");
  PackK k;
  fill_pack(d, k);
  if (pack_tileable(k)) {
    const int T = k.kh * k.kw;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
      GDN_CUDA_CHECK(cudaFuncSetAttribute(pack_tile2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      configured[dev] = true;
    }
    const int tb = pack_tb_for_taps(T);
    dim3 grid2((k.Bpad + tb - 1) / tb, (k.Apad + kPackTA - 1) / kPackTA);
    GDN_CUDA_CHECK(launch_pdl(pack_tile2_kernel, dim3(grid2), dim3(256), pack_smem_bytes(T), (cudaStream_t)stream, 1, w, scale_a, (__nv_bfloat16*)out, k, tb));
    GDN_LAUNCH_CHECK("pack_tile2_kernel");
    return GDN_OK;
  }
  const long long work = (long long)(k.col_c ? 1 : k.kh * k.kw) * k.Apad * k.Bpad;
  GDN_CUDA_CHECK(launch_pdl(pack_weights_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, w, scale_a, (__nv_bfloat16*)out, k));
  GDN_LAUNCH_CHECK("pack_weights_kernel");
  return GDN_OK;
}

GDN_API int gdn_pack_job_size(void) { return (int)sizeof(PackJob); }

// Fill one entry of a host-side job table (job_out: gdn_pack_job_size() bytes).  *n_ctas receives the number of CTAs
// the job needs; the caller lays the jobs out back to back (cta0 = running sum) and copies the table to the device.
// Returns GDN_UNSUPPORTED_SHAPE when the tensor is not tileable (im2col'd thin layers): use gdn_pack_weights for it.
GDN_API int gdn_pack_job_fill(const gdn_pack_desc* d, const float* w, const float* scale_a, void* out, int cta0,
                              void* job_out, int* n_ctas) {
  if (!d || !w || !out || !job_out || !n_ctas || d->a_pad < d->a || d->b_pad < d->b)
    return fail(GDN_INVALID_DESC, "gdn_pack_job_fill: bad arguments");
  PackJob j;
  fill_pack(d, j.k);
  if (!pack_tileable(j.k)) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_pack_job_fill: tensor is not tileable");
  j.w = w;
  j.scale_a = scale_a;
  j.out = (__nv_bfloat16*)out;
  j.cta0 = cta0;
  j.tb = pack_tb_for_taps(j.k.kh * j.k.kw);
  j.tiles_x = (j.k.Bpad + j.tb - 1) / j.tb;
  *n_ctas = j.tiles_x * ((j.k.Apad + kPackTA - 1) / kPackTA);
  memcpy(job_out, &j, sizeof(j));
  return GDN_OK;
}

// one launch for a whole device-resident job table (total_ctas = sum of the jobs' CTA counts, max_taps = largest kh*kw)
GDN_API int gdn_pack_weights_table(const void* jobs_dev, int njobs, int total_ctas, int max_taps, gdn_stream stream) {
  if (!jobs_dev || njobs <= 0 || total_ctas <= 0 || max_taps <= 0 || max_taps > kPackMaxTaps)
    return fail(GDN_INVALID_DESC, "gdn_pack_weights_table: bad arguments");
  // every job of one table has max_taps taps (the caller groups by kernel size); the tile is sized by the tap count
  size_t smem = 0;
  for (int t = 1; t <= max_taps; t++)      // a table may mix tap counts <= max_taps; the tile size is not monotonic in t
    if (pack_smem_bytes(t) > smem) smem = pack_smem_bytes(t);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    GDN_CUDA_CHECK(cudaFuncSetAttribute(pack_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured[dev] = true;
  }
  GDN_CUDA_CHECK(launch_pdl(pack_table_kernel, dim3(total_ctas), dim3(256), smem, (cudaStream_t)stream, 1, (const PackJob*)jobs_dev, njobs));
  GDN_LAUNCH_CHECK("pack_table_kernel");
  return GDN_OK;
}

GDN_API int gdn_unpack_wgrad_slabs(const gdn_pack_desc* d, const float* dw, int slabs, int64_t slab_elems, float* grad,
                                   int accumulate, gdn_stream stream);
GDN_API int gdn_unpack_wgrad(const gdn_pack_desc* d, const float* dw, float* grad, int accumulate, gdn_stream stream) {
  return gdn_unpack_wgrad_slabs(d, dw, 1, 0, grad, accumulate, stream);
}

GDN_API int gdn_unpack_wgrad_slabs(const gdn_pack_desc* d, const float* dw, int slabs, int64_t slab_elems, float* grad,
                                   int accumulate, gdn_stream stream) {
  if (!d || !dw || !grad) return fail(GDN_INVALID_DESC, "gdn_unpack_wgrad: null pointer");
  if (slabs < 1 || (slabs > 1 && slab_elems <= 0)) return fail(GDN_INVALID_DESC, "gdn_unpack_wgrad: %d slabs of %lld floats", slabs, (long long)slab_elems);
  const long long se = (long long)slab_elems;
  PackK k;
  fill_pack(d, k);
  // the tile kernel launches one CTA per 16 x 16 (a, b) tile for ALL taps: a 64 x 64 tensor with 81 taps would be 16 CTAs
  // walking 20 736 elements each (90 us for 1.3 MB, profiles/r02r_ncu_elem.summary.txt) -- such tensors take the
  // element-parallel kernel below (coalesced reads, scattered 4-byte writes that the L2 absorbs)
  const bool enough_tiles = ((k.A + kPackTile - 1) / kPackTile) * ((k.B + kPackTile - 1) / kPackTile) >= 64;
  if (pack_tileable(k) && enough_tiles) {
    const int T = k.kh * k.kw;
    const size_t smem = (size_t)kPackTile * kPackTile * (T + 1) * sizeof(float);
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
      GDN_CUDA_CHECK(cudaFuncSetAttribute(unpack_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      configured[dev] = true;
    }
    dim3 grid((k.B + kPackTile - 1) / kPackTile, (k.A + kPackTile - 1) / kPackTile);
    GDN_CUDA_CHECK(launch_pdl(unpack_tile_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, 1, dw, grad, k, accumulate, slabs, se));
    GDN_LAUNCH_CHECK("unpack_tile_kernel");
    return GDN_OK;
  }
  const long long work = (long long)(k.col_c ? 1 : k.kh * k.kw) * k.A * k.B;
  GDN_CUDA_CHECK(launch_pdl(unpack_wgrad_kernel, dim3(ew_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 1, dw, grad, k, accumulate, slabs, se));
  GDN_LAUNCH_CHECK("unpack_wgrad_kernel");
  return GDN_OK;
}
