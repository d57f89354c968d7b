// Training loss (fwd + analytic gradient), feature MSE, Eigen depth metrics and fused Adam for sm_100a.
// All of it is HBM / latency bound single-pass work: vector loads, warp-shuffle + shared-memory reductions,
// one fp64 atomic per block.  The metric kernel reproduces the reference's fp32 operation order exactly
// (no FMA contraction, IEEE division) so the delta-threshold pixel COUNTS are bit-exact.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gdn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum_d(double (&v)[NV], double* smem /* [32*NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum_d(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) smem[warp * NV + i] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double t = lane < nw ? smem[lane * NV + i] : 0.0;
      v[i] = warp_sum_d(t);
    }
  }
}

// --------------------------------------------------------------------------------------- max |a - b|
__global__ void absdiff_max_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                   unsigned int* __restrict__ out_bits) {
  pdl_trigger();
  pdl_wait();
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(a[i] - b[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = lane < (blockDim.x >> 5) ? s[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order like their bit patterns
  }
}

// ------------------------------------------------------------------------------------------------ loss
struct LossK {
  const float* out;
  const float* gt;
  const float* sparse;
  long long sparse_stride;
  const float* rgb;
  int N, H, W;
  const float* maxabs;
  int mode;  // 0 RtoD, 1 DtoD
  int cy1, cy2, cx1, cx2;
  float inv_count;
  double* sums;
  float* dout;
  float* dpre;
  float grad_scale;
};

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// Sobel cross-correlation (zero padding) of a single-channel map at (y, x); fx = [[1,0,-1],[2,0,-2],[1,0,-1]],
// fy = [[1,2,1],[0,0,0],[-1,-2,-1]]  (utils.py:107,115)
__device__ __forceinline__ void sobel_at(const float* __restrict__ img, int H, int W, int y, int x, float& gx, float& gy) {
  float v[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      const int yy = y + a - 1, xx = x + b - 1;
      v[a][b] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(long long)yy * W + xx] : 0.f;
    }
  gx = (v[0][0] - v[0][2]) + 2.f * (v[1][0] - v[1][2]) + (v[2][0] - v[2][2]);
  gy = (v[0][0] + 2.f * v[0][1] + v[0][2]) - (v[2][0] + 2.f * v[2][1] + v[2][2]);
}

__global__ void loss_kernel(const LossK k) {
  pdl_trigger();
  pdl_wait();
  const long long HW = (long long)k.H * k.W;
  const long long total = (long long)k.N * HW;
  const float c = 0.2f * (*k.maxabs);
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % k.W);
    const int y = (int)((i / k.W) % k.H);
    const long long n = i / HW;
    const float* o = k.out + n * HW;
    const float* g = k.gt + n * HW;
    const float ov = o[(long long)y * k.W + x];
    const float d = ov - g[(long long)y * k.W + x];
    const float a = fabsf(d);
    // masked BerHu (trainer.py:711-720)
    float wgt = 1.f;
    if (k.sparse) {
      const bool crop = (y >= k.cy1 && y < k.cy2 && x >= k.cx1 && x < k.cx2);
      if (!crop) wgt = 0.1f;
      else if (!(k.sparse[n * k.sparse_stride + (long long)y * k.W + x] > -1.f)) wgt = 0.3f;
    }
    float val, dv;
    if (a > c) {
      val = (d * d + c * c) / (2.f * c);
      dv = d / c;
    } else {
      val = a;
      dv = sgn(d);
    }
    acc[0] += (double)(val * wgt);
    acc[2] += (double)(d * d);
    float grad = 3.f * k.inv_count * wgt * dv;
    if (k.mode == 0) {
      // edge-aware smoothness (utils.py:139-178, trainer.py:753-754)
      const float* I = k.rgb + n * 3 * HW;
      auto wx_at = [&](int yy, int xx) {  // weight of the forward difference starting at (yy, xx), xx < W-1
        float s = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) s += fabsf(I[ch * HW + (long long)yy * k.W + xx] - I[ch * HW + (long long)yy * k.W + xx + 1]);
        return expf(-s / 3.f);
      };
      auto wy_at = [&](int yy, int xx) {
        float s = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) s += fabsf(I[ch * HW + (long long)yy * k.W + xx] - I[ch * HW + (long long)(yy + 1) * k.W + xx]);
        return expf(-s / 3.f);
      };
      float sm = 0.f, gs = 0.f;
      if (x < k.W - 1) {
        const float gx = ov - o[(long long)y * k.W + x + 1];
        const float w = wx_at(y, x);
        sm += fabsf(gx) * w;
        gs += sgn(gx) * w;
      }
      if (x >= 1) {
        const float gx = o[(long long)y * k.W + x - 1] - ov;
        gs -= sgn(gx) * wx_at(y, x - 1);
      }
      if (y < k.H - 1) {
        const float gy = ov - o[(long long)(y + 1) * k.W + x];
        const float w = wy_at(y, x);
        sm += fabsf(gy) * w;
        gs += sgn(gy) * w;
      }
      if (y >= 1) {
        const float gy = o[(long long)(y - 1) * k.W + x] - ov;
        gs -= sgn(gy) * wy_at(y - 1, x);
      }
      acc[1] += (double)sm;
      grad += 0.1f * k.inv_count * gs;
    } else {
      // 3 * imgrad_loss (utils.py:105-133, trainer.py:453)
      float gxo, gyo, gxt, gyt;
      sobel_at(o, k.H, k.W, y, x, gxo, gyo);
      sobel_at(g, k.H, k.W, y, x, gxt, gyt);
      acc[1] += (double)(fabsf(gxo - gxt) + fabsf(gyo - gyt));
      // adjoint: d/d out(y,x) = sum over neighbours p' of sign(dG(p')) * f[y - y' + 1][x - x' + 1]
      const float fx[3][3] = {{1.f, 0.f, -1.f}, {2.f, 0.f, -2.f}, {1.f, 0.f, -1.f}};
      const float fy[3][3] = {{1.f, 2.f, 1.f}, {0.f, 0.f, 0.f}, {-1.f, -2.f, -1.f}};
      float gs = 0.f;
#pragma unroll
      for (int a2 = 0; a2 < 3; a2++)
#pragma unroll
        for (int b2 = 0; b2 < 3; b2++) {
          const int yy = y - (a2 - 1), xx = x - (b2 - 1);  // output position whose window tap (a2,b2) is (y,x)
          if (yy < 0 || yy >= k.H || xx < 0 || xx >= k.W) continue;
          float pxo, pyo, pxt, pyt;
          sobel_at(o, k.H, k.W, yy, xx, pxo, pyo);
          sobel_at(g, k.H, k.W, yy, xx, pxt, pyt);
          gs += sgn(pxo - pxt) * fx[a2][b2] + sgn(pyo - pyt) * fy[a2][b2];
        }
      grad += 3.f * k.inv_count * gs;
    }
    grad *= k.grad_scale;
    if (k.dout) k.dout[i] = grad;
    if (k.dpre) k.dpre[i] = grad * (1.f - ov * ov);  // through tanh
  }
  __shared__ double sred[32 * 3];
  block_sum_d<3>(acc, sred);
  if (threadIdx.x == 0) {
    atomicAdd(k.sums + 0, acc[0]);
    atomicAdd(k.sums + 1, acc[1]);
    atomicAdd(k.sums + 2, acc[2]);
  }
}

// RtoD loss, vectorised (mode 0, W % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive pixels of a row.  Everything
// it needs -- the output row and the rows above / below, ground truth, sparse mask, three rows of the three image planes --
// is 14 independent float4 loads + 8 scalar edge loads issued back to back (the scalar kernel above issues ~40 dependent-
// address scalar loads per pixel and spends its time in load latency: 40 us for 30 MB, profiles/r02e_ncu_metrics.summary.txt).
// The per-pixel expressions are the ones of loss_kernel, operation for operation (gradients agree to fp32 rounding); the
// three loss SUMS are accumulated in fp32 over a thread's four pixels, then in fp64.
__global__ void __launch_bounds__(256) loss_rows_kernel(const LossK k, const int lg_tpr) {
  pdl_trigger();
  pdl_wait();
  const int W = k.W, H = k.H;
  const long long HW = (long long)H * W;
  const float c = 0.2f * (*k.maxabs);
  const int tpr = 1 << lg_tpr;                       // threads per row
  const int rpi = (int)blockDim.x >> lg_tpr;         // rows per CTA iteration
  const int tx = threadIdx.x & (tpr - 1), tr = threadIdx.x >> lg_tpr;
  const long long rows = (long long)k.N * H;
  double acc[3] = {0.0, 0.0, 0.0};
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = (long long)blockIdx.x * rpi + tr; row < rows; row += (long long)gridDim.x * rpi) {
    const long long n = row / H;
    const int y = (int)(row - n * H);
    const float* o = k.out + n * HW + (long long)y * W;
    const float* g = k.gt + n * HW + (long long)y * W;
    const float* sp = k.sparse ? k.sparse + n * k.sparse_stride + (long long)y * W : nullptr;
    const float* I = k.rgb + n * 3 * HW + (long long)y * W;
    const bool up = y >= 1, dn = y < H - 1;
    for (int x0 = 4 * tx; x0 < W; x0 += 4 * tpr) {
      // ---- loads (all independent)
      const float4 o1 = *reinterpret_cast<const float4*>(o + x0);
      const float4 o0 = up ? *reinterpret_cast<const float4*>(o - W + x0) : z4;
      const float4 o2 = dn ? *reinterpret_cast<const float4*>(o + W + x0) : z4;
      const float4 g4 = *reinterpret_cast<const float4*>(g + x0);
      const float4 s4 = sp ? *reinterpret_cast<const float4*>(sp + x0) : z4;
      float4 i0[3], i1[3], i2[3];
      float il[3], ir[3];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float* Ic = I + ch * HW;
        i1[ch] = *reinterpret_cast<const float4*>(Ic + x0);
        i0[ch] = up ? *reinterpret_cast<const float4*>(Ic - W + x0) : z4;
        i2[ch] = dn ? *reinterpret_cast<const float4*>(Ic + W + x0) : z4;
        il[ch] = x0 >= 1 ? Ic[x0 - 1] : 0.f;
        ir[ch] = x0 + 4 < W ? Ic[x0 + 4] : 0.f;
      }
      const float ol = x0 >= 1 ? o[x0 - 1] : 0.f, orr = x0 + 4 < W ? o[x0 + 4] : 0.f;
      // ---- edge weights: wx[j] = weight of the forward difference starting at x0 - 1 + j (j = 0..4), wy1 at (y, x), wy0 at (y-1, x)
      const float ov[6] = {ol, o1.x, o1.y, o1.z, o1.w, orr};
      float Iv[3][6];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        Iv[ch][0] = il[ch]; Iv[ch][1] = i1[ch].x; Iv[ch][2] = i1[ch].y; Iv[ch][3] = i1[ch].z; Iv[ch][4] = i1[ch].w; Iv[ch][5] = ir[ch];
      }
      float wx[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        float sd = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) sd += fabsf(Iv[ch][j] - Iv[ch][j + 1]);
        wx[j] = expf(-sd / 3.f);
      }
      const float* p0[3] = {&i0[0].x, &i0[1].x, &i0[2].x};
      const float* p2[3] = {&i2[0].x, &i2[1].x, &i2[2].x};
      const float* po0 = &o0.x;
      const float* po2 = &o2.x;
      const float* pg = &g4.x;
      const float* ps = &s4.x;
      float part0 = 0.f, part1 = 0.f, part2 = 0.f;
      float gr[4], dp[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int x = x0 + e;
        const float ovx = ov[e + 1];
        const float d = ovx - pg[e];
        const float a = fabsf(d);
        float wgt = 1.f;
        if (sp) {
          const bool crop = (y >= k.cy1 && y < k.cy2 && x >= k.cx1 && x < k.cx2);
          if (!crop) wgt = 0.1f;
          else if (!(ps[e] > -1.f)) wgt = 0.3f;
        }
        float val, dv;
        if (a > c) {
          val = (d * d + c * c) / (2.f * c);
          dv = d / c;
        } else {
          val = a;
          dv = sgn(d);
        }
        part0 += val * wgt;
        part2 += d * d;
        float grad = 3.f * k.inv_count * wgt * dv;
        float sm = 0.f, gs = 0.f;
        if (x < W - 1) {
          const float gx = ovx - ov[e + 2];
          const float w = wx[e + 1];
          sm += fabsf(gx) * w;
          gs += sgn(gx) * w;
        }
        if (x >= 1) {
          const float gx = ov[e] - ovx;
          gs -= sgn(gx) * wx[e];
        }
        if (dn) {
          float sd = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ch++) sd += fabsf(Iv[ch][e + 1] - p2[ch][e]);
          const float w = expf(-sd / 3.f);
          const float gy = ovx - po2[e];
          sm += fabsf(gy) * w;
          gs += sgn(gy) * w;
        }
        if (up) {
          float sd = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ch++) sd += fabsf(p0[ch][e] - Iv[ch][e + 1]);
          const float gy = po0[e] - ovx;
          gs -= sgn(gy) * expf(-sd / 3.f);
        }
        part1 += sm;
        grad += 0.1f * k.inv_count * gs;
        grad *= k.grad_scale;
        gr[e] = grad;
        dp[e] = grad * (1.f - ovx * ovx);
      }
      const long long off = row * W + x0;
      if (k.dout) *reinterpret_cast<float4*>(k.dout + off) = make_float4(gr[0], gr[1], gr[2], gr[3]);
      if (k.dpre) *reinterpret_cast<float4*>(k.dpre + off) = make_float4(dp[0], dp[1], dp[2], dp[3]);
      acc[0] += (double)part0;
      acc[1] += (double)part1;
      acc[2] += (double)part2;
    }
  }
  __shared__ double sred[32 * 3];
  block_sum_d<3>(acc, sred);
  if (threadIdx.x == 0) {
    atomicAdd(k.sums + 0, acc[0]);
    atomicAdd(k.sums + 1, acc[1]);
    atomicAdd(k.sums + 2, acc[2]);
  }
}

// ------------------------------------------------------------------------------- sum of squared diffs
__global__ void sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4,
                                  double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  double acc[1] = {0.0};
  float part = 0.f;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    part += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    if (++cnt == 64) {
      acc[0] += (double)part;
      part = 0.f;
      cnt = 0;
    }
  }
  acc[0] += (double)part;
  __shared__ double sred[32];
  block_sum_d<1>(acc, sred);
  if (threadIdx.x == 0) atomicAdd(out, acc[0]);
}

// ------------------------------------------------- guidance-loss gradient (opt-in)
// g[i] = coef * (a[i] - b[i]): gradient of coef/2 * sum (a - b)^2 w.r.t. a (one feature-MSE term of the latent loss)
__global__ void sqdiff_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, float coef,
                                   float* __restrict__ g) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(g)[i] = make_float4(coef * (x.x - y.x), coef * (x.y - y.y), coef * (x.z - y.z), coef * (x.w - y.w));
  }
}

// dpre[i] += scale * dout[i] * (1 - out[i]^2): chains a gradient w.r.t. the tanh output onto dL/d(pre-tanh)
__global__ void tanh_chain_add_kernel(const float* __restrict__ dout, const float* __restrict__ out, long long n, float scale,
                                      float* __restrict__ dpre) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float o = out[i];
    dpre[i] += scale * dout[i] * (1.f - o * o);
  }
}

// ---------------------------------------------------------------------------------------- depth metrics
// Arithmetic follows calculate_error.py operation by operation in fp32.
//   VAR 0  compute_errors        (KITTI / Eigen, :10-103)   out8 = abs_diff abs_rel sq_rel a1 a2 a3 rmse rmse_log
//   VAR 1  compute_errors_NYU    (:105-151)                 out8 = abs_diff abs_rel log10  a1 a2 a3 rmse rmse_log
//   VAR 2  compute_errors_Make3D (:153-182)                 out8 = abs_diff abs_rel log10  -  -  -  rmse -
struct MetricK {
  const float* gt_np;  // [B][H][W] sparse / raw ground truth (unused by VAR 1)
  const float* gt;     // [B][H][W] dense ground truth
  const float* pred;   // [B][H][W]
  int B, H, W;
  int crop, cy1, cy2, cx1, cx2;
  double* out;         // [8] += per-image metric / B
  long long* counts;   // [B][4] = n_valid, n(<1.25), n(<1.25^2), n(<1.25^3)
};

struct MinMax {
  float gmin, gmax, pmin, pmax, nmin, nmax;
};

__device__ __forceinline__ float norm_s(float v, float lo, float hi, float s) {
  return __fmul_rn(__fdiv_rn(__fsub_rn(v, lo), __fsub_rn(hi, lo)), s);
}

// validity of pixel i and the (gt, pred) pair the medians / metrics are computed on, before median scaling
template <int VAR>
__device__ __forceinline__ bool metric_pixel(const MetricK& m, const MinMax& mm, const float* __restrict__ gtn,
                                             const float* __restrict__ g, const float* __restrict__ p, int i, float& vg,
                                             float& vp) {
  const int y = i / m.W, x = i - y * m.W;
  bool valid;
  if (VAR == 0) {
    vg = norm_s(g[i], mm.gmin, mm.gmax, 80.f);                                          // :39,44
    const float n80 = __fmul_rn(__fdiv_rn(__fadd_rn(gtn[i], 1.0f), 2.0f), 80.f);       // :41,46
    valid = (n80 < 80.f) && (vg < 80.f) && (n80 > 1.f) && (vg > 1.f);                   // :78
    if (m.crop) valid = valid && (y >= m.cy1 && y < m.cy2 && x >= m.cx1 && x < m.cx2);
    if (valid) vp = norm_s(p[i], mm.pmin, mm.pmax, 80.f);
  } else if (VAR == 1) {
    vg = norm_s(g[i], mm.gmin, mm.gmax, 10.f);                                          // :122,125
    valid = (vg < 10.f) && (vg > 0.f);                                                  // :128
    if (m.crop) valid = valid && (y >= m.cy1 && y < m.cy2 && x >= m.cx1 && x < m.cx2);
    if (valid) vp = norm_s(p[i], mm.pmin, mm.pmax, 10.f);
  } else {
    const float g1 = __fdiv_rn(__fsub_rn(g[i], mm.gmin), __fsub_rn(mm.gmax, mm.gmin));     // :163
    const float n1 = __fdiv_rn(__fsub_rn(gtn[i], mm.nmin), __fsub_rn(mm.nmax, mm.nmin));   // :165
    const float g80 = __fmul_rn(g1, 80.f), n80 = __fmul_rn(n1, 80.f);
    valid = (n1 > 1e-2f) && (g1 > 1e-2f) && (n80 < 80.f) && (g80 < 80.f);               // :166,171
    if (valid) {
      vg = fminf(fmaxf(g80, 1e-2f), 80.f);                                              // :173
      vp = fminf(fmaxf(norm_s(p[i], mm.pmin, mm.pmax, 80.f), 1e-2f), 80.f);             // :174
    }
  }
  return valid;
}

// ---- staged evaluation: many CTAs per image, per-image state in a caller-owned workspace -------------------------
// Round 1 ran ONE 1024-thread CTA per image through 11 passes (min/max, count, 2 x 4 radix passes, metrics): 0.43 ms
// for 8 images of 128x416 and 3.8 ms at 384x1248 on 8 of 148 SMs (profiles/r02e_ncu_metrics.summary.txt: 12 GB/s).
// Now: 6 launches over (chunks, images) grids -- min/max; four radix passes that histogram the ground-truth AND the
// prediction values at once (the valid count is the total of the first histogram); the metric sums, finished by the last
// CTA of each image.  The per-pixel arithmetic (metric_pixel) is unchanged, medians are still exact radix selects and
// the delta-threshold counts integer-valued sums, so the COUNTS stay bit-exact; the fp64 sums are accumulated with
// atomics (order-dependent in the last bits only).
struct MetricImg {            // per image, zeroed by the host wrapper before the first kernel
  unsigned int mm[6];         // order-preserving encodings of gmin gmax pmin pmax nmin nmax (atomicMin / atomicMax)
  unsigned int done;          // CTAs of the final pass that have contributed
  unsigned int pad;
  unsigned int hist[4][2][256];   // [radix pass][0 = gt, 1 = pred][bin]
  double sums[9];
};

__device__ __forceinline__ unsigned int m_f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float m_ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void metrics_init_kernel(MetricImg* ws, int B) {
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * (int)(sizeof(MetricImg) / 4); i += gridDim.x * blockDim.x) {
    const int img = i / (int)(sizeof(MetricImg) / 4), w = i - img * (int)(sizeof(MetricImg) / 4);
    unsigned int v = 0u;
    if (w < 6) v = (w & 1) ? 0u : 0xffffffffu;            // running max starts at the smallest encoding, min at the largest
    reinterpret_cast<unsigned int*>(ws + img)[w] = v;
  }
}

template <int VAR>
__global__ void __launch_bounds__(256) metrics_minmax_kernel(const MetricK m, MetricImg* ws) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, HW = m.H * m.W;
  const float* g = m.gt + (long long)b * HW;
  const float* p = m.pred + (long long)b * HW;
  const float* gtn = m.gt_np ? m.gt_np + (long long)b * HW : nullptr;
  float v[6] = {INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float gv = g[i], pv = p[i];
    v[0] = fminf(v[0], gv); v[1] = fmaxf(v[1], gv);
    v[2] = fminf(v[2], pv); v[3] = fmaxf(v[3], pv);
    if (VAR == 2) {
      const float nv = gtn[i];
      v[4] = fminf(v[4], nv); v[5] = fmaxf(v[5], nv);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const float t = __shfl_xor_sync(0xffffffffu, v[j], o);
      v[j] = (j & 1) ? fmaxf(v[j], t) : fminf(v[j], t);
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < (VAR == 2 ? 6 : 4); j++) {
      if (j & 1) atomicMax(&ws[b].mm[j], m_f2ord(v[j]));
      else atomicMin(&ws[b].mm[j], m_f2ord(v[j]));
    }
  }
}

__device__ __forceinline__ MinMax metrics_load_mm(const MetricImg& w) {
  MinMax mm;
  mm.gmin = m_ord2f(w.mm[0]); mm.gmax = m_ord2f(w.mm[1]); mm.pmin = m_ord2f(w.mm[2]); mm.pmax = m_ord2f(w.mm[3]);
  mm.nmin = m_ord2f(w.mm[4]); mm.nmax = m_ord2f(w.mm[5]);
  return mm;
}

// radix-select state after `passes` completed passes, recomputed per CTA from the global histograms (<= 4 x 256 bins):
// value prefix, rank still to find, and the number of valid pixels (total of the first histogram).
// Called by a WHOLE WARP (all 32 lanes, every lane returns the same values): lane l holds bins 8l .. 8l+7, a shuffle scan
// finds the lane whose cumulative count crosses the rank, that lane finds the bin.  The first version walked the bins with
// one thread -- up to 255 dependent global loads per pass, which made it the critical path of every metrics kernel (34 us
// for the first radix pass over 5 MB, +5 .. 18 us per further pass, 72 us for the final one:
// profiles/r02r_ncu_metrics.summary.txt).  Integer arithmetic: same result as the sequential walk (bin 255 is never tested).
__device__ void metrics_select_state(const MetricImg& w, int passes, int which, unsigned int& prefix, long long& k,
                                     long long& nvalid) {
  const unsigned int full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += w.hist[0][0][lane * 8 + j];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(full, s, o);
  const long long n = (long long)s;
  nvalid = n;
  prefix = 0;
  long long kk = n > 0 ? (n - 1) / 2 : 0;         // lower median (torch.median), calculate_error.py:86 / :134 / :175
  for (int ps = 0; ps < passes; ps++) {
    const int shift = 24 - 8 * ps;
    unsigned int c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      c[j] = w.hist[ps][which][lane * 8 + j];
      tot += c[j];
    }
    unsigned int incl = tot;                       // inclusive prefix over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned int all = __shfl_sync(full, incl, 31), c255 = __shfl_sync(full, c[7], 31);
    const unsigned int hit = __ballot_sync(full, kk < (long long)incl);
    unsigned int bin;
    if (hit == 0u) {                               // rank beyond the histogram: the sequential walk ends at bin 255
      bin = 255u;
      kk -= (long long)(all - c255);
    } else {
      const int L = __ffs(hit) - 1;                // first lane whose cumulative count exceeds the rank
      long long myk = kk - (long long)(incl - tot);
      int j = 0;
      for (; j < 7; j++) {
        if (myk < (long long)c[j]) break;
        myk -= (long long)c[j];
      }
      bin = __shfl_sync(full, (unsigned int)(lane * 8 + j), L);
      kk = __shfl_sync(full, myk, L);
    }
    prefix |= bin << shift;
  }
  k = kk;
}

template <int VAR>
__global__ void __launch_bounds__(256) metrics_radix_kernel(const MetricK m, MetricImg* ws, const int pass) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, HW = m.H * m.W;
  const float* g = m.gt + (long long)b * HW;
  const float* p = m.pred + (long long)b * HW;
  const float* gtn = m.gt_np ? m.gt_np + (long long)b * HW : nullptr;
  __shared__ unsigned int hist[2][256];
  __shared__ unsigned int s_prefix[2];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&hist[0][0])[i] = 0;
  if (threadIdx.x < 64) {             // warp 0: ground truth, warp 1: prediction
    long long k, nv;
    unsigned int pf;
    metrics_select_state(ws[b], pass, threadIdx.x >> 5, pf, k, nv);
    if ((threadIdx.x & 31) == 0) s_prefix[threadIdx.x >> 5] = pf;
  }
  __syncthreads();
  const MinMax mm = metrics_load_mm(ws[b]);
  const int shift = 24 - 8 * pass;
  const unsigned int mask = pass ? (0xffffffffu << (shift + 8)) : 0u;
  const unsigned int pg = s_prefix[0], pp = s_prefix[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float vg, vp;
    if (!metric_pixel<VAR>(m, mm, gtn, g, p, i, vg, vp)) continue;
    const unsigned int bg = __float_as_uint(vg), bp = __float_as_uint(vp);
    if ((bg & mask) == pg) atomicAdd(&hist[0][(bg >> shift) & 255u], 1u);
    if ((bp & mask) == pp) atomicAdd(&hist[1][(bp >> shift) & 255u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const unsigned int c = (&hist[0][0])[i];
    if (c) atomicAdd(&ws[b].hist[pass][0][0] + i, c);
  }
}

template <int VAR>
__global__ void __launch_bounds__(256) metrics_final_kernel(const MetricK m, MetricImg* ws) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, HW = m.H * m.W;
  const float* g = m.gt + (long long)b * HW;
  const float* p = m.pred + (long long)b * HW;
  const float* gtn = m.gt_np ? m.gt_np + (long long)b * HW : nullptr;
  __shared__ double sred[32 * 9];
  __shared__ float s_med[2];
  __shared__ long long s_n;
  __shared__ int s_last;
  if (threadIdx.x < 64) {             // warp 0: ground truth, warp 1: prediction
    long long k, nv;
    unsigned int pf;
    metrics_select_state(ws[b], 4, threadIdx.x >> 5, pf, k, nv);
    if ((threadIdx.x & 31) == 0) {
      s_med[threadIdx.x >> 5] = __uint_as_float(pf);
      if (threadIdx.x == 0) s_n = nv;
    }
  }
  __syncthreads();
  const long long nvalid = s_n;
  if (nvalid == 0) {
    // the reference would produce NaNs here; callers treat n_valid == 0 as "no measurement"
    if (blockIdx.x == 0 && threadIdx.x == 0 && m.counts) { for (int j = 0; j < 4; j++) m.counts[b * 4 + j] = 0; }
    return;
  }
  const MinMax mm = metrics_load_mm(ws[b]);
  const float med_g = s_med[0], med_p = s_med[1];
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float vg, vp;
    if (!metric_pixel<VAR>(m, mm, gtn, g, p, i, vg, vp)) continue;
    vp = __fdiv_rn(__fmul_rn(vp, med_g), med_p);
    if (VAR == 0) vp = fminf(fmaxf(vp, 1.f), 80.f);       // :87
    if (VAR == 1) vp = fminf(fmaxf(vp, 1e-3f), 10.f);     // :135
    const float thr = fmaxf(__fdiv_rn(vg, vp), __fdiv_rn(vp, vg));
    const float d = __fsub_rn(vg, vp);
    acc[0] += (double)fabsf(d);
    acc[1] += (double)__fdiv_rn(fabsf(d), vg);
    if (VAR == 0) acc[2] += (double)__fdiv_rn(__fmul_rn(d, d), vg);
    else acc[2] += (double)fabsf(__fsub_rn(log10f(vg), log10f(vp)));
    acc[3] += (thr < 1.25f) ? 1.0 : 0.0;
    acc[4] += (thr < 1.5625f) ? 1.0 : 0.0;
    acc[5] += (thr < 1.953125f) ? 1.0 : 0.0;
    acc[6] += (double)__fmul_rn(d, d);
    const float ld = __fsub_rn(logf(vg), logf(vp));
    acc[7] += (double)__fmul_rn(ld, ld);
  }
  block_sum_d<9>(acc, sred);
  if (threadIdx.x == 0) {
    for (int j = 0; j < 8; j++) atomicAdd(&ws[b].sums[j], acc[j]);
    __threadfence();
    s_last = (atomicAdd(&ws[b].done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  double t[8];
  for (int j = 0; j < 8; j++) t[j] = atomicAdd(&ws[b].sums[j], 0.0);     // coherent read of the finished sums
  const double n = (double)nvalid, invB = 1.0 / (double)m.B;
  atomicAdd(m.out + 0, t[0] / n * invB);
  atomicAdd(m.out + 1, t[1] / n * invB);
  atomicAdd(m.out + 2, t[2] / n * invB);
  if (VAR != 2) {
    atomicAdd(m.out + 3, (double)((float)t[3] / (float)nvalid) * invB);
    atomicAdd(m.out + 4, (double)((float)t[4] / (float)nvalid) * invB);
    atomicAdd(m.out + 5, (double)((float)t[5] / (float)nvalid) * invB);
    atomicAdd(m.out + 7, sqrt(t[7] / n) * invB);
  }
  atomicAdd(m.out + 6, sqrt(t[6] / n) * invB);
  if (m.counts) {
    m.counts[b * 4 + 0] = nvalid;
    m.counts[b * 4 + 1] = (long long)t[3];
    m.counts[b * 4 + 2] = (long long)t[4];
    m.counts[b * 4 + 3] = (long long)t[5];
  }
}

template <int VAR>
static int metrics_launch(const MetricK& m, MetricImg* ws, cudaStream_t st) {
  const int HW = m.H * m.W;
  int chunks = (HW + 4095) / 4096;                       // >= 16 pixels per thread
  const int cap = (device_sm_count() * 4 + m.B - 1) / m.B;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  const dim3 grid(chunks, m.B);
  GDN_CUDA_CHECK(launch_pdl(metrics_init_kernel, dim3((m.B * (int)(sizeof(MetricImg) / 4) + 255) / 256), dim3(256), 0, st, 1, ws, m.B));
  GDN_CUDA_CHECK(launch_pdl(metrics_minmax_kernel<VAR>, dim3(grid), dim3(256), 0, st, 1, m, ws));
  for (int pass = 0; pass < 4; pass++) GDN_CUDA_CHECK(launch_pdl(metrics_radix_kernel<VAR>, dim3(grid), dim3(256), 0, st, 1, m, ws, pass));
  GDN_CUDA_CHECK(launch_pdl(metrics_final_kernel<VAR>, dim3(grid), dim3(256), 0, st, 1, m, ws));
  return GDN_OK;
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam semantics with coupled L2 weight decay (GDN_main.py:157,173):
//   g += wd*p ; m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float step_size, float inv_sqrt_bc2, float b1, float b2,
                            float eps, float wd, float grad_scale, const float* __restrict__ dyn, double b1d,
                            double b2d) {
  pdl_trigger();
  pdl_wait();
  if (dyn) {  // learning rate and step count live in device memory so that a captured CUDA graph stays valid
    const double t = (double)dyn[1];
    step_size = (float)((double)dyn[0] / (1.0 - pow(b1d, t)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(b2d, t)));
  }
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp);
    const float* G = reinterpret_cast<const float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm);
    float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float gr = G[j] * grad_scale + wd * P[j];
      M[j] = b1 * M[j] + (1.f - b1) * gr;
      V[j] = b2 * V[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[j]) * inv_sqrt_bc2 + eps;
      P[j] -= step_size * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * grad_scale + wd * p[i];
    const float mn = b1 * m[i] + (1.f - b1) * gr;
    const float vn = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mn;
    v[i] = vn;
    p[i] -= step_size * (mn / (sqrtf(vn) * inv_sqrt_bc2 + eps));
  }
}

}  // namespace gdn

using namespace gdn;
#define GDN_API extern "C" __attribute__((visibility("default")))

static int lm_grid(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

GDN_API int gdn_absdiff_max(const float* a, const float* b, int64_t n, float* out_max, gdn_stream stream) {
  if (!a || !b || !out_max) return fail(GDN_INVALID_DESC, "gdn_absdiff_max: null pointer");
  GDN_CUDA_CHECK(launch_pdl(absdiff_max_kernel, dim3(lm_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream, 1, a, b, n, reinterpret_cast<unsigned int*>(out_max)));
  GDN_LAUNCH_CHECK("absdiff_max_kernel");
  return GDN_OK;
}

GDN_API int gdn_loss(const gdn_loss_desc* d, gdn_stream stream) {
  if (!d || !d->out || !d->gt || !d->maxabs || !d->sums) return fail(GDN_INVALID_DESC, "gdn_loss: null pointer");
  if (d->mode == 0 && !d->rgb) return fail(GDN_INVALID_DESC, "gdn_loss: RtoD mode needs the rgb input");
  LossK k{};
  k.out = d->out; k.gt = d->gt; k.sparse = d->sparse; k.sparse_stride = d->sparse_stride; k.rgb = d->rgb;
  k.N = d->n; k.H = d->h; k.W = d->w;
  k.maxabs = d->maxabs;
  k.mode = d->mode;
  // Garg ECCV16 crop of the training loss (trainer.py:644-645)
  k.cy1 = (int)(0.40810811 * d->h); k.cy2 = (int)(0.99189189 * d->h);
  k.cx1 = (int)(0.03594771 * d->w); k.cx2 = (int)(0.96405229 * d->w);
  k.inv_count = 1.0f / (float)((long long)d->n * d->h * d->w);
  k.sums = d->sums; k.dout = d->dout; k.dpre = d->dpre;
  k.grad_scale = d->grad_scale;
  {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = d->mode == 0 && d->w % 4 == 0 && d->w >= 8 && al16(d->out) && al16(d->gt) && al16(d->rgb) &&
                     (!d->sparse || (al16(d->sparse) && d->sparse_stride % 4 == 0)) && (!d->dout || al16(d->dout)) &&
                     (!d->dpre || al16(d->dpre));
    if (vec) {
      int lg = 1;                                  // threads per row: power of two >= w / 4, at most the CTA
      while ((1 << lg) < d->w / 4 && lg < 8) lg++;
      const long long rows = (long long)d->n * d->h;
      const int rpi = 256 >> lg;
      long long ctas = (rows + rpi - 1) / rpi;
      const long long cap = (long long)device_sm_count() * 4;
      if (ctas > cap) ctas = cap;
      GDN_CUDA_CHECK(launch_pdl(loss_rows_kernel, dim3((unsigned)ctas), dim3(256), 0, (cudaStream_t)stream, 1, k, lg));
      GDN_LAUNCH_CHECK("loss_rows_kernel");
      return GDN_OK;
    }
  }
  GDN_CUDA_CHECK(launch_pdl(loss_kernel, dim3(lm_grid((long long)d->n * d->h * d->w, 256)), dim3(256), 0, (cudaStream_t)stream, 1, k));
  GDN_LAUNCH_CHECK("loss_kernel");
  return GDN_OK;
}

GDN_API int gdn_sqdiff_sum(const float* a, const float* b, int64_t n, double* out, gdn_stream stream) {
  if (!a || !b || !out || (n & 3)) return fail(GDN_INVALID_DESC, "gdn_sqdiff_sum: bad arguments (n must be a multiple of 4)");
  GDN_CUDA_CHECK(launch_pdl(sqdiff_sum_kernel, dim3(lm_grid(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, 1, a, b, n / 4, out));
  GDN_LAUNCH_CHECK("sqdiff_sum_kernel");
  return GDN_OK;
}

GDN_API int gdn_sqdiff_grad(const float* a, const float* b, int64_t n, float coef, float* grad, gdn_stream stream) {
  if (!a || !b || !grad || (n & 3)) return fail(GDN_INVALID_DESC, "gdn_sqdiff_grad: bad arguments (n must be a multiple of 4)");
  GDN_CUDA_CHECK(launch_pdl(sqdiff_grad_kernel, dim3(lm_grid(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, 1, a, b, n / 4, coef, grad));
  GDN_LAUNCH_CHECK("sqdiff_grad_kernel");
  return GDN_OK;
}

GDN_API int gdn_tanh_chain_add(const float* dout, const float* out, int64_t n, float scale, float* dpre, gdn_stream stream) {
  if (!dout || !out || !dpre || n < 0) return fail(GDN_INVALID_DESC, "gdn_tanh_chain_add: bad arguments");
  GDN_CUDA_CHECK(launch_pdl(tanh_chain_add_kernel, dim3(lm_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream, 1, dout, out, n, scale, dpre));
  GDN_LAUNCH_CHECK("tanh_chain_add_kernel");
  return GDN_OK;
}

GDN_API size_t gdn_depth_metrics_workspace_bytes(int b) { return b > 0 ? (size_t)b * sizeof(MetricImg) : 0; }

GDN_API int gdn_depth_metrics(int variant, const float* gt_np, const float* gt, const float* pred, int b, int h, int w,
                              int crop, double* out8, int64_t* counts, void* workspace, size_t workspace_bytes,
                              gdn_stream stream) {
  if (variant < 0 || variant > 2) return fail(GDN_INVALID_DESC, "gdn_depth_metrics: variant %d", variant);
  if ((!gt_np && variant != GDN_METRICS_NYU) || !gt || !pred || !out8 || b < 1)
    return fail(GDN_INVALID_DESC, "gdn_depth_metrics: bad arguments");
  if (!workspace || workspace_bytes < gdn_depth_metrics_workspace_bytes(b) || (reinterpret_cast<uintptr_t>(workspace) & 7))
    return fail(GDN_WORKSPACE_TOO_SMALL, "gdn_depth_metrics: needs %zu bytes of 8-byte aligned workspace",
                gdn_depth_metrics_workspace_bytes(b));
  MetricK m{};
  m.gt_np = gt_np; m.gt = gt; m.pred = pred;
  m.B = b; m.H = h; m.W = w; m.crop = (variant == GDN_METRICS_MAKE3D) ? 0 : crop;
  if (variant == GDN_METRICS_KITTI) {
    // crop used by Godard CVPR17 (calculate_error.py:28-29)
    m.cy1 = (int)(0.3324324 * h); m.cy2 = (int)(0.91351351 * h);
  } else {
    m.cy1 = (int)(0.0359477 * h); m.cy2 = (int)(0.96405229 * h);   // calculate_error.py:111
  }
  m.cx1 = (int)(0.0359477 * w); m.cx2 = (int)(0.96405229 * w);
  m.out = out8;
  m.counts = reinterpret_cast<long long*>(counts);
  MetricImg* ws = reinterpret_cast<MetricImg*>(workspace);
  int lrc;
  if (variant == GDN_METRICS_KITTI) lrc = metrics_launch<0>(m, ws, (cudaStream_t)stream);
  else if (variant == GDN_METRICS_NYU) lrc = metrics_launch<1>(m, ws, (cudaStream_t)stream);
  else lrc = metrics_launch<2>(m, ws, (cudaStream_t)stream);
  if (lrc) return lrc;
  GDN_LAUNCH_CHECK("depth metrics kernels");
  return GDN_OK;
}

GDN_API int gdn_eigen_metrics(const float* gt_np, const float* gt, const float* pred, int b, int h, int w, int crop,
                              double* out8, int64_t* counts, void* workspace, size_t workspace_bytes, gdn_stream stream) {
  return gdn_depth_metrics(GDN_METRICS_KITTI, gt_np, gt, pred, b, h, w, crop, out8, counts, workspace, workspace_bytes, stream);
}

GDN_API int gdn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float grad_scale, gdn_stream stream) {
  if (!p || !g || !m || !v || n < 0 || step < 1) return fail(GDN_INVALID_DESC, "gdn_adam_step: bad arguments");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(GDN_INVALID_DESC, "gdn_adam_step: buffers must be 16-byte aligned");
  if (n == 0) return GDN_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  GDN_CUDA_CHECK(launch_pdl(adam_kernel, dim3(lm_grid(n / 4 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, 1, p, g, m, v, n, step_size, inv_sqrt_bc2, beta1, beta2, eps, weight_decay, grad_scale, nullptr, 0.0, 0.0));
  GDN_LAUNCH_CHECK("adam_kernel");
  return GDN_OK;
}

__global__ void adam_tick_kernel(float* dyn) {
  pdl_trigger();
  pdl_wait(); dyn[1] += 1.0f; }

GDN_API int gdn_adam_step_dyn(float* p, const float* g, float* m, float* v, int64_t n, float* dyn, double beta1,
                              double beta2, float eps, float weight_decay, float grad_scale, gdn_stream stream) {
  if (!p || !g || !m || !v || !dyn || n < 0) return fail(GDN_INVALID_DESC, "gdn_adam_step_dyn: bad arguments");
  if (n == 0) return GDN_OK;
  GDN_CUDA_CHECK(launch_pdl(adam_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, 1, dyn));
  GDN_CUDA_CHECK(launch_pdl(adam_kernel, dim3(lm_grid(n / 4 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, 1, p, g, m, v, n, 0.f, 0.f, (float)beta1, (float)beta2, eps, weight_decay, grad_scale, dyn, beta1, beta2));
  GDN_LAUNCH_CHECK("adam_kernel");
  return GDN_OK;
}
