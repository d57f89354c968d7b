// Adjoint of a convolution's input transform (x2 bilinear upsample / zero dilation, then reflection or zero border),
// row-per-CTA version with the column candidates tabulated once per CTA ("fold_rows2").
//
// The pieces are __host__ __device__ so that the CPU test-suite can run the SAME index arithmetic thread by thread on
// the host (tests/host/fold_sim.cu, a test-only shared object; nothing here is a CPU path of the product).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define GDN_HD __host__ __device__ __forceinline__
#else
#define GDN_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define GDN_LDG4(p) __ldg(reinterpret_cast<const float4*>(p))
#else
#define GDN_LDG4(p) (*reinterpret_cast<const float4*>(p))
#endif

namespace gdn {

struct FoldK {
  const float* dpad;   // [N][OH+2P][OW+2P][ctot]: gradient w.r.t. the conv's input buffer (fp32, or bf16 when dpad_bf16)
  int ctot, c_off;
  int N, H, W, C;      // source activation extent
  int P, reflect, up, dilate;
  float* dact;         // fp32 [N][H][W][C]
  int accumulate;
  int dpad_bf16;       // dpad holds bf16 (an input-gradient convolution wrote it through its bf16 output)
};

struct FoldV4 { float x, y, z, w; };

// four consecutive channels at ELEMENT offset `off` of dpad (bf16 -> fp32 is a 16-bit shift: same code on host and device)
GDN_HD FoldV4 fold_load4(const FoldK& f, size_t off) {
  FoldV4 r;
  if (f.dpad_bf16) {
    const uint64_t u = *reinterpret_cast<const uint64_t*>(reinterpret_cast<const uint16_t*>(f.dpad) + off);
    const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
    const uint32_t b0 = lo << 16, b1 = lo & 0xffff0000u, b2 = hi << 16, b3 = hi & 0xffff0000u;
    r.x = *reinterpret_cast<const float*>(&b0); r.y = *reinterpret_cast<const float*>(&b1);
    r.z = *reinterpret_cast<const float*>(&b2); r.w = *reinterpret_cast<const float*>(&b3);
  } else {
    const float4 v = GDN_LDG4(f.dpad + off);
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
  }
  return r;
}

// source coordinates / weight of hi-res index o for x2 bilinear upsampling (mode 1: align_corners=False, 2: True)
GDN_HD void up_coord(int o, int in, int mode, int& i0, int& i1, float& w1) {
  float src;
  if (mode == 1) {
    src = (o + 0.5f) * 0.5f - 0.5f;
    if (src < 0.f) src = 0.f;
  } else {
    src = in > 1 ? o * (float)(in - 1) / (float)(2 * in - 1) : 0.f;
  }
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  w1 = src - (float)i0;
}

// buffer positions (border P) that hold a copy of interior index Y: itself and its reflection images
GDN_HD int mirror_set(int Y, int OH, int P, int reflect, int* out) {
  int n = 0;
  out[n++] = Y + P;
  if (reflect) {
    if (Y >= 1 && Y <= P) out[n++] = P - Y;
    if (Y >= OH - 1 - P && Y <= OH - 2) out[n++] = P + 2 * (OH - 1) - Y;
  }
  return n;
}

constexpr int kFoldColInts = 12;   // per source column: 4 candidate hi-res columns x up to 3 reflection images (-1 = unused)

// candidate buffer columns feeding source column x, and the weight of each candidate
GDN_HD void fold_col_entry(const FoldK& f, int OW, int x, int* pc, float* qw) {
  for (int k = 0; k < 4; k++) {
    qw[k] = 0.f;
    pc[3 * k] = pc[3 * k + 1] = pc[3 * k + 2] = -1;
    int X;
    if (f.up) X = 2 * x - 1 + k;
    else if (k == 0) X = f.dilate ? 2 * x : x;
    else continue;
    if (X < 0 || X >= OW) continue;
    float w = 1.f;
    if (f.up) {
      int a0, a1;
      float w1;
      up_coord(X, f.W, f.up, a0, a1, w1);
      w = (a0 == x ? 1.f - w1 : 0.f) + (a1 == x ? w1 : 0.f);
      if (w == 0.f) continue;
    }
    qw[k] = w;
    pc[3 * k] = X + f.P;
    if (f.reflect) {
      if (X >= 1 && X <= f.P) pc[3 * k + 1] = f.P - X;
      if (X >= OW - 1 - f.P && X <= OW - 2) pc[3 * k + 2] = f.P + 2 * (OW - 1) - X;
    }
  }
}

// buffer rows feeding source row y (with reflection images) and their weights; returns the count (<= 12)
GDN_HD int fold_row_entry(const FoldK& f, int OH, int y, int* prow, float* pw) {
  int np = 0;
  const int ylo = f.up ? 2 * y - 1 : (f.dilate ? 2 * y : y), yhi = f.up ? 2 * y + 2 : ylo;
  for (int Y = ylo; Y <= yhi; Y++) {
    if (Y < 0 || Y >= OH) continue;
    float w = 1.f;
    if (f.up) {
      int a0, a1;
      float w1;
      up_coord(Y, f.H, f.up, a0, a1, w1);
      w = (a0 == y ? 1.f - w1 : 0.f) + (a1 == y ? w1 : 0.f);
      if (w == 0.f) continue;
    }
    int my[3];
    const int cy = mirror_set(Y, OH, f.P, f.reflect, my);
    for (int p = 0; p < cy; p++) { prow[np] = my[p]; pw[np++] = w; }
  }
  return np;
}

// one (x, 4-channel group) item of source row `row`: same accumulation order as the first version of the kernel
GDN_HD void fold_item(const FoldK& f, int lg_cg, int Hq, int Wq, int row, int it, int np, const int* prow, const float* pw,
                      const int* pcs, const float* qws) {
  const int cgm = (1 << lg_cg) - 1;
  const int x = it >> lg_cg, c4 = (it & cgm) * 4;
  const int n = row / f.H;
  const size_t img = (size_t)n * Hq;
  const int* pc = pcs + (size_t)x * kFoldColInts;
  const float* qw = qws + (size_t)x * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < np; p++) {
    const size_t rowo = ((img + prow[p]) * Wq) * f.ctot + f.c_off + c4;
    const float wr = pw[p];
    for (int k = 0; k < 4; k++) {
      const float w = wr * qw[k];
      for (int m = 0; m < 3; m++) {
        const int col = pc[3 * k + m];
        if (col < 0) continue;
        const FoldV4 v = fold_load4(f, rowo + (size_t)col * f.ctot);
#if defined(__CUDA_ARCH__)
        acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]); acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
#else
        acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
#endif
      }
    }
  }
  float* o = f.dact + ((size_t)row * f.W + x) * f.C + c4;
  if (f.accumulate) {
    const float4 v = *reinterpret_cast<const float4*>(o);
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
  *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

}  // namespace gdn
