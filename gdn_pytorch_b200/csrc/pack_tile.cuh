// Weight re-pack, wide-tile version ("pack v2"): fp32 parameter tensor -> bf16 [tap][a][b] operand layout.
// A CTA moves a 16 (a) x TB (b) tile for ALL taps through shared memory; TB = 64 / 32 / 16 by tap count, so that the
// packed side is written in runs of up to 128 bytes (b is the contiguous index of the operand layout) and the CTA
// count of the big 3x3 / 512-channel tensors drops 4x against the first version (16 x 16 tiles, 32-byte runs,
// latency-bound at ~1 TB/s).  The two phases are __host__ __device__ so that the CPU test-suite can run the SAME index
// arithmetic thread by thread on the host (tests/host/pack_sim.cu, test-only; not a CPU path of the product).
#pragma once
#include <cuda_bf16.h>
#include <cstddef>

#if defined(__CUDACC__)
#define GDN_PHD __host__ __device__ __forceinline__
#else
#define GDN_PHD inline
#endif

namespace gdn {

struct PackK {
  int kh, kw, A, B, Apad, Bpad;
  long long sa, sb, sr, ss;
  int flip, col_c;
};

constexpr int kPackTA = 16;

GDN_PHD int pack_tb_for_taps(int T) { return T <= 9 ? 64 : (T <= 25 ? 32 : 16); }
GDN_PHD int pack_smem_stride(int T) { return (T & 1) ? T : T + 1; }       // odd: conflict-free reads across b
GDN_PHD size_t pack_smem_bytes(int T) { return (size_t)kPackTA * pack_tb_for_taps(T) * pack_smem_stride(T) * sizeof(float); }

// phase 1: parameter tensor -> s_tile[(al * TB + bl) * ST + tap]; address of (a, b, tap) = a*sa + b*sb + tap
// (one of sa / sb equals T: [a][b][taps] for Conv2d, [b][a][taps] for ConvTranspose2d / input-gradient views).
// TT > 0: the tap count as a compile-time constant (divisions by it become multiply-shifts: the run-time divisions of
// the first version made this kernel instruction-bound at ~1 TB/s, profiles/r02e_ncu_elem.summary.txt); full interior
// tiles whose rows are 16-byte aligned are read as float4.
template <int TT>
GDN_PHD void pack_v2_phase1_t(const float* w, const PackK& k, int a0, int b0, int TB, float* s_tile, int tid, int nthr) {
  const int T = TT > 0 ? TT : k.kh * k.kw, ST = pack_smem_stride(T);
  const bool b_inner = (k.sb == T);
  const int n_outer = b_inner ? kPackTA : TB, n_inner = b_inner ? TB : kPackTA;
  const int run = n_inner * T;
  const long long s_outer = b_inner ? k.sa : k.sb;
  const int o0 = b_inner ? a0 : b0, i0 = b_inner ? b0 : a0;           // first outer / inner index of the tile
  const int n_o = b_inner ? k.A : k.B, n_i = b_inner ? k.B : k.A;     // valid extents
  const float* base = w + (long long)o0 * s_outer + (long long)i0 * T;
  const bool vec = (run % 4 == 0) && (i0 + n_inner <= n_i) && (o0 + n_outer <= n_o) && (s_outer % 4 == 0) &&
                   ((reinterpret_cast<size_t>(base) & 15) == 0);
  if (vec) {
    const int q_per = run / 4;
    for (int i = tid; i < n_outer * q_per; i += nthr) {
      const int outer = i / q_per, q = i - outer * q_per;
      const float* p = base + (long long)outer * s_outer + 4 * q;
#if defined(__CUDA_ARCH__)
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(p));
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#else
      const float v[4] = {p[0], p[1], p[2], p[3]};
#endif
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int rem = 4 * q + e;
        const int inner = rem / T, tap = rem - inner * T;
        const int al = b_inner ? outer : inner, bl = b_inner ? inner : outer;
        s_tile[(al * TB + bl) * ST + tap] = v[e];
      }
    }
    return;
  }
  for (int i = tid; i < n_outer * run; i += nthr) {
    const int outer = i / run, rem = i - outer * run;
    const int inner = rem / T, tap = rem - inner * T;
    const int al = b_inner ? outer : inner, bl = b_inner ? inner : outer;
    const int a = a0 + al, b = b0 + bl;
    float v = 0.f;
    if (a < k.A && b < k.B) v = w[(long long)a * k.sa + (long long)b * k.sb + tap];
    s_tile[(al * TB + bl) * ST + tap] = v;
  }
}

GDN_PHD void pack_v2_phase1(const float* w, const PackK& k, int a0, int b0, int TB, float* s_tile, int tid, int nthr) {
  switch (k.kh * k.kw) {
    case 1: pack_v2_phase1_t<1>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    case 9: pack_v2_phase1_t<9>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    case 16: pack_v2_phase1_t<16>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    case 25: pack_v2_phase1_t<25>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    case 49: pack_v2_phase1_t<49>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    case 81: pack_v2_phase1_t<81>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
    default: pack_v2_phase1_t<0>(w, k, a0, b0, TB, s_tile, tid, nthr); break;
  }
}

// phase 2: s_tile -> out[(t * Apad + a) * Bpad + b] (taps flipped when flip, rows scaled by scale_a, padding = 0).
// One thread = two consecutive b of one (tap, a): a warp writes 128 contiguous bytes; TB and the a-tile are powers of two,
// so the index decode is shifts and masks.
GDN_PHD void pack_v2_phase2(const float* scale_a, __nv_bfloat16* out, const PackK& k, int a0, int b0, int TB,
                            const float* s_tile, int tid, int nthr) {
  const int T = k.kh * k.kw, ST = pack_smem_stride(T);
  const int lg_hb = (TB == 64) ? 5 : ((TB == 32) ? 4 : 3);            // log2(TB / 2)
  const int hb_mask = (1 << lg_hb) - 1;
  const bool pairs = (k.Bpad % 2 == 0);
  if (pairs) {
    for (int i = tid; i < (T * kPackTA) << lg_hb; i += nthr) {
      const int bl = (i & hb_mask) * 2, al = (i >> lg_hb) & (kPackTA - 1), t = i >> (lg_hb + 4);
      const int a = a0 + al, b = b0 + bl;
      if (a >= k.Apad || b >= k.Bpad) continue;
      const int tap = k.flip ? T - 1 - t : t;
      float v0 = s_tile[(al * TB + bl) * ST + tap], v1 = s_tile[(al * TB + bl + 1) * ST + tap];
      if (scale_a && a < k.A) { const float sc = scale_a[a]; v0 *= sc; v1 *= sc; }
      __nv_bfloat16* o = out + ((long long)t * k.Apad + a) * k.Bpad + b;
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v0, v1);
#else
      o[0] = __float2bfloat16(v0);
      o[1] = __float2bfloat16(v1);
#endif
    }
    return;
  }
  for (int i = tid; i < T * kPackTA * TB; i += nthr) {
    const int bl = i % TB, al = (i / TB) % kPackTA, t = i / (TB * kPackTA);
    const int a = a0 + al, b = b0 + bl;
    if (a >= k.Apad || b >= k.Bpad) continue;
    const int tap = k.flip ? T - 1 - t : t;
    float v = s_tile[(al * TB + bl) * ST + tap];
    if (scale_a && a < k.A) v *= scale_a[a];
    out[((long long)t * k.Apad + a) * k.Bpad + b] = __float2bfloat16(v);
  }
}

}  // namespace gdn
