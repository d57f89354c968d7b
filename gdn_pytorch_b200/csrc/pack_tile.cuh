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
// (one of sa / sb equals T: [a][b][taps] for Conv2d, [b][a][taps] for ConvTranspose2d / input-gradient views)
GDN_PHD void pack_v2_phase1(const float* w, const PackK& k, int a0, int b0, int TB, float* s_tile, int tid, int nthr) {
  const int T = k.kh * k.kw, ST = pack_smem_stride(T);
  const bool b_inner = (k.sb == T);
  const int n_outer = b_inner ? kPackTA : TB, n_inner = b_inner ? TB : kPackTA;
  const int run = n_inner * T;
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

// phase 2: s_tile -> out[(t * Apad + a) * Bpad + b] (taps flipped when flip, rows scaled by scale_a, padding = 0)
GDN_PHD void pack_v2_phase2(const float* scale_a, __nv_bfloat16* out, const PackK& k, int a0, int b0, int TB,
                            const float* s_tile, int tid, int nthr) {
  const int T = k.kh * k.kw, ST = pack_smem_stride(T);
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
