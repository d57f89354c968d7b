// Weight-gradient implicit GEMM for sm_100a (tcgen05, MN-major operands straight out of NHWC tiles).
//
//   dW[tap][ci][co] (fp32, +=)  =  sum over pixels  X[pix (+) tap][ci] * dY[pix][co]
//
// GEMM view per CTA: D[M = 128 rows, N = WN output channels] with the pixel index as K.  The 128 rows are a
// "unit": TWO (tap, 64-channel chunk) slices of the forward input stacked along M, so the tensor core always runs
// at M = 128 even for 64-channel layers.  WN = 64 / 128 / 256 (the widest that divides Cout: an N = 64 MMA is
// shared-memory-bandwidth bound at 48 cycles, N >= 128 runs at the tensor pipe's rate).  Each CTA owns up to
// 512 / WN units (512 TMEM columns), one WN-wide block of output channels and a contiguous range of pixel tiles
// (split-K); it accumulates in TMEM over its whole range and flushes once with vectorised fp32 reductions.
//
// Operand staging (both operands are MN-major: the channel index is contiguous in NHWC, pixels are K):
//   HALO   (stride 1): the X halo of a 16 x 8J pixel tile is brought in by ONE TMA box per pixel tile; every unit
//          reads its two taps as shifted windows (descriptor start address / LBO), exactly like the forward kernel.
//   TAPBOX (stride 2, 1x1): one TMA box per (unit half, pixel tile).
// dY tiles are plain [128 pixels][64 co] boxes.
#include <cuda_bf16.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace gdn {

constexpr int kWgThreads = 8 * 32;
constexpr int kWgWarpX = 4, kWgWarpMMA = 5, kWgWarpY = 6, kWgWarpMMA2 = 7;  // two MMA issuers split the units (even / odd)
constexpr int kMaxUnits = 64;
constexpr int kWgStagesMax = 8;

struct WgUnit {
  // half A / half B: tap row, tap col, 64-channel chunk (global chunk index over src0|src1); sB < 0 -> no second half
  int16_t rA, sA, cA, rB, sB, cB;
};

struct WgK {
  int n_img, out_h, out_w;
  int mode, J;
  int tw_log2, th_log2, nb;
  int tiles_x, tiles_y, groups_img, total_pt;  // pixel tiles
  int stride;
  int chunks0, c0_total, c1_total;
  int off_y, off_x, off_y1, off_x1;            // buffer coordinates (include source pad)
  int halo_w;
  uint32_t x_bytes, y_bytes, x_tx;             // per-stage strides; x_tx = bytes one X stage actually receives
  int nstage;
  int n_units, units_per_cta, n_groups;        // unit groups
  int ci_chunks;                               // HALO: separate CTAs per 64-ci chunk; TAPBOX: 1 (chunk is in the unit)
  int co_blocks, splits, total_work;
  int cin_total, cout_pad;
  int kw;
  float* dw;
  size_t det_slab;   // 0: splits add to dw with red.global.add; else split s stores to dw + s*det_slab (floats)
  WgUnit units[kMaxUnits];
};

struct WRing {
  int i = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void next(int n) {
    if (++i == n) {
      i = 0;
      ph ^= 1;
    }
  }
};

template <int WN>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ WgK p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + (size_t)p.nstage * p.x_bytes;

  __shared__ uint64_t x_full[kWgStagesMax], x_empty[kWgStagesMax], y_full[kWgStagesMax], y_empty[kWgStagesMax];
  __shared__ uint64_t acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool halo = (p.mode == GDN_CONV_HALO);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStagesMax; i++) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 2);   // both issuers observe (wait + commit) every X stage, see the MMA loop
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 2);
    }
    mbar_init(&acc_full, 2);
    mbar_init(&acc_empty, 4);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == kWgWarpMMA) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == kWgWarpX && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
  }
  if (warp == kWgWarpY && lane == 0) tma_prefetch_desc(&tmY);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // work unit -> (unit group, ci chunk, co block, split)
  auto decode_work = [&](int w, int& g, int& cic, int& cob, int& pt0, int& pt1) {
    const int sp = w % p.splits;
    w /= p.splits;
    cob = w % p.co_blocks;
    w /= p.co_blocks;
    cic = w % p.ci_chunks;
    g = w / p.ci_chunks;
    pt0 = (int)((long long)p.total_pt * sp / p.splits);
    pt1 = (int)((long long)p.total_pt * (sp + 1) / p.splits);
  };
  auto decode_pt = [&](int t, int& tx, int& ty, int& img) {
    tx = t % p.tiles_x;
    t /= p.tiles_x;
    ty = t % p.tiles_y;
    img = t / p.tiles_y;
  };
  const int tile_h = halo ? 16 : (1 << p.th_log2);
  const int tile_w = halo ? 8 * p.J : (1 << p.tw_log2);

  if (warp == kWgWarpX) {
    // ------------------------------------------------------------ X producer
    if (lane == 0) {
      WRing rx;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int g, cic, cob, pt0, pt1;
        decode_work(w, g, cic, cob, pt0, pt1);
        const int u0 = g * p.units_per_cta;
        const int u1 = min(u0 + p.units_per_cta, p.n_units);
        for (int t = pt0; t < pt1; t++) {
          int tx, ty, img;
          decode_pt(t, tx, ty, img);
          const int oy0 = ty * tile_h, ox0 = tx * tile_w, n0 = img * p.nb;
          if (halo) {
            const bool s1 = cic >= p.chunks0;
            mbar_wait(&x_empty[rx.i], rx.ph ^ 1);
            mbar_expect_tx(&x_full[rx.i], p.x_tx);
            tma_load_4d(s1 ? &tmX1 : &tmX0, &x_full[rx.i], sX + (size_t)rx.i * p.x_bytes,
                        (s1 ? cic - p.chunks0 : cic) * 64, ox0 + (s1 ? p.off_x1 : p.off_x),
                        oy0 + (s1 ? p.off_y1 : p.off_y), img);
            rx.next(p.nstage);
          } else {
            for (int u = u0; u < u1; u++) {
              const WgUnit un = p.units[u];
              mbar_wait(&x_empty[rx.i], rx.ph ^ 1);
              mbar_expect_tx(&x_full[rx.i], p.x_tx);
              uint8_t* dst = sX + (size_t)rx.i * p.x_bytes;
#pragma unroll
              for (int h = 0; h < 2; h++) {
                // a missing second half re-reads the first one (its rows are discarded by the epilogue)
                const int r = (h == 0 || un.sB < 0) ? un.rA : un.rB;
                const int s = (h == 0 || un.sB < 0) ? un.sA : un.sB;
                const int c = (h == 0 || un.sB < 0) ? un.cA : un.cB;
                const bool s1 = c >= p.chunks0;
                const CUtensorMap* tm = s1 ? &tmX1 : &tmX0;
                const int cc = s1 ? c - p.chunks0 : c;
                const int offy = s1 ? p.off_y1 : p.off_y, offx = s1 ? p.off_x1 : p.off_x;
                if (p.stride == 1) {
                  tma_load_4d(tm, &x_full[rx.i], dst + h * 16384, cc * 64, ox0 + s + offx, oy0 + r + offy, n0);
                } else {
                  const int bx = s + offx, by = r + offy;
                  const int ctot = s1 ? p.c1_total : p.c0_total;
                  tma_load_5d(tm, &x_full[rx.i], dst + h * 16384, (bx & 1) * ctot + cc * 64, ox0 + (bx >> 1), by & 1,
                              oy0 + (by >> 1), n0);
                }
              }
              rx.next(p.nstage);
            }
          }
        }
      }
    }
  } else if (warp == kWgWarpY) {
    // ------------------------------------------------------------ dY producer
    if (lane == 0) {
      WRing ry;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int g, cic, cob, pt0, pt1;
        decode_work(w, g, cic, cob, pt0, pt1);
        for (int t = pt0; t < pt1; t++) {
          int tx, ty, img;
          decode_pt(t, tx, ty, img);
          const int oy0 = ty * tile_h, ox0 = tx * tile_w, n0 = img * p.nb;
          mbar_wait(&y_empty[ry.i], ry.ph ^ 1);
          mbar_expect_tx(&y_full[ry.i], p.y_bytes);
          uint8_t* dst = sY + (size_t)ry.i * p.y_bytes;
          // [co block of 64][sub-tile j][128 pixels][64 co]: the MMA's N index walks the co blocks LBO = J*16 KB apart
          for (int cb = 0; cb < WN / 64; cb++)
            for (int j = 0; j < p.J; j++)
              tma_load_4d(&tmY, &y_full[ry.i], dst + (cb * p.J + j) * 16384, cob * WN + cb * 64,
                          ox0 + 8 * j * (halo ? 1 : 0), oy0, n0);
          ry.next(p.nstage);
        }
      }
    }
  } else if (warp == kWgWarpMMA || warp == kWgWarpMMA2) {
    // ------------------------------------------------------------ MMA issuers (whole warp, elected lane issues)
    // Issuer `who` takes units u0 + who, u0 + who + 2, ...; both wait on the same full barriers and commit to the
    // same empty barriers (arrival count 2).  That includes TAPBOX X stages, which carry one unit each: the issuer
    // that does not own the unit still waits and commits, because a parity wait may lag the barrier by at most one
    // phase -- an issuer that skipped a stage's phase would mistake the phase before it for the one it wants.
    const int who = (warp == kWgWarpMMA) ? 0 : 1;
    WRing rx, ry;
    uint32_t accph = 0;
    constexpr uint32_t idesc = make_idesc_bf16(128, WN, 1, 1);  // both operands MN-major
    const uint32_t sX_u = smem_u32(sX), sY_u = smem_u32(sY);
    const uint32_t x_sbo = halo ? (uint32_t)p.halo_w * 128u : 1024u;
    const int J = p.J;
    const uint64_t b_hi = make_smem_desc_sw128(0, (uint32_t)J * 16384u, 1024u);
    const uint64_t a_hi_tap = make_smem_desc_sw128(0, 16384u, 1024u);
    const uint32_t kstep = halo ? (uint32_t)(2 * p.halo_w * 128) >> 4 : (2048u >> 4);  // 16 pixels of K per MMA
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      int g, cic, cob, pt0, pt1;
      decode_work(w, g, cic, cob, pt0, pt1);
      const int u0 = g * p.units_per_cta;
      const int u1 = min(u0 + p.units_per_cta, p.n_units);
      mbar_wait(&acc_empty, accph ^ 1);
      tc_fence_after();
      for (int t = pt0; t < pt1; t++) {
        mbar_wait(&y_full[ry.i], ry.ph);
        const uint32_t yb = sY_u + (uint32_t)ry.i * p.y_bytes;
        const uint64_t bd0 = b_hi + (uint64_t)((yb & 0x3FFFFu) >> 4);
        const uint32_t first_t = (t > pt0) ? 1u : 0u;
        if (halo) {
          mbar_wait(&x_full[rx.i], rx.ph);
          tc_fence_after();
          const uint32_t xb = sX_u + (uint32_t)rx.i * p.x_bytes;
          for (int u = u0 + who; u < u1; u += 2) {
            const WgUnit un = p.units[u];
            const uint32_t a0 = xb + (uint32_t)(un.rA * p.halo_w + un.sA) * 128u;
            const uint32_t lbo = un.sB < 0 ? 128u : (uint32_t)((un.rB - un.rA) * p.halo_w + (un.sB - un.sA)) * 128u;
            const uint64_t ad0 = make_smem_desc_sw128(a0, lbo, x_sbo);
            const uint32_t acc = tmem_base + (uint32_t)((u - u0) * WN);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 2; j++) {
                if (j < J) {
#pragma unroll
                  for (int ks = 0; ks < 8; ks++)
                    umma_bf16(acc, ad0 + (uint64_t)(ks * kstep + j * 64), bd0 + (uint64_t)(j * 1024 + ks * 128), idesc,
                              first_t | (uint32_t)(j | ks));
                }
              }
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&x_empty[rx.i]);
          __syncwarp();
          rx.next(p.nstage);
        } else {
          for (int u = u0; u < u1; u++) {
            mbar_wait(&x_full[rx.i], rx.ph);
            tc_fence_after();
            const uint32_t xb = sX_u + (uint32_t)rx.i * p.x_bytes;
            const uint64_t ad0 = a_hi_tap + (uint64_t)((xb & 0x3FFFFu) >> 4);
            const uint32_t acc = tmem_base + (uint32_t)((u - u0) * WN);
            const bool mine = ((u - u0) & 1) == who;
            if (elect_one()) {
              if (mine) {
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                  umma_bf16(acc, ad0 + (uint64_t)(ks * 128), bd0 + (uint64_t)(ks * 128), idesc, first_t | (uint32_t)ks);
              }
              umma_commit(&x_empty[rx.i]);
            }
            __syncwarp();
            rx.next(p.nstage);
          }
        }
        if (elect_one()) umma_commit(&y_empty[ry.i]);
        __syncwarp();
        ry.next(p.nstage);
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
      accph ^= 1;
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> fp32 atomics
    uint32_t accph = 0;
    const int q = warp;
    const int m = q * 32 + lane;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      int g, cic, cob, pt0, pt1;
      decode_work(w, g, cic, cob, pt0, pt1);
      const int u0 = g * p.units_per_cta;
      const int u1 = min(u0 + p.units_per_cta, p.n_units);
      mbar_wait(&acc_full, accph);
      tc_fence_after();
      if (pt1 > pt0) {
        for (int u = u0; u < u1; u++) {
          const WgUnit un = p.units[u];
          const bool second = m >= 64;
          const bool live = !second || un.sB >= 0;
          const int r = second ? un.rB : un.rA, s = second ? un.sB : un.sA;
          const int chunk = halo ? cic : (second ? un.cB : un.cA);
          const int ci = chunk * 64 + (m & 63);
          const int tap = r * p.kw + s;
          // deterministic mode: split sp owns slab sp of the gradient and stores to it (no atomics)
          float* dst = p.dw + (size_t)(w % p.splits) * p.det_slab + ((size_t)tap * p.cin_total + ci) * p.cout_pad + cob * WN;
#pragma unroll 1
          for (int cc = 0; cc < WN; cc += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((u - u0) * WN + cc), v);
            tmem_ld_wait();
            if (live) {
              if (p.det_slab) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<uint4*>(dst + cc + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + cc + i), "r"(v[i]), "r"(v[i + 1]),
                               "r"(v[i + 2]), "r"(v[i + 3])
                               : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      accph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgWarpMMA) tmem_dealloc(tmem_base, 512);
}

static int wg_ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

static int wg_act_map(CUtensorMap* tm, const gdn_act& a, int stride, const uint32_t* box4) {
  const uint64_t Hp = a.h + 2 * a.pad, Wp = a.w + 2 * a.pad, C = a.c;
  if (stride == 1) {
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)a.n};
    uint64_t str[3] = {C * 2, C * 2 * Wp, C * 2 * Wp * Hp};
    return encode_tmap_bf16(tm, a.ptr, 4, dims, str, box4);
  }
  if ((Hp & 1) || (Wp & 1)) return fail(GDN_UNSUPPORTED_SHAPE, "stride-2 source needs even padded dims");
  uint64_t dims[5] = {2 * C, Wp / 2, 2, Hp / 2, (uint64_t)a.n};
  uint64_t str[4] = {2 * C * 2, C * 2 * Wp, 2 * C * 2 * Wp, C * 2 * Wp * Hp};
  uint32_t box[5] = {box4[0], box4[1], 1, box4[2], box4[3]};
  return encode_tmap_bf16(tm, a.ptr, 5, dims, str, box);
}

}  // namespace gdn

using namespace gdn;

extern "C" __attribute__((visibility("default"))) int gdn_conv2d_wgrad(const gdn_wgrad_desc* d, gdn_stream stream) {
  if (!d || !d->x0.ptr || !d->dy.ptr || !d->dw) return fail(GDN_INVALID_DESC, "gdn_conv2d_wgrad: null pointer");
  const bool two = d->x1.ptr != nullptr;
  const int WN = (d->cout_pad % 256 == 0) ? 256 : (d->cout_pad % 128 == 0) ? 128 : 64;
  if (d->x0.c % 64 || (two && d->x1.c % 64) || d->cout_pad % 64 || d->dy.c != d->cout_pad)
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: channels must be multiples of 64 (cin %d/%d, cout %d, dy.c %d)",
                d->x0.c, two ? d->x1.c : 0, d->cout_pad, d->dy.c);
  if (d->stride != 1 && d->stride != 2) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: stride %d", d->stride);
  if (d->dy.h != d->out_h || d->dy.w != d->out_w || d->dy.pad != 0) return fail(GDN_INVALID_DESC, "gdn_conv2d_wgrad: dy extent / border");
  const int taps = d->kh * d->kw;
  const int chunks0 = d->x0.c / 64, chunks1 = two ? d->x1.c / 64 : 0, chunks = chunks0 + chunks1;

  WgK* kp = new WgK();
  WgK& k = *kp;
  struct Guard { WgK* p; ~Guard() { delete p; } } guard{kp};
  k.n_img = d->x0.n;
  k.out_h = d->out_h;
  k.out_w = d->out_w;
  k.stride = d->stride;
  k.chunks0 = chunks0;
  k.c0_total = d->x0.c;
  k.c1_total = two ? d->x1.c : 0;
  k.off_y = d->off_y + d->x0.pad;
  k.off_x = d->off_x + d->x0.pad;
  k.off_y1 = d->off_y + (two ? d->x1.pad : 0);
  k.off_x1 = d->off_x + (two ? d->x1.pad : 0);
  k.cin_total = d->x0.c + (two ? d->x1.c : 0);
  k.cout_pad = d->cout_pad;
  k.kw = d->kw;
  k.dw = d->dw;
  k.det_slab = 0;
  k.co_blocks = d->cout_pad / WN;
  const int upc_max = 512 / WN;           // units per CTA (tensor memory columns)

  const bool halo = (d->stride == 1 && taps > 1 && !two);
  k.mode = halo ? GDN_CONV_HALO : GDN_CONV_TAPBOX;
  const size_t smem_budget = 227 * 1024 - 4096 - 1024;
  CUtensorMap tmX0, tmX1, tmY;
  int rc;
  int nu = 0;
  if (halo) {
    // units: horizontal tap pairs (s, s+1) in a row; the odd last column is paired vertically
    for (int r = 0; r < d->kh; r++)
      for (int s = 0; s + 1 < d->kw; s += 2) {
        if (nu >= kMaxUnits) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: too many tap units");
        k.units[nu++] = WgUnit{(int16_t)r, (int16_t)s, 0, (int16_t)r, (int16_t)(s + 1), 0};
      }
    if (d->kw & 1) {
      const int s = d->kw - 1;
      for (int r = 0; r < d->kh; r += 2) {
        if (nu >= kMaxUnits) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: too many tap units");
        if (r + 1 < d->kh) k.units[nu++] = WgUnit{(int16_t)r, (int16_t)s, 0, (int16_t)(r + 1), (int16_t)s, 0};
        else k.units[nu++] = WgUnit{(int16_t)r, (int16_t)s, 0, 0, -1, 0};
      }
    }
    k.ci_chunks = chunks;
    int J = 2;
    {
      const int cols2 = (d->out_w + 15) / 16 * 16;
      if ((cols2 - d->out_w) * 4 >= cols2) J = 1;
    }
    int halo_w = 0;
    const int halo_h = 16 + d->kh - 1;
    for (;; J = 1) {
      halo_w = 8 * J + d->kw - 1;
      k.x_tx = (uint32_t)halo_h * halo_w * 128;
      k.x_bytes = k.x_tx + 1024;  // slack for the dummy second half of an unpaired tap
      k.x_bytes = (k.x_bytes + 1023) & ~1023u;
      k.y_bytes = (uint32_t)J * 16384 * (WN / 64);
      if (J == 1 || 2 * ((size_t)k.x_bytes + k.y_bytes) <= smem_budget) break;   // keep at least two stages
    }
    k.J = J;
    k.halo_w = halo_w;
    k.tiles_x = (d->out_w + 8 * J - 1) / (8 * J);
    k.tiles_y = (d->out_h + 15) / 16;
    k.nb = 1;
    k.groups_img = k.n_img;
    uint32_t box[4] = {64, (uint32_t)halo_w, (uint32_t)halo_h, 1};
    if ((rc = wg_act_map(&tmX0, d->x0, 1, box))) return rc;
    tmX1 = tmX0;
    uint32_t boxy[4] = {64, 8, 16, 1};
    gdn_act dy = d->dy;
    if ((rc = wg_act_map(&tmY, dy, 1, boxy))) return rc;
  } else {
    // units: pairs of (tap, chunk) slices in (tap-major, chunk-minor) order
    const int slices = taps * chunks;
    for (int i = 0; i < slices; i += 2) {
      if (nu >= kMaxUnits) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: too many (tap, chunk) units (%d)", slices);
      const int t0 = i / chunks, c0 = i % chunks;
      WgUnit u{(int16_t)(t0 / d->kw), (int16_t)(t0 % d->kw), (int16_t)c0, 0, -1, 0};
      if (i + 1 < slices) {
        const int t1 = (i + 1) / chunks, c1 = (i + 1) % chunks;
        u.rB = (int16_t)(t1 / d->kw);
        u.sB = (int16_t)(t1 % d->kw);
        u.cB = (int16_t)c1;
      }
      k.units[nu++] = u;
    }
    k.ci_chunks = 1;
    int tw = 8;
    while (tw > d->out_w && tw > 1) tw >>= 1;
    int th = 128 / tw;
    while (th / 2 >= d->out_h && th > 1) th >>= 1;
    const int nb = 128 / (tw * th);
    k.J = 1;
    k.tw_log2 = wg_ilog2(tw);
    k.th_log2 = wg_ilog2(th);
    k.nb = nb;
    k.x_bytes = 32768;
    k.x_tx = 32768;
    k.y_bytes = 16384 * (WN / 64);
    k.tiles_x = (d->out_w + tw - 1) / tw;
    k.tiles_y = (d->out_h + th - 1) / th;
    k.groups_img = (k.n_img + nb - 1) / nb;
    uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)nb};
    if ((rc = wg_act_map(&tmX0, d->x0, d->stride, box))) return rc;
    if (two) {
      if ((rc = wg_act_map(&tmX1, d->x1, d->stride, box))) return rc;
    } else {
      tmX1 = tmX0;
    }
    if ((rc = wg_act_map(&tmY, d->dy, 1, box))) return rc;
  }
  k.n_units = nu;
  k.n_groups = (nu + upc_max - 1) / upc_max;
  k.units_per_cta = (nu + k.n_groups - 1) / k.n_groups;
  k.total_pt = k.tiles_x * k.tiles_y * k.groups_img;
  const int items = k.n_groups * k.ci_chunks * k.co_blocks;
  const int sms = device_sm_count();
  int splits = sms / items;
  if (splits < 1) splits = 1;
  if (splits > k.total_pt) splits = k.total_pt;
  if (d->slabs) {
    // deterministic split-K: every split stores its partial gradient to its own slab (each split holds >= 1 pixel tile,
    // so every slab is written completely); gdn_unpack_wgrad_slabs sums them in order
    if (d->max_slabs < 1 || !d->splits_used) return fail(GDN_INVALID_DESC, "gdn_conv2d_wgrad: slabs need max_slabs >= 1 and splits_used");
    if (splits > d->max_slabs) splits = d->max_slabs;
    k.dw = d->slabs;
    k.det_slab = (size_t)taps * k.cin_total * k.cout_pad;
    *d->splits_used = splits;
  }
  k.splits = splits;
  k.total_work = items * splits;
  {
    const size_t per = (size_t)k.x_bytes + k.y_bytes;
    int st = (int)(smem_budget / per);
    if (st > kWgStagesMax) st = kWgStagesMax;
    if (st < 1) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d_wgrad: stage does not fit shared memory");
    k.nstage = st;
  }
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    GDN_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    GDN_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    GDN_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    configured[dev] = true;
  }
  const size_t smem = (size_t)k.nstage * (k.x_bytes + k.y_bytes) + 1024;
  const int grid = k.total_work < sms ? k.total_work : sms;
  const cudaStream_t cs = (cudaStream_t)stream;
  if (WN == 256) GDN_CUDA_CHECK(launch_pdl(conv_wgrad_kernel<256>, dim3(grid), dim3(kWgThreads), smem, cs, 1, tmX0, tmX1, tmY, k));
  else if (WN == 128) GDN_CUDA_CHECK(launch_pdl(conv_wgrad_kernel<128>, dim3(grid), dim3(kWgThreads), smem, cs, 1, tmX0, tmX1, tmY, k));
  else GDN_CUDA_CHECK(launch_pdl(conv_wgrad_kernel<64>, dim3(grid), dim3(kWgThreads), smem, cs, 1, tmX0, tmX1, tmY, k));
  GDN_LAUNCH_CHECK("conv_wgrad_kernel");
  return GDN_OK;
}
