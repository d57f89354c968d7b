// Implicit-GEMM convolution for sm_100a: tcgen05.mma (BF16 x BF16 -> FP32 in TMEM), TMA-staged NHWC tiles.
//
// One persistent CTA per SM, 8 warps (12 with the second epilogue group):
//   warps 0-3  epilogue  (TMEM -> registers -> bias / ReLU / residual / tanh / BN statistics -> global)
//   warp  4    A producer (activations, TMA)
//   warp  5    MMA issuer 0 (one elected thread) + TMEM owner
//   warp  6    B producer (weights, TMA)
//   warp  7    MMA issuer 1: when a tile has J >= 2 sub-tiles the two issuers split them (disjoint accumulators).
//              One issuer alone is instruction-latency bound (~70 cycles per MMA measured with ncu source
//              counters against the 48 / 64 cycles an N = 64 / 128 MMA needs), two are not.
//   warps 8-11 (template EW = 8 only) second epilogue group: the same tensor-memory lanes as warps 0-3 (a warp reaches
//              lanes 32*(warp % 4) ...), the other half of the 32-column groups.  Why: with ONE warp per scheduler the
//              epilogue runs at ~5 cycles per instruction (every instruction waits for its predecessor; ncu source
//              counters of the 1x1 convolutions, profiles/r02l_ncu_conv1x1.txt: ~530 instructions = 2 900 cycles per
//              32-column group with the BatchNorm statistics), so every launch whose reduction is shorter than that
//              (K = Cin*kh*kw < ~2 900: 1x1, stride-2, im2col'd and head layers) is epilogue-bound.  A second warp per
//              scheduler hides the dependency stalls.  Picked per layer by the caller's autotuner (bit 28 of algo).
//
// GEMM view: D[pixels(128), Cout tile(BN)] += A[pixels, 64 ch] * W[tap][Cout tile, 64 ch]^T over taps x 64-ch chunks.
//
// Two ways of staging A:
//   HALO   (stride 1, large maps): one TMA box brings the tile's whole input halo
//          [(16+kh-1) x (8J+kw-1)] pixels x 64 ch into smem ONCE per chunk; every tap then reads a shifted
//          window of it through the UMMA descriptor (start address += (r*halo_w + s)*128 B, SBO = halo_w*128 B),
//          so activations cross L2->smem once instead of kh*kw times.  J sub-tiles (16 rows x 8 cols = 128
//          pixels each) share every weight tile.
//   TAPBOX (anything: stride 2, 1x1 on a virtual concat of two sources, tiny maps): one TMA box per (tap, chunk).
// (tools/probe_umma.cu is the hardware check of the descriptor semantics this relies on.)
//
// CTA pairs (template CG = 2): two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per step on two
// adjacent pixel tiles (M = 256) against the same weight tile.  Each CTA stages its own activation halo and only
// HALF of the weight rows, so the shared-memory operand traffic per MMA drops from 4 KB + BN*32 B to
// 4 KB + BN*16 B per SM -- an M = 128, N = 64 MMA is shared-memory-bandwidth bound (6 KB per 32-cycle slot against
// 128 B/clk; measured 48 cycles), the pair needs 5 KB.  The leader CTA (cluster rank 0) issues all MMAs; both CTAs'
// TMA loads count their bytes on the leader's "full" barriers; tcgen05.commit multicasts to both CTAs' "empty" /
// "accumulator full" barriers; both CTAs' epilogue warps arrive on the leader's "accumulator empty" barrier.
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace gdn {


constexpr int kWarpA = 4, kWarpMMA = 5, kWarpB = 6, kWarpMMA2 = 7;
constexpr int kMaxA = 8, kMaxB = 8;

struct ConvK {
  int n_img, out_h, out_w;
  int tiles_x, tiles_y, cout_blocks, total_tiles;
  int mode, J;
  int tw_log2, th_log2, nb;  // TAPBOX tile decode
  int kh, kw, stride;
  int chunks0, chunks1, c0_total, c1_total;
  int off_y, off_x;          // buffer coordinates: in = out*stride + tap + off
  int off_y1, off_x1;        // same for source 1
  int halo_w;
  uint32_t a_bytes;
  int na, nbst, acc_bufs;
  int nmma;                  // MMA issuer warps in use (2 when J >= 2)
  const float* bias;
  int relu, tanh_out;
  const float* resid;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ob_pad, ob_reflect, ob_half;
  int dst_h, dst_w, dst_sy, dst_sx, dst_oy, dst_ox;
  int cout, cout_pad;
  double* stat_sum;
  double* stat_sq;
  // fused BatchNorm-backward statistics (input-gradient launches that write the LAST contribution to a tensor's gradient):
  // the unit that produced the tensor: its stored pre-BN output and per-channel (scale, shift, mean, rstd)
  const __half* bs_raw;      // fp16 [n][dst_h][dst_w][cout], NULL = off
  const float4* bs_coef;     // [cout]
  int bs_relu;
  // split-K (small maps: fewer pixel tiles than SMs): work item = (channel block, K split, pixel tile); split s reduces
  // the 64-channel chunks [s*chunks/ksplit, (s+1)*chunks/ksplit) and stores its fp32 partial tile to ws + s*ws_slab;
  // conv_splitk_combine_kernel sums the partials in a fixed order and runs the epilogue operators
  int ksplit;
  float* ws;
  size_t ws_slab;
  int det;                   // GDN_DETERMINISTIC: fixed-order accumulation of the per-CTA statistics
  // fused BatchNorm finalisation by the last CTA (see gdn_conv_desc.fin_counter)
  unsigned int* fin_counter;
  const float* fin_gamma;
  const float* fin_beta;
  float* fin_rmean;
  float* fin_rvar;
  float* fin_scale;
  float* fin_shift;
  float* fin_mean;
  float* fin_rstd;
  float4* fin_coef4;
  double fin_count;
  float fin_eps, fin_momentum;
};

struct Ring {
  int i = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void next(int n) {
    if (++i == n) {
      i = 0;
      ph ^= 1;
    }
  }
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  // saturate instead of overflowing to inf: a pre-BatchNorm value beyond +-65504 (possible with loaded checkpoints or
  // huge fan-in) must stay finite through the normalisation that follows (the reference keeps it in fp32).  One
  // F2FP.SATFINITE per pair (round to nearest, then clamp to the largest finite half) instead of four FMNMX + a convert.
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Sum v[0..NV) over the 32 lanes; afterwards lane l holds the total of element l (valid for l < NV).
template <int NV>
__device__ __forceinline__ float lane_transpose_sum(float (&v)[NV], int lane) {
  if (NV < 32) {
    // fold lanes together first so that log2(NV) scatter steps remain
#pragma unroll
    for (int off = 16; off >= NV; off >>= 1) {
#pragma unroll
      for (int i = 0; i < NV; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
  }
#pragma unroll
  for (int len = NV / 2; len >= 1; len >>= 1) {
    const bool up = (lane & len) != 0;
#pragma unroll
    for (int i = 0; i < len; i++) {
      float send = up ? v[i] : v[i + len];
      float keep = up ? v[i + len] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, len);
    }
  }
  return v[0];
}

template <int BN, int CG, int EW>
__global__ void __launch_bounds__((4 + EW) * 32, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const ConvK p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();   // the next kernel of the stream may set up while this one runs (it waits for our completion itself)
  // (the dynamic shared memory starts at the same offset in both CTAs of a pair, so the aligned layout is symmetric)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t B_BYTES = BN * 128 / CG;   // a pair splits the weight rows between its CTAs
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.na * p.a_bytes;

  __shared__ uint64_t a_full[kMaxA], a_empty[kMaxA], b_full[kMaxB], b_empty[kMaxB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_stat[2][BN >= 32 ? BN : 32];
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool halo = (p.mode == GDN_CONV_HALO);
  const int chunks = p.chunks0 + p.chunks1;
  const int taps = p.kh * p.kw;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxA; i++) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], p.nmma);
    }
    for (int i = 0; i < kMaxB; i++) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], p.nmma);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&acc_full[i], p.nmma);
      mbar_init(&acc_empty[i], EW * CG);  // pair: the epilogue warps of both CTAs release the leader's accumulators
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * (BN >= 32 ? BN : 32); i += (4 + EW) * 32) (&s_stat[0][0])[i] = 0.f;
  __syncwarp();
  if (warp == kWarpMMA) {
    if (CG == 2) {
      tmem_alloc_2sm(&tmem_base_s, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(&tmem_base_s, 512);
      tmem_relinquish();
    }
  }
  if (warp == kWarpA && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
  }
  if (warp == kWarpB && lane == 0) tma_prefetch_desc(&tmB);
  pdl_wait();      // everything above overlapped the preceding kernel; its results are visible from here on
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers are initialised too
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int acc_stride = p.J * BN;

  // work item -> (output-channel block, pixel tile); a pair works on pixel tiles 2m and 2m + 1 of the same block (the
  // odd one out lies beyond the last image: its loads are zero-filled, its pixels fail the `valid` test)
  const int cta_first = blockIdx.x / CG, cta_step = gridDim.x / CG;
  auto decode = [&](int t, int& nblk, int& tx, int& ty, int& img, int& ks) {
    nblk = t % p.cout_blocks;
    t /= p.cout_blocks;
    ks = t % p.ksplit;
    t /= p.ksplit;
    if (CG == 2) t = 2 * t + (int)cta_rank;
    tx = t % p.tiles_x;
    t /= p.tiles_x;
    ty = t % p.tiles_y;
    img = t / p.tiles_y;
  };
  const int tile_h = halo ? 16 : (1 << p.th_log2);
  const int tile_w = halo ? 8 * p.J : (1 << p.tw_log2);

  if (warp == kWarpA) {
    // ------------------------------------------------------------------ A producer
    if (lane == 0) {
      Ring ra;
      for (int t = cta_first; t < p.total_tiles; t += cta_step) {
        int nblk, tx, ty, img, ks;
        decode(t, nblk, tx, ty, img, ks);
        const int oy0 = ty * tile_h, ox0 = tx * tile_w, n0 = img * p.nb;
        const int cpk = chunks / p.ksplit;
        for (int c = ks * cpk; c < (ks + 1) * cpk; c++) {
          const bool s1 = c >= p.chunks0;
          const CUtensorMap* tm = s1 ? &tmA1 : &tmA0;
          const int cc = s1 ? c - p.chunks0 : c;
          const int offy = s1 ? p.off_y1 : p.off_y, offx = s1 ? p.off_x1 : p.off_x;
          const int ctot = s1 ? p.c1_total : p.c0_total;
          if (halo) {
            mbar_wait(&a_empty[ra.i], ra.ph ^ 1);
            if (CG == 2) {
              if (leader) mbar_expect_tx(&a_full[ra.i], 2 * p.a_bytes);
              tma_load_4d_2sm(tm, &a_full[ra.i], sA + (size_t)ra.i * p.a_bytes, cc * 64, ox0 + offx, oy0 + offy, img);
            } else {
              mbar_expect_tx(&a_full[ra.i], p.a_bytes);
              tma_load_4d(tm, &a_full[ra.i], sA + (size_t)ra.i * p.a_bytes, cc * 64, ox0 + offx, oy0 + offy, img);
            }
            ra.next(p.na);
          } else {
            for (int r = 0; r < p.kh; r++)
              for (int s = 0; s < p.kw; s++) {
                mbar_wait(&a_empty[ra.i], ra.ph ^ 1);
                if (CG == 1 || leader) mbar_expect_tx(&a_full[ra.i], CG * p.a_bytes);
                uint8_t* dst = sA + (size_t)ra.i * p.a_bytes;
                if (p.stride == 1) {
                  if (CG == 2) tma_load_4d_2sm(tm, &a_full[ra.i], dst, cc * 64, ox0 + s + offx, oy0 + r + offy, n0);
                  else tma_load_4d(tm, &a_full[ra.i], dst, cc * 64, ox0 + s + offx, oy0 + r + offy, n0);
                } else {
                  // 5D view (2C, Wp/2, 2, Hp/2, N): buffer x = 2*ox + s + off -> (x >> 1, x & 1)
                  const int bx = s + offx, by = r + offy;
                  const int fx = bx >> 1, px = bx & 1, fy = by >> 1, py = by & 1;  // arithmetic shift = floor
                  if (CG == 2) tma_load_5d_2sm(tm, &a_full[ra.i], dst, px * ctot + cc * 64, ox0 + fx, py, oy0 + fy, n0);
                  else tma_load_5d(tm, &a_full[ra.i], dst, px * ctot + cc * 64, ox0 + fx, py, oy0 + fy, n0);
                }
                ra.next(p.na);
              }
          }
        }
      }
    }
  } else if (warp == kWarpB) {
    // ------------------------------------------------------------------ B producer (weights)
    if (lane == 0) {
      Ring rb;
      for (int t = cta_first; t < p.total_tiles; t += cta_step) {
        int nblk, tx, ty, img, ks;
        decode(t, nblk, tx, ty, img, ks);
        const int cpk = chunks / p.ksplit;
        for (int c = ks * cpk; c < (ks + 1) * cpk; c++)
          for (int tap = 0; tap < taps; tap++) {
            mbar_wait(&b_empty[rb.i], rb.ph ^ 1);
            if (CG == 2) {
              if (leader) mbar_expect_tx(&b_full[rb.i], 2 * B_BYTES);
              tma_load_3d_2sm(&tmB, &b_full[rb.i], sB + (size_t)rb.i * B_BYTES, c * 64, nblk * BN + (int)cta_rank * (BN / 2), tap);
            } else {
              mbar_expect_tx(&b_full[rb.i], B_BYTES);
              tma_load_3d(&tmB, &b_full[rb.i], sB + (size_t)rb.i * B_BYTES, c * 64, nblk * BN, tap);
            }
            rb.next(p.nbst);
          }
      }
    }
  } else if (warp == kWarpMMA || warp == kWarpMMA2) {
    // ------------------------------------------------------------------ MMA issuers
    // A WHOLE warp runs this (warp-uniform) loop so that descriptors live in uniform registers; one elected
    // lane issues all MMAs of a (tap, chunk) step back to back as "64-bit descriptor base + immediate"
    // (one UIADD3.64 + one UTCHMMA per MMA -- see tools/probe_mma_rate.cu for the issue-rate measurements).
    // Issuer `who` owns sub-tiles [j0, j1) of every tile; every issuer waits on the same full barriers and
    // commits to the same empty barriers (their arrival count is the number of issuers).
    const int who = (warp == kWarpMMA) ? 0 : 1;
    if (who < p.nmma && leader) {
      Ring ra, rb, rc;
      constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN, 0, 0);
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (CG == 2) umma_bf16_2sm(d, a, b, id, acc); else umma_bf16(d, a, b, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if (CG == 2) umma_commit_2sm(bar, 3); else umma_commit(bar);
      };
      const uint32_t a_sbo = halo ? (uint32_t)p.halo_w * 128u : 1024u;
      const uint64_t a_hi = make_smem_desc_sw128(0, 0, a_sbo);
      const uint64_t b_hi = make_smem_desc_sw128(0, 0, 1024u);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      const int jn = p.J / p.nmma;                 // sub-tiles per issuer (1 or 2; J itself when alone)
      const int j0 = who * jn;
      const uint32_t a_j0 = (uint32_t)(j0 * 64);   // descriptor units (16 B): sub-tile j is 1024 B along the halo row
      const uint32_t acc_j0 = (uint32_t)(j0 * BN);
      const int kh = p.kh, kw = p.kw, na = p.na, nbst = p.nbst;
      const uint32_t a_bytes = p.a_bytes;
      const uint32_t row_skip = halo ? (uint32_t)(p.halo_w - kw) * 128u : 0u;
      for (int t = cta_first; t < p.total_tiles; t += cta_step) {
        mbar_wait(&acc_empty[rc.i], rc.ph ^ 1);
        tc_fence_after();
        const uint32_t acc_addr = tmem_base + (uint32_t)(rc.i * acc_stride) + acc_j0;
        uint32_t accum = 0u;
        for (int c = 0; c < chunks / p.ksplit; c++) {
          if (halo) mbar_wait(&a_full[ra.i], ra.ph);
          uint32_t a_off = sA_u + (uint32_t)ra.i * a_bytes;   // halo: shifted window start, advanced tap by tap
          for (int r = 0; r < kh; r++) {
            for (int s = 0; s < kw; s++) {
              if (!halo) {
                mbar_wait(&a_full[ra.i], ra.ph);
                a_off = sA_u + (uint32_t)ra.i * a_bytes;
              }
              mbar_wait(&b_full[rb.i], rb.ph);
              tc_fence_after();
              const uint64_t ad0 = a_hi + (uint64_t)(((a_off & 0x3FFFFu) >> 4) + a_j0);
              const uint64_t bd0 = b_hi + (uint64_t)(((sB_u + (uint32_t)rb.i * B_BYTES) & 0x3FFFFu) >> 4);
              if (elect_one()) {
                if (jn == 4) {
#pragma unroll
                  for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++)
                      mma(acc_addr + (uint32_t)(j * BN), ad0 + (uint64_t)(j * 64 + k4 * 2), bd0 + (uint64_t)(k4 * 2),
                          idesc, k4 ? 1u : accum);
                } else if (jn == 2) {
#pragma unroll
                  for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++)
                      mma(acc_addr + (uint32_t)(j * BN), ad0 + (uint64_t)(j * 64 + k4 * 2), bd0 + (uint64_t)(k4 * 2),
                          idesc, k4 ? 1u : accum);
                } else {
#pragma unroll
                  for (int k4 = 0; k4 < 4; k4++)
                    mma(acc_addr, ad0 + (uint64_t)(k4 * 2), bd0 + (uint64_t)(k4 * 2), idesc, k4 ? 1u : accum);
                }
                commit(&b_empty[rb.i]);
                if (!halo) commit(&a_empty[ra.i]);
              }
              __syncwarp();
              accum = 1u;
              rb.next(nbst);
              if (!halo) ra.next(na);
              a_off += 128u;
            }
            a_off += row_skip;
          }
          if (halo) {
            if (elect_one()) commit(&a_empty[ra.i]);
            __syncwarp();
            ra.next(na);
          }
        }
        if (elect_one()) commit(&acc_full[rc.i]);
        __syncwarp();
        rc.next(p.acc_bufs);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0..3 (and 8..11 when EW = 8)
    constexpr int CW = BN >= 32 ? 32 : 16;  // columns per TMEM load
    constexpr int ET = EW * 32;             // epilogue threads
    Ring rc;
    const int q = warp & 3;                 // tensor-memory lane quarter this warp can read
    const int half = warp >> 3;             // second group: the odd 32-column groups
    const int et = (q + 4 * half) * 32 + lane;
    const int m = q * 32 + lane;
    const bool do_bstats = p.bs_raw != nullptr;                  // sum g, sum g*xhat of the masked total gradient
    const bool do_stats = p.stat_sum != nullptr && !do_bstats;   // forward: sum x, sum x^2 of the raw accumulators
    const bool any_stats = p.stat_sum != nullptr && p.ksplit == 1;
    for (int t = cta_first; t < p.total_tiles; t += cta_step) {
      int nblk, tx, ty, img, ks;
      decode(t, nblk, tx, ty, img, ks);
      mbar_wait(&acc_full[rc.i], rc.ph);
      tc_fence_after();
      const uint32_t acc_addr = tmem_base + (uint32_t)(rc.i * acc_stride) + ((uint32_t)(q * 32) << 16);
      for (int j = 0; j < p.J; j++) {
        int oy, ox, n;
        if (halo) {
          oy = ty * 16 + (m >> 3);
          ox = tx * (8 * p.J) + 8 * j + (m & 7);
          n = img;
        } else {
          const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
          ox = tx * tw + (m & (tw - 1));
          oy = ty * th + ((m >> p.tw_log2) & (th - 1));
          n = img * p.nb + (m >> (p.tw_log2 + p.th_log2));
        }
        const bool valid = (oy < p.out_h) && (ox < p.out_w) && (n < p.n_img);
        const int dy = oy * p.dst_sy + p.dst_oy, dx = ox * p.dst_sx + p.dst_ox;
        const size_t pix = ((size_t)n * p.dst_h + dy) * p.dst_w + dx;
        // destinations in the padded bf16 buffer (interior + reflection images)
        int ys[3], xs[3], ny = 0, nx = 0;
        if (p.out_bf16) {
          const int P = p.ob_pad;
          ys[ny++] = dy + P;
          xs[nx++] = dx + P;
          if (p.ob_reflect) {
            if (dy >= 1 && dy <= P) ys[ny++] = P - dy;
            if (dy >= p.dst_h - 1 - P && dy <= p.dst_h - 2) ys[ny++] = P + 2 * (p.dst_h - 1) - dy;
            if (dx >= 1 && dx <= P) xs[nx++] = P - dx;
            if (dx >= p.dst_w - 1 - P && dx <= p.dst_w - 2) xs[nx++] = P + 2 * (p.dst_w - 1) - dx;
          }
        }
        const int Hp = p.dst_h + 2 * p.ob_pad, Wp = p.dst_w + 2 * p.ob_pad;
#pragma unroll 1
        for (int cc = half * CW; cc < BN; cc += (EW / 4) * CW) {
          float v[CW];
          {
            uint32_t r[CW];
            if constexpr (CW == 32) tmem_ld_32x32(acc_addr + (uint32_t)(j * BN + cc), r);
            else tmem_ld_32x16(acc_addr + (uint32_t)(j * BN + cc), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CW; i++) v[i] = __uint_as_float(r[i]);
          }
          const int cb = nblk * BN + cc;
          if (p.ksplit > 1) {
            // partial sums only: the combine kernel owns bias / ReLU / residual / statistics / conversions
            if (valid) {
              float* wp = p.ws + (size_t)ks * p.ws_slab + pix * p.cout + cb;
#pragma unroll
              for (int i = 0; i < CW; i += 4)
                *reinterpret_cast<float4*>(wp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            continue;
          }
          if (do_stats) {
            float s1[CW], s2[CW];
#pragma unroll
            for (int i = 0; i < CW; i++) {
              const float x = valid ? v[i] : 0.f;
              s1[i] = x;
              s2[i] = x * x;
            }
            const float t1 = lane_transpose_sum<CW>(s1, lane);
            const float t2 = lane_transpose_sum<CW>(s2, lane);
            if (!p.det) {
              if (lane < CW) {
                atomicAdd(&s_stat[0][cc + lane], t1);
                atomicAdd(&s_stat[1][cc + lane], t2);
              }
            } else {
              // fixed order: the epilogue warps that share this column group take turns (all of them run this loop
              // the same number of times: the conditions around it are uniform over the CTA)
              for (int w = 0; w < 4; w++) {
                if (q == w && lane < CW) {
                  s_stat[0][cc + lane] += t1;
                  s_stat[1][cc + lane] += t2;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
              }
            }
          }
          const bool full = (cb + CW <= p.cout);
          if (valid) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < CW; i++)
                if (cb + i < p.cout) v[i] += __ldg(p.bias + cb + i);
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < CW; i++) v[i] = fmaxf(v[i], 0.f);
            }
            if (p.resid) {
              const float* rp = p.resid + pix * p.cout + cb;
              if (full) {
#pragma unroll
                for (int i = 0; i < CW; i += 4) {
                  const float4 rv = *reinterpret_cast<const float4*>(rp + i);
                  v[i] += rv.x; v[i + 1] += rv.y; v[i + 2] += rv.z; v[i + 3] += rv.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < CW; i++)
                  if (cb + i < p.cout) v[i] += rp[i];
              }
            }
          }
          if constexpr (CW == 32) {
            if (do_bstats) {
              // v is now the TOTAL gradient of the destination tensor (this launch is its last contribution: host
              // contract).  Apply the ReLU mask of the unit that produced the tensor and reduce, per channel, sum g and
              // sum g*xhat -- what gdn_bn_bwd_reduce would compute in a separate pass over the fp32 gradient -- with the
              // same expressions (mask: fma(x, scale, shift) <= 0; xhat = (x - mean)*rstd).
              float s2[CW];
              if (valid) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.bs_raw + pix * p.cout + cb);
#pragma unroll
                for (int q4 = 0; q4 < CW / 8; q4++) {
                  const uint4 u = __ldg(rp + q4);
                  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int h2 = 0; h2 < 4; h2++) {
                    const float2 xy = __half22float2(*reinterpret_cast<const __half2*>(&w4[h2]));
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                      const int i = q4 * 8 + h2 * 2 + e;
                      const float x = e ? xy.y : xy.x;
                      const float4 cf = __ldg(p.bs_coef + cb + i);
                      if (p.bs_relu && fmaf(x, cf.x, cf.y) <= 0.f) v[i] = 0.f;
                      s2[i] = v[i] * ((x - cf.z) * cf.w);
                    }
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < CW; i++) s2[i] = 0.f;
              }
              float s1[CW];
#pragma unroll
              for (int i = 0; i < CW; i++) s1[i] = valid ? v[i] : 0.f;
              const float t1 = lane_transpose_sum<CW>(s1, lane);
              const float t2 = lane_transpose_sum<CW>(s2, lane);
              if (!p.det) {
                atomicAdd(&s_stat[0][cc + lane], t1);
                atomicAdd(&s_stat[1][cc + lane], t2);
              } else {
                for (int w = 0; w < 4; w++) {
                  if (q == w) {
                    s_stat[0][cc + lane] += t1;
                    s_stat[1][cc + lane] += t2;
                  }
                  asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
                }
              }
            }
          }
          if (valid) {
            if (p.tanh_out) {
#pragma unroll
              for (int i = 0; i < CW; i++) v[i] = tanhf(v[i]);
            }
            if (p.out_f32) {
              float* op = p.out_f32 + pix * p.cout + cb;
              if (full) {
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                  *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < CW; i++)
                  if (cb + i < p.cout) op[i] = v[i];
              }
            }
            if (p.out_bf16) {
              uint32_t w[CW / 2];
#pragma unroll
              for (int i = 0; i < CW / 2; i++) w[i] = p.ob_half ? pack_f16(v[2 * i], v[2 * i + 1]) : pack_bf16(v[2 * i], v[2 * i + 1]);
              for (int a = 0; a < ny; a++)
                for (int b = 0; b < nx; b++) {
                  __nv_bfloat16* op = p.out_bf16 + (((size_t)n * Hp + ys[a]) * Wp + xs[b]) * p.cout + cb;
                  if (full) {
#pragma unroll
                    for (int i = 0; i < CW / 2; i += 4)
                      *reinterpret_cast<uint4*>(op + 2 * i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
                  } else {
#pragma unroll
                    for (int i = 0; i < CW; i++)
                      if (cb + i < p.cout) {
                        if (p.ob_half) reinterpret_cast<__half*>(op)[i] = __float2half_rn(fminf(fmaxf(v[i], -65504.f), 65504.f));
                        else op[i] = __float2bfloat16(v[i]);
                      }
                  }
                }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&acc_empty[rc.i]); else mbar_arrive(&acc_empty[rc.i]);
      }
      rc.next(p.acc_bufs);
      if (any_stats && p.cout_blocks > 1) {
        // the channel block changes from tile to tile: flush the per-CTA partials now (epilogue warps only)
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
        for (int i = et; i < BN; i += ET) {
          atomicAdd(p.stat_sum + nblk * BN + i, (double)s_stat[0][i]);
          atomicAdd(p.stat_sq + nblk * BN + i, (double)s_stat[1][i]);
          s_stat[0][i] = 0.f;
          s_stat[1][i] = 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      }
    }
    if (any_stats && p.cout_blocks == 1) {
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      for (int i = et; i < BN; i += ET) {
        if (i < p.cout) {
          atomicAdd(p.stat_sum + i, (double)s_stat[0][i]);
          atomicAdd(p.stat_sq + i, (double)s_stat[1][i]);
        }
      }
    }
    if (p.fin_counter) {
      // BatchNorm finalisation by the LAST CTA whose statistics have landed (every CTA of the grid owns >= 1 tile and has
      // flushed above): the same expressions as bn_finalize_kernel (elementwise.cu), on the finished fp64 sums
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (et == 0) s_last = (atomicAdd(p.fin_counter, 1u) == gridDim.x - 1) ? 1 : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      if (s_last) {
        __threadfence();
        for (int c = et; c < p.cout; c += ET) {
          const double sum = atomicAdd(p.stat_sum + c, 0.0), sq = atomicAdd(p.stat_sq + c, 0.0);   // coherent reads
          const BnFinOut o = bn_finalize_channel(sum, sq, p.fin_count, p.fin_gamma[c], p.fin_beta[c], p.fin_eps);
          p.fin_scale[c] = o.scale;
          p.fin_shift[c] = o.shift;
          p.fin_mean[c] = o.mean;
          p.fin_rstd[c] = o.rstd;
          if (p.fin_coef4) p.fin_coef4[c] = make_float4(o.scale, o.shift, o.mean, o.rstd);
          if (p.fin_rmean) {
            p.fin_rmean[c] = bn_running_update(p.fin_rmean[c], p.fin_momentum, o.mean);
            p.fin_rvar[c] = bn_running_update(p.fin_rvar[c], p.fin_momentum, o.unbiased_var);
          }
        }
        if (et == 0) *p.fin_counter = 0u;     // ready for the next launch on this stream
      }
    }
  }
  tc_fence_before();
  if (CG == 2) {
    cluster_sync_all();   // neither CTA may leave (or free tensor memory) while its partner still works on the pair
    if (warp == kWarpMMA) tmem_dealloc_2sm(tmem_base, 512);
  } else {
    __syncthreads();
    if (warp == kWarpMMA) tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------- split-K combine
// Sums the ksplit fp32 partial tiles of a split-K convolution in a fixed order (s = 0, 1, ...: no atomics on the data
// path) and applies the operators of the convolution epilogue: forward statistics of the raw sums, bias, ReLU, fp32
// residual, fused BatchNorm-backward mask + statistics, fp32 / bf16 / fp16 stores.  One thread = 8 channels of a pixel;
// a thread keeps the same 8 channels over its grid-stride loop, so the statistics stay in registers until the end.
struct CombK {
  const float* ws;
  size_t slab;
  int S;
  long long npix;
  int C;
  const float* bias;
  int relu;
  const float* resid;
  float* out_f32;
  void* out16;
  int half;
  double* stat_sum;
  double* stat_sq;
  const __half* bs_raw;
  const float4* bs_coef;
  int bs_relu;
  int det;
};

__global__ void __launch_bounds__(256) conv_splitk_combine_kernel(const CombK k) {
  extern __shared__ float s_red[];   // [2][C]
  pdl_trigger();
  pdl_wait();
  const int cg = k.C >> 3;
  const int c8 = (threadIdx.x & (cg - 1)) * 8;
  const bool fstats = k.stat_sum && !k.bs_raw, bstats = k.bs_raw != nullptr;
  if (k.stat_sum) {
    for (int i = threadIdx.x; i < 2 * k.C; i += blockDim.x) s_red[i] = 0.f;
    __syncthreads();
  }
  float s1[8], s2[8], bi[8];
  float4 cf[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    s1[j] = s2[j] = 0.f;
    bi[j] = k.bias ? __ldg(k.bias + c8 + j) : 0.f;
    cf[j] = bstats ? __ldg(k.bs_coef + c8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long total = k.npix * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const size_t off = (size_t)i * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < k.S; s++) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(k.ws + (size_t)s * k.slab + off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(k.ws + (size_t)s * k.slab + off + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (fstats) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        s1[j] += v[j];
        s2[j] = fmaf(v[j], v[j], s2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      v[j] += bi[j];
      if (k.relu) v[j] = fmaxf(v[j], 0.f);
    }
    if (k.resid) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(k.resid + off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(k.resid + off + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (bstats) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(k.bs_raw + off));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int h2 = 0; h2 < 4; h2++) {
        const float2 xy = __half22float2(*reinterpret_cast<const __half2*>(&w4[h2]));
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int j = h2 * 2 + e;
          const float x = e ? xy.y : xy.x;
          if (k.bs_relu && fmaf(x, cf[j].x, cf[j].y) <= 0.f) v[j] = 0.f;
          s1[j] += v[j];
          s2[j] = fmaf(v[j], (x - cf[j].z) * cf[j].w, s2[j]);
        }
      }
    }
    if (k.out_f32) {
      *reinterpret_cast<float4*>(k.out_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(k.out_f32 + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (k.out16) {
      uint4 o;
      if (k.half) {
        o = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
      } else {
        o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(k.out16) + off) = o;
    }
  }
  if (k.stat_sum) {
    if (!k.det) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        atomicAdd(&s_red[c8 + j], s1[j]);
        atomicAdd(&s_red[k.C + c8 + j], s2[j]);
      }
    } else {
      // fixed order: the blockDim.x / cg threads that hold the same 8 channels take turns
      for (int r = 0; r < (int)blockDim.x / cg; r++) {
        if ((int)threadIdx.x / cg == r) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            s_red[c8 + j] += s1[j];
            s_red[k.C + c8 + j] += s2[j];
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k.C; i += blockDim.x) {
      atomicAdd(k.stat_sum + i, (double)s_red[i]);
      atomicAdd(k.stat_sq + i, (double)s_red[k.C + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------- host side
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) l++;
  return l;
}

static int make_act_map(CUtensorMap* tm, const gdn_act& a, int stride, const uint32_t* box4 /*c,w,h,n*/) {
  const uint64_t Hp = a.h + 2 * a.pad, Wp = a.w + 2 * a.pad, C = a.c;
  if (stride == 1) {
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)a.n};
    uint64_t str[3] = {C * 2, C * 2 * Wp, C * 2 * Wp * Hp};
    return encode_tmap_bf16(tm, a.ptr, 4, dims, str, box4);
  }
  if ((Hp & 1) || (Wp & 1)) return fail(GDN_UNSUPPORTED_SHAPE, "stride-2 source needs even padded dims (%d x %d)", (int)Hp, (int)Wp);
  uint64_t dims[5] = {2 * C, Wp / 2, 2, Hp / 2, (uint64_t)a.n};
  uint64_t str[4] = {2 * C * 2, C * 2 * Wp, 2 * C * 2 * Wp, C * 2 * Wp * Hp};
  uint32_t box[5] = {box4[0], box4[1], 1, box4[2], box4[3]};
  return encode_tmap_bf16(tm, a.ptr, 5, dims, str, box);
}

template <int BN, int CG, int EW>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvK& k, size_t smem,
                  cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    GDN_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<BN, CG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    configured[dev] = true;
  }
  const int slots = device_sm_count() / CG;   // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
  const int units = k.total_tiles < slots ? k.total_tiles : slots;
  GDN_CUDA_CHECK(launch_pdl(conv_igemm_kernel<BN, CG, EW>, dim3(units * CG), dim3((4 + EW) * 32), smem, st, CG, a0, a1, b, k));
  GDN_LAUNCH_CHECK("conv_igemm_kernel");
  return GDN_OK;
}

}  // namespace gdn

using namespace gdn;

extern "C" __attribute__((visibility("default"))) size_t gdn_conv2d_workspace_bytes(const gdn_conv_desc* d);

extern "C" __attribute__((visibility("default"))) int gdn_conv2d(const gdn_conv_desc* d, gdn_stream stream) {
  if (!d || !d->src0.ptr || !d->weights) return fail(GDN_INVALID_DESC, "gdn_conv2d: null descriptor / source / weights");
  const bool two = d->src1.ptr != nullptr;
  if (d->src0.c % 64 || (two && d->src1.c % 64))
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: input channels must be multiples of 64 (got %d, %d)", d->src0.c, two ? d->src1.c : 0);
  if (d->stride != 1 && d->stride != 2) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: stride %d", d->stride);
  if (d->cout_pad % 16 || d->cout > d->cout_pad) return fail(GDN_INVALID_DESC, "gdn_conv2d: cout %d / cout_pad %d", d->cout, d->cout_pad);
  if (d->cout != d->cout_pad && d->cout_pad != 16) return fail(GDN_INVALID_DESC, "gdn_conv2d: padded cout only for the 16-wide head");
  if (two && (d->src1.n != d->src0.n)) return fail(GDN_INVALID_DESC, "gdn_conv2d: source batch mismatch");
  int BN = d->cout_pad >= 256 ? 256 : d->cout_pad;
  {
    // bits 16-23 of algo: output-channel tile requested by the caller's autotuner (in units of 64; 0 = widest)
    const int bn_req = ((d->algo >> 16) & 0xff) * 64;
    if (bn_req) {
      if ((bn_req != 64 && bn_req != 128 && bn_req != 256) || bn_req > d->cout_pad || d->cout_pad % bn_req)
        return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: output-channel tile %d does not divide %d", bn_req, d->cout_pad);
      BN = bn_req;
    }
  }
  if (BN != 16 && BN != 64 && BN != 128 && BN != 256) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: cout_pad %d", d->cout_pad);
  if (d->cout_pad % BN) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: cout_pad %d not a multiple of %d", d->cout_pad, BN);
  if (d->cout % 8 && d->cout != 1) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: cout %d", d->cout);

  ConvK k{};
  k.n_img = d->src0.n;
  k.out_h = d->out_h;
  k.out_w = d->out_w;
  k.kh = d->kh;
  k.kw = d->kw;
  k.stride = d->stride;
  k.chunks0 = d->src0.c / 64;
  k.chunks1 = two ? d->src1.c / 64 : 0;
  k.c0_total = d->src0.c;
  k.c1_total = two ? d->src1.c : 0;
  k.off_y = d->off_y + d->src0.pad;
  k.off_x = d->off_x + d->src0.pad;
  k.off_y1 = d->off_y + (two ? d->src1.pad : 0);
  k.off_x1 = d->off_x + (two ? d->src1.pad : 0);
  k.cout_blocks = d->cout_pad / BN;
  k.bias = d->bias;
  k.relu = d->relu;
  k.tanh_out = d->tanh_out;
  k.resid = d->resid;
  k.out_f32 = d->out_f32;
  k.out_bf16 = (__nv_bfloat16*)d->out_bf16.ptr;
  k.ob_pad = d->out_bf16.ptr ? d->out_bf16.pad : 0;
  k.ob_reflect = d->out_reflect;
  k.ob_half = d->out16_is_half;
  k.dst_h = d->dst_h;
  k.dst_w = d->dst_w;
  k.dst_sy = d->dst_sy ? d->dst_sy : 1;
  k.dst_sx = d->dst_sx ? d->dst_sx : 1;
  k.dst_oy = d->dst_oy;
  k.dst_ox = d->dst_ox;
  k.cout = d->cout;
  k.cout_pad = d->cout_pad;
  k.stat_sum = d->stat_sum;
  k.stat_sq = d->stat_sqsum;
  k.bs_raw = reinterpret_cast<const __half*>(d->bwd_raw);
  k.bs_coef = reinterpret_cast<const float4*>(d->bwd_coef);
  k.bs_relu = d->bwd_relu;
  if (d->bwd_raw) {
    if (!d->bwd_coef || !d->stat_sum || !d->stat_sqsum)
      return fail(GDN_INVALID_DESC, "gdn_conv2d: bwd_raw needs bwd_coef and the stat_sum / stat_sqsum accumulators");
    if (d->cout != d->cout_pad || d->cout % 32 || d->tanh_out || d->out_reflect || d->out16_is_half ||
        (d->out_bf16.ptr && d->out_bf16.pad))
      return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: fused BatchNorm-backward statistics need cout %% 32 == 0 and plain outputs");
  }
  if (d->out_bf16.ptr && (d->out_bf16.h != d->dst_h || d->out_bf16.w != d->dst_w || d->out_bf16.c != d->cout))
    return fail(GDN_INVALID_DESC, "gdn_conv2d: out_bf16 extent %dx%dx%d != dst %dx%dx%d", d->out_bf16.h, d->out_bf16.w,
                d->out_bf16.c, d->dst_h, d->dst_w, d->cout);
  if ((k.out_h - 1) * k.dst_sy + k.dst_oy >= k.dst_h || (k.out_w - 1) * k.dst_sx + k.dst_ox >= k.dst_w)
    return fail(GDN_INVALID_DESC, "gdn_conv2d: outputs fall outside the destination");
  if (d->out_reflect && d->out_bf16.ptr && (d->out_bf16.pad >= d->dst_h || d->out_bf16.pad >= d->dst_w))
    return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: reflection border %d >= extent", d->out_bf16.pad);

  int mode = d->algo & 0xff;
  const int cg = ((d->algo >> 24) & 1) ? 2 : 1;   // bit 24 of algo: CTA pairs (tcgen05 cta_group::2), HALO mode only
  // bits 25-27 of algo: split-K factor (2 or 4) for maps with fewer pixel tiles than SMs; needs the caller's workspace
  int ksplit = (d->algo >> 25) & 7;
  if (ksplit < 2) ksplit = 1;
  if (ksplit > 1) {
    const int chunks_all = d->src0.c / 64 + (two ? d->src1.c / 64 : 0);
    const size_t need = gdn_conv2d_workspace_bytes(d);
    if ((ksplit != 2 && ksplit != 4) || chunks_all % ksplit || k.dst_sy != 1 || k.dst_sx != 1 || k.dst_oy || k.dst_ox ||
        d->dst_h != d->out_h || d->dst_w != d->out_w || d->tanh_out || d->out_reflect || (d->out_bf16.ptr && d->out_bf16.pad) ||
        d->cout != d->cout_pad || d->cout % 8 || (d->cout & (d->cout - 1)) || d->cout > 2048)
      return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: split-K %d not applicable to this launch", ksplit);
    if (!d->workspace || d->workspace_bytes < need)
      return fail(GDN_WORKSPACE_TOO_SMALL, "gdn_conv2d: split-K needs %zu bytes of workspace (got %zu)", need,
                  (size_t)d->workspace_bytes);
  }
  k.ksplit = ksplit;
  k.det = det_enabled() ? 1 : 0;
  if (d->fin_counter) {
    if (ksplit > 1 || d->bwd_raw || !d->stat_sum || !d->stat_sqsum || !d->fin_gamma || !d->fin_beta || !d->fin_scale ||
        !d->fin_shift || !d->fin_mean || !d->fin_rstd || !(d->fin_count > 0) || (d->fin_running_mean && !d->fin_running_var))
      return fail(GDN_INVALID_DESC, "gdn_conv2d: fused BatchNorm finalisation needs the forward statistics, all its outputs, no split-K");
    k.fin_counter = d->fin_counter;
    k.fin_gamma = d->fin_gamma; k.fin_beta = d->fin_beta;
    k.fin_rmean = d->fin_running_mean; k.fin_rvar = d->fin_running_var;
    k.fin_scale = d->fin_scale; k.fin_shift = d->fin_shift; k.fin_mean = d->fin_mean; k.fin_rstd = d->fin_rstd;
    k.fin_coef4 = reinterpret_cast<float4*>(d->fin_coef4);
    k.fin_count = d->fin_count; k.fin_eps = d->fin_eps; k.fin_momentum = d->fin_momentum;
  }
  k.ws = reinterpret_cast<float*>(d->workspace);
  k.ws_slab = (size_t)d->src0.n * d->out_h * d->out_w * d->cout;
  const int j_req = (d->algo >> 8) & 0xff;  // HALO: sub-tiles per tile requested by the caller's autotuner (0 = heuristic)
  const int taps = d->kh * d->kw;
  if (mode == GDN_CONV_AUTO)
    mode = (d->stride == 1 && !two && taps > 1 && d->out_h >= 16 && d->out_w >= 16) ? GDN_CONV_HALO : GDN_CONV_TAPBOX;
  if (mode == GDN_CONV_HALO && (d->stride != 1 || two)) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: HALO needs stride 1, one source");
  k.mode = mode;
  if (cg == 2 && BN < 64) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: CTA pairs need >= 64 output channels per tile");

  const size_t smem_budget = 227 * 1024 - 4096 - 1024;  // dynamic smem minus alignment slack
  const uint32_t b_bytes = BN * 128 / cg;
  CUtensorMap tmA0, tmA1, tmB;
  int rc;
  if (mode == GDN_CONV_HALO) {
    int J = BN <= 64 ? 4 : 2;
    if (j_req) {
      if ((j_req != 1 && j_req != 2 && j_req != 4) || j_req * BN > 512)
        return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: %d sub-tiles of %d channels do not fit tensor memory", j_req, BN);
      J = j_req;
    } else {
      // halve J while a quarter or more of the computed columns would fall outside the image
      while (J > 1) {
        const int cols = (d->out_w + 8 * J - 1) / (8 * J) * (8 * J);
        if ((cols - d->out_w) * 4 >= cols) J /= 2; else break;
      }
    }
    const int halo_h = 16 + d->kh - 1;
    int halo_w = 8 * J + d->kw - 1;
    if (halo_h > 256 || halo_w > 256) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: halo too large");
    k.J = J;
    k.halo_w = halo_w;
    k.a_bytes = (uint32_t)halo_h * halo_w * 128;
    // two halo buffers whenever they fit: with one 64-channel chunk the second buffer prefetches the NEXT tile's halo
    k.na = (2 * (size_t)k.a_bytes + 3 * b_bytes <= smem_budget) ? 2 : 1;
    size_t rem = smem_budget - (size_t)k.na * k.a_bytes;
    k.nbst = (int)(rem / b_bytes);
    if (k.nbst > kMaxB) k.nbst = kMaxB;
    if (k.nbst < 2) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: halo tile does not fit shared memory");
    k.acc_bufs = (2 * J * BN <= 512) ? 2 : 1;
    k.nmma = J >= 2 ? 2 : 1;
    k.tiles_x = (d->out_w + 8 * J - 1) / (8 * J);
    k.tiles_y = (d->out_h + 15) / 16;
    k.nb = 1;
    k.total_tiles = (k.tiles_x * k.tiles_y * k.n_img + cg - 1) / cg * k.cout_blocks;   // work items (pairs when cg = 2)
    uint32_t box[4] = {64, (uint32_t)halo_w, (uint32_t)halo_h, 1};
    if ((rc = make_act_map(&tmA0, d->src0, 1, box))) return rc;
    tmA1 = tmA0;
  } else {
    int tw = 8;
    while (tw > d->out_w && tw > 1) tw >>= 1;
    if (taps == 1 && d->stride == 1 && d->out_w >= 16) {
      // 1x1 convolutions need no 2-D locality: make the 128-pixel tile as WIDE as the row allows without padding it
      // (416 columns: 4 rows x 32 pixels), so that every TMA row and every warp-wide store is one contiguous run of
      // 32 pixels instead of 8 -- these launches are bound by HBM / the epilogue, not by the MMAs
      int best = tw, best_cols = (d->out_w + tw - 1) / tw * tw;
      for (int c = 16; c <= 128 && c <= d->out_w; c <<= 1) {
        const int cols = (d->out_w + c - 1) / c * c;
        if (cols <= best_cols) { best = c; best_cols = cols; }
      }
      tw = best;
    }
    int th = 128 / tw;
    while (th / 2 >= d->out_h && th > 1) th >>= 1;
    int nb = 128 / (tw * th);
    k.J = 1;
    k.tw_log2 = ilog2(tw);
    k.th_log2 = ilog2(th);
    k.nb = nb;
    k.a_bytes = 128 * 128;
    k.nmma = 1;
    k.acc_bufs = (2 * BN <= 512) ? 2 : 1;
    size_t per = k.a_bytes + b_bytes;
    int st = (int)(smem_budget / per);
    if (st > kMaxB) st = kMaxB;
    if (st < 2) return fail(GDN_UNSUPPORTED_SHAPE, "gdn_conv2d: stage does not fit");
    k.nbst = st;
    k.na = st;
    k.tiles_x = (d->out_w + tw - 1) / tw;
    k.tiles_y = (d->out_h + th - 1) / th;
    const int groups = (k.n_img + nb - 1) / nb;
    k.total_tiles = (k.tiles_x * k.tiles_y * groups + cg - 1) / cg * k.cout_blocks;   // work items (pairs when cg = 2)
    uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)nb};
    if ((rc = make_act_map(&tmA0, d->src0, d->stride, box))) return rc;
    if (two) {
      if ((rc = make_act_map(&tmA1, d->src1, d->stride, box))) return rc;
    } else {
      tmA1 = tmA0;
    }
  }
  {
    const uint64_t cin_total = (uint64_t)d->src0.c + (two ? d->src1.c : 0);
    uint64_t dims[3] = {cin_total, (uint64_t)d->cout_pad, (uint64_t)taps};
    uint64_t str[2] = {cin_total * 2, cin_total * 2 * d->cout_pad};
    uint32_t box[3] = {64, (uint32_t)(BN / cg), 1};
    if ((rc = encode_tmap_bf16(&tmB, const_cast<void*>(d->weights), 3, dims, str, box))) return rc;
  }
  const size_t smem = (size_t)k.na * k.a_bytes + (size_t)k.nbst * b_bytes + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  k.total_tiles *= k.ksplit;
  // bit 28 of algo: second epilogue warp group (8 epilogue warps) -- for launches with a short reduction
  const bool ew8 = ((d->algo >> 28) & 1) && BN >= 64;
  if (cg == 2) {
    switch (BN) {
      case 64: rc = ew8 ? launch<64, 2, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<64, 2, 4>(tmA0, tmA1, tmB, k, smem, st); break;
      case 128: rc = ew8 ? launch<128, 2, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<128, 2, 4>(tmA0, tmA1, tmB, k, smem, st); break;
      default: rc = ew8 ? launch<256, 2, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<256, 2, 4>(tmA0, tmA1, tmB, k, smem, st); break;
    }
  } else {
    switch (BN) {
      case 16: rc = launch<16, 1, 4>(tmA0, tmA1, tmB, k, smem, st); break;
      case 64: rc = ew8 ? launch<64, 1, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<64, 1, 4>(tmA0, tmA1, tmB, k, smem, st); break;
      case 128: rc = ew8 ? launch<128, 1, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<128, 1, 4>(tmA0, tmA1, tmB, k, smem, st); break;
      default: rc = ew8 ? launch<256, 1, 8>(tmA0, tmA1, tmB, k, smem, st) : launch<256, 1, 4>(tmA0, tmA1, tmB, k, smem, st); break;
    }
  }
  if (rc || k.ksplit == 1) return rc;
  CombK c{};
  c.ws = k.ws; c.slab = k.ws_slab; c.S = k.ksplit;
  c.npix = (long long)d->src0.n * d->out_h * d->out_w;
  c.C = d->cout;
  c.bias = d->bias; c.relu = d->relu; c.resid = d->resid; c.out_f32 = d->out_f32;
  c.out16 = d->out_bf16.ptr; c.half = d->out16_is_half;
  c.stat_sum = d->stat_sum; c.stat_sq = d->stat_sqsum;
  c.bs_raw = k.bs_raw; c.bs_coef = k.bs_coef; c.bs_relu = k.bs_relu;
  c.det = k.det;
  const long long items = c.npix * (c.C / 8);
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)device_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  GDN_CUDA_CHECK(launch_pdl(conv_splitk_combine_kernel, dim3((int)blocks), dim3(256), 2 * c.C * sizeof(float), st, 1, c));
  GDN_LAUNCH_CHECK("conv_splitk_combine_kernel");
  return GDN_OK;
}

extern "C" __attribute__((visibility("default"))) size_t gdn_conv2d_workspace_bytes(const gdn_conv_desc* d) {
  if (!d) return 0;
  int ksplit = (d->algo >> 25) & 7;
  if (ksplit < 2) return 0;
  return (size_t)ksplit * d->src0.n * d->out_h * d->out_w * d->cout * sizeof(float);
}
