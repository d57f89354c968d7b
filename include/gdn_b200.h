/* gdn_b200 -- C ABI of the B200-native GDN hot path (libgdn_b200.so).
 *
 * Plain C, no torch / C++ types.  Every pointer is a raw device address owned by the caller (the PyTorch
 * caching allocator in practice: tensor.data_ptr()); the library never allocates or frees on the hot path,
 * never synchronises, and launches everything on the caller's stream.  Return value: 0 on success, a negative
 * gdn_status otherwise; gdn_last_error() gives a thread-local message.  There is no CPU fallback.
 *
 * The reference has no FFI of its own (pure PyTorch); each entry point replaces the library-op call sites the
 * reference makes on this path (file:line into the reference tree):
 *   gdn_conv2d            nn.Conv2d / nn.ConvTranspose2d forward and input-gradient:
 *                         src/AE_model_unet.py:50,53,67,85,101-104,127-134,292,521  (cuDNN via ATen)
 *   gdn_conv2d_wgrad      weight gradient of the same convolutions (autograd of the above)
 *   gdn_im2col            thin-channel layers (Cin = 1 / 3, Cout = 1) reshaped for the tensor-core path:
 *                         src/AE_model_unet.py:272 (3->64), :494 (1->64), :292/:521 (64->1 heads)
 *   gdn_head_gather       second half of the 64 -> 1 heads (src/AE_model_unet.py:300,362 / :521,570): the 81 taps run as the
 *                         N dimension of ONE 1x1 gdn_conv2d (Z[p][tap]), this sums Z over the 9x9 neighbourhood + tanh
 *   gdn_act_forward       nn.BatchNorm2d apply + nn.ReLU + residual add + nn.ReflectionPad2d + F.interpolate(x2):
 *                         src/AE_model_unet.py:51-57,66-69,86-87,135,336-355
 *   gdn_bn_finalize       batch statistics -> scale/shift, running-stat update (nn.BatchNorm2d training mode)
 *   gdn_bn_bwd_reduce / gdn_act_backward   autograd of the above
 *   gdn_fold_grad         adjoint of reflection padding / bilinear x2 upsampling / zero-dilation
 *   gdn_loss (mode 0 = RtoD, 1 = DtoD) / gdn_sqdiff_sum (feature MSE)   src/trainer.py:433-456, 705-757; src/utils.py:105-178
 *   gdn_eigen_metrics     src/calculate_error.py:10-103
 *   gdn_adam_step         optim.Adam(lr, [0.9,0.999], eps=1e-8, weight_decay=5e-4): src/GDN_main.py:157,173
 *   gdn_pack_weights / gdn_unpack_wgrad   fp32 OIHW <-> bf16 [tap][Cout][Cin] operand layout (BN folding at eval)
 *   gdn_act_backward_frozen / gdn_sqdiff_grad / gdn_tanh_chain_add   opt-in guidance gradient: autograd of the frozen
 *                         DtoD encoder passes of src/trainer.py:699-703,726-733 with the no_grad removed
 *   gdn_preprocess_u8     src/transform_list.py:84-113,161-203 (data-loader transforms)
 *   gdn_bytescale / gdn_resize_u8   scipy.misc.imresize in src/depth_extract.py:23-58,86,138 (demo path)
 */
#ifndef GDN_B200_H
#define GDN_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GDN_OK = 0,
  GDN_INVALID_DESC = -1,
  GDN_UNSUPPORTED_SHAPE = -2,
  GDN_WORKSPACE_TOO_SMALL = -3,
  GDN_CUDA_ERROR = -4
} gdn_status;

typedef void* gdn_stream; /* cudaStream_t */

/* NHWC bf16 activation with a physical border of `pad` pixels on every side:
 * buffer is [n][h + 2*pad][w + 2*pad][c], interior pixel (y,x) lives at (y+pad, x+pad). */
typedef struct {
  void* ptr;
  int32_t n, h, w, c, pad;
} gdn_act;

enum { GDN_CONV_AUTO = 0, GDN_CONV_TAPBOX = 1, GDN_CONV_HALO = 2 };

/* Implicit-GEMM convolution  out[n, oy, ox, co] = sum_{r,s,ci} in[n, oy*stride + r + off_y, ox*stride + s + off_x, ci] * w[r*kw+s][co][ci]
 * Coordinates are INTERIOR coordinates of the source activation(s); reads that fall outside the physical
 * buffer return 0 (TMA out-of-bounds fill), reads inside the border return whatever the producer wrote there
 * (reflection halo).  src1 (optional, c1 > 0) is a second source whose channels follow src0's (virtual concat).
 * Epilogue, in this order:  v = acc (+ bias[co]) ; ReLU ; (+ resid) ; tanh ; stores.  */
typedef struct {
  gdn_act src0, src1;        /* src1.ptr == NULL when unused */
  const void* weights;       /* bf16 [kh*kw][cout_pad][cin_total] */
  int32_t kh, kw, stride;    /* stride 1 or 2 */
  int32_t off_y, off_x;
  int32_t out_h, out_w;      /* outputs computed per image */
  int32_t cout;              /* real output channels */
  int32_t cout_pad;          /* channels in the weight tensor (multiple of 16; == cout unless cout < 16) */
  int32_t algo;              /* GDN_CONV_* in the low byte; bits 8-15: HALO sub-tiles per tile (1, 2, 4), bits 16-23: output-channel
                                tile / 64 (1, 2, 4); bit 24: CTA pairs; bits 25-27: split-K (see workspace); bit 28: eight
                                epilogue warps instead of four (launches with a short reduction are epilogue-bound);
                                0 = library heuristic.  Every choice WITHOUT split-K produces bit-identical
                                outputs (BN statistics aside: atomics); callers may time them. */
  /* epilogue */
  const float* bias;         /* [cout] or NULL */
  int32_t relu, tanh_out;
  const float* resid;        /* fp32 [n][dst_h][dst_w][cout] or NULL; added after ReLU */
  float* out_f32;            /* fp32 [n][dst_h][dst_w][cout] or NULL */
  gdn_act out_bf16;          /* ptr NULL when unused; interior dims = dst_h x dst_w */
  int32_t out_reflect;       /* also write the reflection halo of out_bf16 (border = out_bf16.pad) */
  int32_t dst_h, dst_w;      /* destination extent; output (oy,ox) goes to (oy*dst_sy + dst_oy, ox*dst_sx + dst_ox) */
  int32_t dst_sy, dst_sx, dst_oy, dst_ox;
  double* stat_sum;          /* [cout] += sum over pixels of acc, or NULL */
  double* stat_sqsum;        /* [cout] += sum of acc^2 */
  int32_t out16_is_half;     /* store out_bf16 as IEEE fp16 instead (pre-BatchNorm conv outputs: they only feed
                                elementwise kernels, and fp16's 10-bit mantissa keeps BN's mean subtraction accurate) */
  /* Fused BatchNorm(+ReLU)-backward statistics, for an input-gradient launch that writes the LAST contribution to the
   * gradient of a tensor t (the caller knows the order of its backward plan).  With v = acc (+ resid) the total gradient:
   *   g = v * [fma(raw, scale, shift) > 0]   (only when bwd_relu);   stat_sum[co] += sum g;   stat_sqsum[co] += sum g*xhat,
   *   xhat = (raw - mean) * rstd;   the stores then write g (fp32 and / or bf16)
   * -- exactly what gdn_bn_bwd_reduce computes from the fp32 gradient, without the extra pass over it.
   * bwd_raw: fp16 pre-BN output of the unit that PRODUCED t, [n][dst_h][dst_w][cout]; bwd_coef: that unit's
   * (scale, shift, mean, rstd) per channel as written by gdn_bn_finalize(coef4).  Needs cout % 32 == 0. */
  const void* bwd_raw;       /* NULL = off (then stat_sum / stat_sqsum are the forward statistics above) */
  const float* bwd_coef;
  int32_t bwd_relu;
  /* Split-K (bits 25-27 of algo = 2 or 4): for maps with fewer 128-pixel tiles than SMs the reduction over the input
   * channels is split across CTAs; the fp32 partial tiles go to `workspace` (caller-owned, gdn_conv2d_workspace_bytes(d)
   * bytes, private to the stream) and a second kernel on the same stream sums them in a fixed order and applies the
   * epilogue above.  Plain destinations only (no strides, border, reflection, tanh; cout a power of two). */
  void* workspace;
  size_t workspace_bytes;
  /* Fused BatchNorm finalisation (training forward, optional): with `fin_counter` set (a device uint32 that is 0 before
   * the launch; the launch leaves it 0), the LAST CTA to flush its statistics turns stat_sum / stat_sqsum into what
   * gdn_bn_finalize would compute from them -- same expressions, same results -- so the separate finalise launch between
   * the convolution and the BatchNorm-apply pass disappears: scale = gamma*rstd, shift = beta - mean*scale, mean, rstd,
   * coef4 (optional), running statistics (momentum, unbiased variance; optional).  Not with split-K or bwd_raw. */
  uint32_t* fin_counter;
  const float* fin_gamma;
  const float* fin_beta;
  float* fin_running_mean;   /* may be NULL */
  float* fin_running_var;
  float* fin_scale;
  float* fin_shift;
  float* fin_mean;
  float* fin_rstd;
  float* fin_coef4;          /* may be NULL */
  double fin_count;
  float fin_eps, fin_momentum;
} gdn_conv_desc;

int gdn_conv2d(const gdn_conv_desc* d, gdn_stream stream);
size_t gdn_conv2d_workspace_bytes(const gdn_conv_desc* d);   /* 0 unless the algo word asks for split-K */

/* Weight gradient: dw[tap][ci][co] (fp32, += ) = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*stride + r + off_y, ox*stride + s + off_x, ci] */
typedef struct {
  gdn_act x0, x1;            /* forward input(s) of the convolution (virtual concat like gdn_conv_desc) */
  gdn_act dy;                /* gradient of the conv output, interior = out_h x out_w, c = cout_pad */
  float* dw;                 /* fp32 [kh*kw][cin_total][cout_pad] (co fastest), accumulated into (caller zeroes) */
  int32_t kh, kw, stride, off_y, off_x, out_h, out_w, cout_pad;
  /* Deterministic split-K (optional; the engine sets it under GDN_DETERMINISTIC=1).  The reduction over the pixels is
   * split over the SMs; by default every split ADDS its partial gradient to dw with fp32 atomics (order varies from run
   * to run).  With `slabs` set, split s STORES its partial to slabs + s * kh*kw*cin_total*cout_pad floats instead (dw is
   * not touched, nothing needs zeroing), at most max_slabs splits are used, and their number is written to *splits_used
   * (host memory, at call time).  gdn_unpack_wgrad_slabs sums the slabs in index order. */
  float* slabs;
  int32_t max_slabs;
  int32_t* splits_used;
} gdn_wgrad_desc;

int gdn_conv2d_wgrad(const gdn_wgrad_desc* d, gdn_stream stream);

/* ---- HBM-bound helpers (NHWC; bf16 activations, fp32 residual / gradient streams) ---------------------- */

/* thin-channel lowering: src fp32 planar [n][c][h][w] (c <= 4) -> dst bf16 [n*h*w][kpad],
 * column k = (r*kw + s)*c + ch holds src[n][ch][y + r - pad][x + s - pad] (reflection or zero padding). */
int gdn_im2col(const float* src, void* dst, int n, int c, int h, int w, int kh, int kw, int pad, int reflect,
               int kpad, gdn_stream stream);

/* Single-output-channel k x k convolution with zero padding, given Z[p][t] = <x[p][:], w[t][:]> for every tap t = r*k + s
 * (a 1x1 gdn_conv2d with the taps as output channels; z: [n][h][w][zc] fp16 (z_is_half) or bf16, zc % 8 == 0, zc >= k*k
 * rounded up to 8):  out[n][y][x] = act( sum_{r,s} z[n][y + r - pad][x + s - pad][r*k + s] ), taps outside the image
 * skipped (= zero padding of x), act = tanh when tanh_out.  Fixed summation order (r-major).  k = 9. */
int gdn_head_gather(const void* z, int z_is_half, int zc, int n, int h, int w, int k, int pad, int tanh_out, float* out,
                    gdn_stream stream);

/* batch statistics -> per-channel scale = gamma*rstd, shift = beta - mean*scale; saves mean / rstd for backward;
 * updates running stats (momentum, unbiased variance) when running_mean != NULL.  coef4 (optional, 16-byte aligned
 * [c][4] floats) receives (scale, shift, mean, rstd) interleaved: the operand of gdn_conv_desc.bwd_coef. */
int gdn_bn_finalize(const double* sum, const double* sqsum, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                    float* mean, float* rstd, float* coef4, int c, gdn_stream stream);
/* eval mode: scale = gamma / sqrt(running_var + eps), bias = beta - running_mean*scale */
int gdn_bn_fold(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps, float* scale,
                float* bias, int c, gdn_stream stream);

/* y = [ReLU](src*scale + shift) [+ resid];  out_f32 <- y ;  out_bf16 <- pad/upsample/dilate(y)
 * out_bf16 is [n][OH + 2*pad][OW + 2*pad][c] with OH = 2h when up or dilate; reflect fills the border by
 * reflection (otherwise the border is left untouched); up: 1 = bilinear x2 align_corners=False, 2 = True;
 * dilate: y at even positions, zeros elsewhere. */
typedef struct {
  const void* src_bf16;
  const float* src_f32;
  const float* scale;
  const float* shift;
  const float* resid;
  int32_t relu;
  int32_t n, h, w, c;
  float* out_f32;
  void* out_bf16;
  int32_t pad, reflect, up, dilate;
  int32_t src16_is_half;    /* src_bf16 holds fp16 (raw conv output written with out16_is_half) */
} gdn_act_fwd_desc;
int gdn_act_forward(const gdn_act_fwd_desc* d, gdn_stream stream);

/* BatchNorm(+ReLU) backward.  reduce: sum_g[c] += sum g, sum_gx[c] += sum g*xhat with g = dact*[bn(raw) > 0];
 * apply (gdn_act_backward): dy = scale*(g - sum_g/n - xhat*sum_gx/n) as bf16 (optionally zero-dilated x2),
 * dgamma += sum_gx, dbeta += sum_g. */
typedef struct {
  const void* dact;        /* fp32 [n][h][w][c], or bf16 when dact_is_bf16 */
  const void* raw;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  int32_t relu;
  int32_t n, h, w, c;
  double* sum_g;
  double* sum_gx;
  void* dy;
  int32_t dilate;
  float* dgamma;
  float* dbeta;
  int32_t raw_is_half;
  int32_t dact_is_bf16;    /* dact is the bf16 buffer a gdn_conv2d epilogue wrote in bwd_stats mode (ReLU mask applied) */
} gdn_bn_bwd_desc;
int gdn_bn_bwd_reduce(const gdn_bn_bwd_desc* d, gdn_stream stream);
int gdn_act_backward(const gdn_bn_bwd_desc* d, gdn_stream stream);

/* Backward through a FROZEN unit (eval mode, BatchNorm folded into the weights): dy = dact * [y > 0] * scale[c] as
 * bf16, y = the unit's own post-ReLU output (fp32 NHWC or plain bf16 NHWC; exactly one when relu).  Opt-in
 * guidance gradient through the frozen DtoD encoder (the published trainer.py:699-703 blocks it
 * with no_grad).  No statistics, no parameter gradients. */
typedef struct {
  const float* dact;
  const float* y_f32;
  const void* y_bf16;
  const float* scale;     /* folded BatchNorm scale per channel, or NULL */
  int32_t relu;
  int32_t n, h, w, c;
  void* dy;
} gdn_frozen_bwd_desc;
int gdn_act_backward_frozen(const gdn_frozen_bwd_desc* d, gdn_stream stream);

/* adjoint of the input transform of a conv: dpad is the fp32 gradient w.r.t. the conv's (padded / upsampled /
 * dilated) input buffer [n][OH + 2*pad][OW + 2*pad][ctot]; channels [c_off, c_off + c) are folded back onto the
 * source activation gradient dact [n][h][w][c] (+= when accumulate).  Channel counts that are not multiples of 4
 * (the 1- / 3-channel network input) are supported for reflection / zero padding only. */
typedef struct {
  const void* dpad;        /* fp32, or bf16 when dpad_is_bf16 (written by gdn_conv2d through out_bf16) */
  int32_t ctot, c_off;
  int32_t n, h, w, c;
  int32_t pad, reflect, up, dilate;
  float* dact;
  int32_t accumulate;
  int32_t dpad_is_bf16;
} gdn_fold_desc;
int gdn_fold_grad(const gdn_fold_desc* d, gdn_stream stream);

/* fp32 parameter tensor <-> packed operand layout.  packed[t][ai][bi] = w[ai*stride_a + bi*stride_b + r*stride_r +
 * s*stride_s] (t = r*kw + s, taps flipped when flip), zero padded to a_pad x b_pad, optionally scaled per a.
 * col_c > 0: im2col'd layer, packed[0][ai][(r*kw + s)*col_c + ch] with stride_b the channel stride.
 * gdn_unpack_wgrad scatters the wgrad result dw[t][bi][ai] back to the parameter-gradient layout. */
typedef struct {
  int32_t kh, kw, a, b, a_pad, b_pad;
  int64_t stride_a, stride_b, stride_r, stride_s;
  int32_t flip, col_c;
} gdn_pack_desc;
int gdn_pack_weights(const gdn_pack_desc* d, const float* w, const float* scale_a, void* out, gdn_stream stream);
int gdn_unpack_wgrad(const gdn_pack_desc* d, const float* dw, float* grad, int accumulate, gdn_stream stream);
/* same, from `slabs` partial gradients lying slab_elems floats apart (deterministic split-K of gdn_conv2d_wgrad), summed in
 * index order */
int gdn_unpack_wgrad_slabs(const gdn_pack_desc* d, const float* dw, int slabs, int64_t slab_elems, float* grad, int accumulate,
                           gdn_stream stream);
/* Batched re-pack: every packed tensor of a network in ONE launch.  The caller builds a job table on the host with
 * gdn_pack_job_fill (entries of gdn_pack_job_size() bytes, cta0 = running sum of the n_ctas returned so far), copies
 * it to device memory once, and calls gdn_pack_weights_table after every optimizer step.  Non-tileable tensors
 * (the im2col'd thin first layers) return GDN_UNSUPPORTED_SHAPE from gdn_pack_job_fill and keep using
 * gdn_pack_weights. */
int gdn_pack_job_size(void);
int gdn_pack_job_fill(const gdn_pack_desc* d, const float* w, const float* scale_a, void* out, int cta0, void* job_out,
                      int* n_ctas);
int gdn_pack_weights_table(const void* jobs_dev, int njobs, int total_ctas, int max_taps, gdn_stream stream);

/* ---- device-side input pipeline (the reference's per-sample CPU transforms) ----------------
 * src: uint8 [n][h][w][c] (HWC, as decoded); dst: fp32 [n][c][h][w] in [-1, 1] = Normalize(0.5, 0.5)(ArrayToTensor(.)),
 * transform_list.py:84-113.  flip[n] != 0 mirrors sample n horizontally (RandomHorizontalFlip, :161-169); crop[n] =
 * (scaled_h, scaled_w, off_y, off_x) zooms sample n to scaled_h x scaled_w (scaled >= h, w) and keeps the h x w window
 * at the offset (RandomScaleCrop, :189-203).  The zoom is scipy.misc.imresize on the loader's float32 image: per-image
 * min-max byte scaling of the whole array + PIL's two-pass 8-bit BILINEAR resize, bit-exact (the arithmetic of
 * gdn_bytescale / gdn_resize_u8 below).  Both may be NULL.  scratch: 8*n bytes of device memory (per-image min / max),
 * required when crop != NULL.  The random draws stay on the host (gdn_pytorch_b200/data.py replays the reference's RNG
 * calls). */
int gdn_preprocess_u8(const uint8_t* src, float* dst, int n, int h, int w, int c, const int32_t* flip, const float* crop,
                      void* scratch, gdn_stream stream);

/* ---- demo path: the image resizing of src/depth_extract.py:23-58,86,138 ------------------------
 * scipy.misc.imresize(arr, size, 'bilinear') = bytescale + PIL BILINEAR resize of the 8-bit image, bit-exact.
 * gdn_bytescale: dst[i] = uint8(clip((src[i] - min) * (255 / (max - min or 1)), 0, 255) + 0.5); src is float32 (or
 *   uint8 storage of the float image when src_is_u8); f64_math = 0: float32 arithmetic (NumPy on the float32 input
 *   image, :86), 1: double (the demo copies the depth map into a float64 array first, :135-138);
 *   scratch8 = 8 bytes of device scratch.
 * gdn_resize_u8: src uint8 [n][h][w][c] -> dst uint8 [n][oh][ow][c]; PIL's two-pass (horizontal, then vertical)
 *   antialiased triangle filter with 22-bit fixed-point coefficients (Pillow src/libImaging/Resample.c).  The caller
 *   owns the workspace (coefficient tables + the 8-bit intermediate): gdn_resize_u8_workspace() bytes. */
int gdn_bytescale(const void* src, int src_is_u8, int64_t n, int f64_math, uint8_t* dst, void* scratch8,
                  gdn_stream stream);
size_t gdn_resize_u8_workspace(int n, int h, int w, int c, int oh, int ow);
int gdn_resize_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int c, int oh, int ow, void* workspace,
                  size_t ws_bytes, gdn_stream stream);

/* ---- training loss, metrics, optimizer ------------------------------------------------------------------ */

/* out_max[0] = max(out_max[0], max_i |a[i] - b[i]|)  (caller zeroes; non-negative float compared as bits).
 * Replaces c = 0.2*max|diff| of trainer.py:714 (the 0.2 is applied by gdn_loss). */
int gdn_absdiff_max(const float* a, const float* b, int64_t n, float* out_max, gdn_stream stream);

/* Fused training loss, forward sums + analytic gradient w.r.t. the network output (trainer.py:705-757 RtoD,
 * :433-456 DtoD).  sums[0] += sum w*BerHu(d), sums[1] += sum of the second term (mode 0: |gx|wx + |gy|wy edge-aware
 * smoothness; mode 1: |Sobel_x diff| + |Sobel_y diff|), sums[2] += sum d^2.  The loss value is
 *   mode 0: 3*sums[0]/P + 0.1*sums[1]/P (+ latent)      mode 1: 3*sums[0]/P + 3*sums[1]/P,   P = n*h*w.
 * dout / dpre (optional) receive grad_scale * dL/d(out) and the same chained through tanh (x (1 - out^2)). */
typedef struct {
  const float* out;
  const float* gt;
  const float* sparse;      /* channel 0 of the sparse depth (validity = value > -1), or NULL: no crop/validity weights */
  int64_t sparse_stride;    /* floats between consecutive images of `sparse` */
  const float* rgb;         /* [n][3][h][w] planar, mode 0 only */
  int32_t n, h, w;
  const float* maxabs;      /* device scalar from gdn_absdiff_max (all-reduced MAX by the caller when sharded) */
  int32_t mode;
  double* sums;             /* [3], caller zeroes */
  float* dout;
  float* dpre;
  float grad_scale;
} gdn_loss_desc;
int gdn_loss(const gdn_loss_desc* d, gdn_stream stream);

/* out[0] += sum_i (a[i] - b[i])^2   (feature MSE of the guidance loss, trainer.py:726-733); n % 4 == 0 */
int gdn_sqdiff_sum(const float* a, const float* b, int64_t n, double* out, gdn_stream stream);

/* grad[i] = coef * (a[i] - b[i])  -- d/da of (coef/2) * sum (a - b)^2: one feature term of the latent (guidance) loss
 * when its gradient is enabled; n % 4 == 0 */
int gdn_sqdiff_grad(const float* a, const float* b, int64_t n, float coef, float* grad, gdn_stream stream);
/* dpre[i] += scale * dout[i] * (1 - out[i]^2): adds a gradient w.r.t. the network's tanh output to dL/d(pre-tanh) */
int gdn_tanh_chain_add(const float* dout, const float* out, int64_t n, float scale, float* dpre, gdn_stream stream);

/* calculate_error.compute_errors (calculate_error.py:10-103): per-image min-max normalisation, validity mask + crop,
 * exact lower medians by radix select, median scaling, the 8 Eigen metrics.  Staged over many CTAs per image (min/max,
 * four radix passes that histogram ground truth and prediction together, metric sums finished by the last CTA of each
 * image); the per-pixel fp32 operation order is the reference's, so the delta-threshold COUNTS are bit-exact.
 * out8 += [abs_diff, abs_rel, sq_rel, a1, a2, a3, rmse, rmse_log] averaged over the b images (caller zeroes);
 * counts[b][4] = (n_valid, n(thr<1.25), n(thr<1.25^2), n(thr<1.25^3)) may be NULL.
 * workspace: gdn_depth_metrics_workspace_bytes(b) bytes of caller-owned device memory (8-byte aligned; the library
 * initialises it on the stream). */
size_t gdn_depth_metrics_workspace_bytes(int b);
int gdn_eigen_metrics(const float* gt_np, const float* gt, const float* pred, int b, int h, int w, int crop,
                      double* out8, int64_t* counts, void* workspace, size_t workspace_bytes, gdn_stream stream);

/* The same kernels with the constants of the other two evaluation protocols of calculate_error.py:
 *   GDN_METRICS_KITTI  = compute_errors        (:10-103)  = gdn_eigen_metrics
 *   GDN_METRICS_NYU    = compute_errors_NYU    (:105-151): 10 m range, valid = 0 < gt < 10, border crop, clamp(1e-3, 10);
 *                        out8 = [abs_diff, abs_rel, log10, a1, a2, a3, rmse, rmse_log]; gt_np unused (may be NULL)
 *   GDN_METRICS_MAKE3D = compute_errors_Make3D (:153-182): min-max normalised gt_np / gt / pred, clamp(1e-2, 80) BEFORE
 *                        the median scaling, no crop; out8 = [abs_diff, abs_rel, log10, 0, 0, 0, rmse, 0] */
#define GDN_METRICS_KITTI 0
#define GDN_METRICS_NYU 1
#define GDN_METRICS_MAKE3D 2
int gdn_depth_metrics(int variant, const float* gt_np, const float* gt, const float* pred, int b, int h, int w, int crop,
                      double* out8, int64_t* counts, void* workspace, size_t workspace_bytes, gdn_stream stream);

/* Fused Adam with coupled L2 decay over flat fp32 buffers (torch.optim.Adam semantics); step is 1-based;
 * the gradient is multiplied by grad_scale first. */
int gdn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, gdn_stream stream);

/* same, with the learning rate and the step counter in device memory: dyn[0] = lr, dyn[1] = number of steps taken so
 * far (incremented by this call before use) -- lets a whole training step live in one captured CUDA graph. */
int gdn_adam_step_dyn(float* p, const float* g, float* m, float* v, int64_t n, float* dyn, double beta1, double beta2,
                      float eps, float weight_decay, float grad_scale, gdn_stream stream);

/* ---- the coverage contract's entry-point names -------------------------------------------------
 * One descriptor-driven function implements several contract entries (forward and input-gradient convolutions are
 * the same implicit GEMM on different weight packs; the two losses differ by `mode`); pointers travel in the
 * descriptors.  These names are exported as aliases: gdn_conv2d_fwd = gdn_conv2d_dgrad = gdn_conv2d,
 * gdn_bn_act_apply = gdn_act_forward, gdn_bn_act_apply_bwd = gdn_act_backward (after gdn_bn_bwd_reduce),
 * gdn_loss_rtod_fwd_bwd / gdn_loss_dtod_fwd_bwd = gdn_loss with mode 0 / 1, gdn_feature_mse = gdn_sqdiff_sum,
 * gdn_workspace_bytes = 0 (the convolutions need no global workspace). */
int gdn_conv2d_fwd(const gdn_conv_desc* d, gdn_stream stream);
int gdn_conv2d_dgrad(const gdn_conv_desc* d, gdn_stream stream);
int gdn_bn_act_apply(const gdn_act_fwd_desc* d, gdn_stream stream);
int gdn_bn_act_apply_bwd(const gdn_bn_bwd_desc* d, gdn_stream stream);
int gdn_loss_rtod_fwd_bwd(const gdn_loss_desc* d, gdn_stream stream);
int gdn_loss_dtod_fwd_bwd(const gdn_loss_desc* d, gdn_stream stream);
int gdn_feature_mse(const float* a, const float* b, int64_t n, double* out, gdn_stream stream);
size_t gdn_workspace_bytes(const gdn_conv_desc* d);

const char* gdn_last_error(void);
int gdn_version(void);
int gdn_sm_count(void);
/* 1 when the library runs its fp32 reductions in a fixed order (environment GDN_DETERMINISTIC=1, read once) */
int gdn_deterministic(void);

#ifdef __cplusplus
}
#endif
#endif
