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
 *   gdn_act_forward       nn.BatchNorm2d apply + nn.ReLU + residual add + nn.ReflectionPad2d + F.interpolate(x2):
 *                         src/AE_model_unet.py:51-57,66-69,86-87,135,336-355
 *   gdn_bn_finalize       batch statistics -> scale/shift, running-stat update (nn.BatchNorm2d training mode)
 *   gdn_bn_bwd_reduce / gdn_act_backward   autograd of the above
 *   gdn_fold_grad         adjoint of reflection padding / bilinear x2 upsampling / zero-dilation
 *   gdn_loss_rtod / gdn_loss_dtod / gdn_feature_mse   src/trainer.py:433-456, 705-757; src/utils.py:105-178
 *   gdn_eigen_metrics     src/calculate_error.py:10-103
 *   gdn_adam_step         optim.Adam(lr, [0.9,0.999], eps=1e-8, weight_decay=5e-4): src/GDN_main.py:157,173
 *   gdn_pack_weights / gdn_unpack_wgrad   fp32 OIHW <-> bf16 [tap][Cout][Cin] operand layout (BN folding at eval)
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
  int32_t algo;              /* GDN_CONV_* */
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
} gdn_conv_desc;

int gdn_conv2d(const gdn_conv_desc* d, gdn_stream stream);

/* Weight gradient: dw[tap][co][ci] (fp32, += ) = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*stride + r + off_y, ox*stride + s + off_x, ci] */
typedef struct {
  gdn_act x0, x1;            /* forward input(s) of the convolution (virtual concat like gdn_conv_desc) */
  gdn_act dy;                /* gradient of the conv output, interior = out_h x out_w, c = cout_pad */
  float* dw;                 /* fp32 [kh*kw][cout_pad][cin_total], accumulated into (caller zeroes) */
  int32_t kh, kw, stride, off_y, off_x, out_h, out_w, cout_pad;
} gdn_wgrad_desc;

int gdn_conv2d_wgrad(const gdn_wgrad_desc* d, gdn_stream stream);

const char* gdn_last_error(void);
int gdn_version(void);
int gdn_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif
