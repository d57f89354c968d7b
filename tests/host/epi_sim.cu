// TEST INFRASTRUCTURE (never linked into libgdn_b200.so): warp-level simulation of the transposed fp32 epilogue of
// conv_igemm_kernel<BN, CG, EPI_T = true> (gdn_pytorch_b200/csrc/conv_igemm.cu: lane_transpose32 + the store loop).
// Shuffles need all 32 lanes at once, so the device code is mirrored here statement by statement with explicit lane
// loops: __shfl_xor_sync(x, s) -> x of lane (l ^ s), __shfl_sync(x, r) -> x of lane r, __ballot_sync -> bit mask.
#include <cmath>
#include <cstddef>

extern "C" int epi_t_host(const float* acc, const unsigned long long* pix, const int* valid, int cout, int cb,
                          const float* bias, int relu, const float* resid, int tanh_out, float* out) {
  float v[32][32];                                   // v[lane][i]: lane = pixel row of the TMEM block, i = column
  for (int l = 0; l < 32; l++)
    for (int i = 0; i < 32; i++) v[l][i] = acc[l * 32 + i];
  // lane_transpose32
  for (int s = 16; s >= 1; s >>= 1)
    for (int i = 0; i < 32; i++) {
      if (i & s) continue;
      float x[32], y[32];
      for (int l = 0; l < 32; l++) x[l] = (l & s) ? v[l][i] : v[l][i | s];
      for (int l = 0; l < 32; l++) y[l] = x[l ^ s];              // __shfl_xor_sync
      for (int l = 0; l < 32; l++) {
        if (l & s) v[l][i] = y[l]; else v[l][i | s] = y[l];
      }
    }
  unsigned vmask = 0;
  for (int l = 0; l < 32; l++)
    if (valid[l]) vmask |= 1u << l;                              // __ballot_sync
  for (int r = 0; r < 32; r++) {
    const size_t pr = (size_t)pix[r];                            // __shfl_sync(hi / lo, r)
    if (!((vmask >> r) & 1u)) continue;
    for (int l = 0; l < 32; l++) {
      float x = v[l][r] + (bias ? bias[cb + l] : 0.f);
      if (relu) x = fmaxf(x, 0.f);
      const size_t o = pr * cout + cb + l;
      if (resid) x += resid[o];
      if (tanh_out) x = tanhf(x);
      out[o] = x;
    }
  }
  return 0;
}
