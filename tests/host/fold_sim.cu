// TEST INFRASTRUCTURE (never linked into libgdn_b200.so): runs the index arithmetic of fold_rows2_kernel
// (gdn_pytorch_b200/csrc/fold_rows.cuh, shared __host__ __device__ pieces) on the host, CTA by CTA and thread by
// thread in the kernel's own loop structure, so that tests/test_host_sim_cpu.py can check it against the autograd
// adjoint without a GPU.  Built on demand by the test with `nvcc -shared` (host code only).
#include <vector>
#include "../../gdn_pytorch_b200/csrc/fold_rows.cuh"

using namespace gdn;

extern "C" int fold_rows2_host(const float* dpad, int ctot, int c_off, int N, int H, int W, int C, int P, int reflect, int up,
                               int dilate, float* dact, int accumulate, int nthreads, int nblocks) {
  FoldK f{dpad, ctot, c_off, N, H, W, C, P, reflect, up, dilate, dact, accumulate, 0};
  int lg_cg = -1;
  for (int l = 0; l < 16; l++)
    if ((1 << l) == C / 4) lg_cg = l;
  if (lg_cg < 0 || C % 4) return -1;
  const int OH = (up || dilate) ? 2 * H : H, OW = (up || dilate) ? 2 * W : W;
  const int Hq = OH + 2 * P, Wq = OW + 2 * P;
  const int rows = N * H, items = W << lg_cg;
  for (int b = 0; b < nblocks; b++) {                      // one simulated CTA
    std::vector<int> s_pc((size_t)W * kFoldColInts);
    std::vector<float> s_qw((size_t)W * 4);
    for (int tid = 0; tid < nthreads; tid++)
      for (int x = tid; x < W; x += nthreads) fold_col_entry(f, OW, x, s_pc.data() + (size_t)x * kFoldColInts, s_qw.data() + (size_t)x * 4);
    for (int row = b; row < rows; row += nblocks) {
      int s_prow[12];
      float s_pw[12];
      const int np = fold_row_entry(f, OH, row % H, s_prow, s_pw);   // thread 0 between the two barriers
      for (int tid = 0; tid < nthreads; tid++)
        for (int it = tid; it < items; it += nthreads) fold_item(f, lg_cg, Hq, Wq, row, it, np, s_prow, s_pw, s_pc.data(), s_qw.data());
    }
  }
  return 0;
}
