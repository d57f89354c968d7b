// TEST INFRASTRUCTURE (never linked into libgdn_b200.so): the two phases of the wide-tile weight re-pack
// (gdn_pytorch_b200/csrc/pack_tile.cuh) run on the host, CTA by CTA and thread by thread, with the kernel's own grid.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../gdn_pytorch_b200/csrc/pack_tile.cuh"

using namespace gdn;

extern "C" int pack_v2_host(int kh, int kw, int A, int B, int Apad, int Bpad, long long sa, long long sb, long long sr,
                            long long ss, int flip, const float* w, const float* scale_a, uint16_t* out, int nthreads) {
  PackK k{kh, kw, A, B, Apad, Bpad, sa, sb, sr, ss, flip, 0};
  const int T = kh * kw;
  if (!(ss == 1 && (sr == kw || kh == 1))) return -2;                       // pack_tileable()
  if (!((sb == T && sa >= (long long)B * T) || (sa == T && sb >= (long long)A * T))) return -2;
  const int tb = pack_tb_for_taps(T);
  const int tiles_x = (Bpad + tb - 1) / tb, tiles_y = (Apad + kPackTA - 1) / kPackTA;
  std::vector<float> s_tile(pack_smem_bytes(T) / sizeof(float));
  for (int ty = 0; ty < tiles_y; ty++)
    for (int tx = 0; tx < tiles_x; tx++) {
      const int a0 = ty * kPackTA, b0 = tx * tb;
      std::fill(s_tile.begin(), s_tile.end(), -12345.f);                    // poison: phase 2 must only read what phase 1 wrote
      for (int tid = 0; tid < nthreads; tid++) pack_v2_phase1(w, k, a0, b0, tb, s_tile.data(), tid, nthreads);
      for (int tid = 0; tid < nthreads; tid++)
        pack_v2_phase2(scale_a, reinterpret_cast<__nv_bfloat16*>(out), k, a0, b0, tb, s_tile.data(), tid, nthreads);
    }
  return tb;
}
