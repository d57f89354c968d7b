// Hardware probe for the descriptor semantics the conv kernels rely on (run on a B200 via gpurun).
// Checks, against a host model:
//   * TMA SWIZZLE_128B smem image (2D and 4D boxes, negative/OOB coordinates -> zero fill)
//   * tcgen05.mma K-major operands whose start address is a 128 B multiple that is NOT 1024 B aligned
//     (shifted windows into a resident halo tile) and whose 8-row groups are SBO != 1024 apart
//   * tcgen05.mma MN-major operands (wgrad), incl. LBO = 128 B (two taps stacked along M)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../gdn_pytorch_b200/csrc/sm100_ptx.cuh"

using namespace gdn;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeTiledFn)fn;
}

__device__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) return false;
  }
  return true;
}

struct Probe {
  // operand descriptors
  uint32_t a_off, a_lbo, a_sbo, a_base, a_mn;   // byte offsets relative to smem A base
  uint32_t b_off, b_lbo, b_sbo, b_base, b_mn;
  uint32_t a_kstep, b_kstep;                    // bytes added to the start address per K=16 step
  uint32_t M, N, ksteps;
  // loads: up to 2 A loads + 2 B loads via 2D maps, or one 4D A load
  int a_rows[2], b_rows[2];                     // rows per 2D load (0 = none)
  int use4d; int c4[4]; int rows4d;
  int dump_a_bytes;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapA4, Probe p, float* __restrict__ D, uint8_t* __restrict__ dumpA) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 96 KB
  uint8_t* sB = smem + 96 * 1024;     // 64 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (p.use4d) bytes += p.rows4d * 128;
    for (int i = 0; i < 2; i++) bytes += (p.a_rows[i] + p.b_rows[i]) * 128;
    mbar_expect_tx(&bar_load, bytes);
    if (p.use4d) tma_load_4d(&mapA4, &bar_load, sA, p.c4[0], p.c4[1], p.c4[2], p.c4[3]);
    int ra = 0, rb = 0;
    for (int i = 0; i < 2; i++) {
      if (p.a_rows[i]) { tma_load_2d(&mapA, &bar_load, sA + ra * 128, 0, ra); ra += p.a_rows[i]; }
      if (p.b_rows[i]) { tma_load_2d(&mapB, &bar_load, sB + rb * 128, 0, rb); rb += p.b_rows[i]; }
    }
    if (!wait_bounded(&bar_load, 0)) { printf("TIMEOUT waiting for TMA\n"); __trap(); }
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(p.M, p.N, p.a_mn, p.b_mn);
    for (uint32_t k = 0; k < p.ksteps; k++) {
      uint64_t ad = make_smem_desc_sw128(smem_u32(sA) + p.a_off + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_base);
      uint64_t bd = make_smem_desc_sw128(smem_u32(sB) + p.b_off + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_base);
      umma_bf16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  if (!wait_bounded(&bar_mma, 0)) { if (threadIdx.x == 1) printf("TIMEOUT waiting for MMA\n"); __trap(); }
  tc_fence_after();
  for (uint32_t c = 0; c < p.N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++) D[(warp * 32 + lane) * 256 + c + j] = __uint_as_float(r[j]);
  }
  for (int i = threadIdx.x; i < p.dump_a_bytes; i += 128) dumpA[i] = sA[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  EncodeTiledFn enc = get_encode();
  const int ROWS = 768;  // source rows (64 bf16 each)
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(ROWS * 64);
  std::vector<float> fA(ROWS * 64), fB(ROWS * 64);
  srand(1);
  for (int i = 0; i < ROWS * 64; i++) {
    fA[i] = bf((rand() % 2001 - 1000) / 1000.f);
    fB[i] = bf((rand() % 2001 - 1000) / 1000.f);
    hA[i] = __float2bfloat16(fA[i]);
    hB[i] = __float2bfloat16(fB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  uint8_t* dDump;
  CK(cudaMalloc(&dA, ROWS * 128));
  CK(cudaMalloc(&dB, ROWS * 128));
  CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMalloc(&dDump, 96 * 1024));
  CK(cudaMemcpy(dA, hA.data(), ROWS * 128, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), ROWS * 128, cudaMemcpyHostToDevice));

  auto make2d = [&](void* ptr, int boxrows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, (cuuint64_t)ROWS};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode2d failed %d\n", (int)r); exit(1); }
    return m;
  };
  // 4D view of A source: [n=2][h=12][w=32][c=64]  (768 rows)
  const int N4 = 2, H4 = 12, W4 = 32;
  auto make4d = [&](int bw, int bh, int bn) {
    CUtensorMap m;
    cuuint64_t dims[4] = {64, (cuuint64_t)W4, (cuuint64_t)H4, (cuuint64_t)N4};
    cuuint64_t strides[3] = {128, 128ull * W4, 128ull * W4 * H4};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode4d failed %d\n", (int)r); exit(1); }
    return m;
  };
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024 + 1024));

  std::vector<float> hD(128 * 256);
  std::vector<uint8_t> hDump(96 * 1024);

  // logical element accessors given a probe (host model)
  auto run = [&](const char* name, Probe p, CUtensorMap mA, CUtensorMap mB, CUtensorMap mA4,
                 auto Aelem /*(m,k)->float*/, auto Belem /*(n,k)->float*/) {
    CK(cudaMemset(dD, 0, 128 * 256 * 4));
    probe_kernel<<<1, 128, 161 * 1024 + 1024>>>(mA, mB, mA4, p, dD, dDump);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("[%s] KERNEL ERROR %s\n", name, cudaGetErrorString(e));
      cudaDeviceReset();
      exit(2);
    }
    CK(cudaMemcpy(hD.data(), dD, 128 * 256 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hDump.data(), dDump, 96 * 1024, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    const int K = p.ksteps * 16;
    for (uint32_t m = 0; m < p.M; m++)
      for (uint32_t n = 0; n < p.N; n++) {
        double ref = 0;
        for (int k = 0; k < K; k++) ref += (double)Aelem(m, k) * (double)Belem(n, k);
        double err = fabs(ref - hD[m * 256 + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-2) bad++;
      }
    printf("[%s] M=%u N=%u K=%d maxerr=%.5f bad=%d %s\n", name, p.M, p.N, K, maxerr, bad, bad ? "FAIL" : "ok");
    return bad == 0;
  };

  CUtensorMap mA128 = make2d(dA, 128), mB64 = make2d(dB, 64), mB128 = make2d(dB, 128), mA256 = make2d(dA, 256),
              mB256 = make2d(dB, 256);
  CUtensorMap m4 = make4d(8, 8, 2);

  // ---- T0: TMA swizzle image check (2D, 16 rows dumped)
  {
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_sbo = 1024; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.a_rows[0] = 128; p.b_rows[0] = 64; p.dump_a_bytes = 128 * 128;
    run("T1 basic K-major", p, mA128, mB64, m4, [&](int m, int k) { return fA[m * 64 + k]; },
        [&](int n, int k) { return fB[n * 64 + k]; });
    int mism = 0;
    for (int r = 0; r < 128; r++)
      for (int c = 0; c < 8; c++) {
        const uint8_t* phys = &hDump[r * 128 + ((c ^ (r & 7)) * 16)];
        if (memcmp(phys, &hA[r * 64 + c * 8], 16) != 0) mism++;
      }
    printf("[T0 swizzle image: chunk c of row r at (c ^ (r&7))] mismatches=%d %s\n", mism, mism ? "FAIL" : "ok");
  }
  // ---- T2: shifted start (row offset j), base_offset 0 / j
  for (int j : {1, 3, 7}) {
    for (int usebase = 0; usebase < 2; usebase++) {
      Probe p{};
      p.M = 128; p.N = 64; p.ksteps = 4; p.a_sbo = 1024; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
      p.a_rows[0] = 256; p.b_rows[0] = 64; p.a_off = j * 128; p.a_base = usebase ? j : 0;
      char nm[64]; snprintf(nm, 64, "T2 A start row %d base_off=%d", j, p.a_base);
      run(nm, p, mA256, mB64, m4, [&](int m, int k) { return fA[(m + j) * 64 + k]; },
          [&](int n, int k) { return fB[n * 64 + k]; });
    }
  }
  // ---- T3: SBO = 16 rows (2048 B) with start rows 0, 3 : logical row m -> src row start + (m/8)*16 + m%8
  for (int j : {0, 3, 11}) {
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_sbo = 2048; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.a_rows[0] = 256; p.a_rows[1] = 256; p.b_rows[0] = 64; p.a_off = j * 128;
    char nm[64]; snprintf(nm, 64, "T3 A SBO=2048 start row %d", j);
    run(nm, p, mA256, mB64, m4, [&](int m, int k) { return fA[(j + (m / 8) * 16 + m % 8) * 64 + k]; },
        [&](int n, int k) { return fB[n * 64 + k]; });
  }
  // ---- T3b: SBO = 24 rows (3072 B, not a power of two), start row 5
  {
    int j = 5;
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_sbo = 3072; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.a_rows[0] = 256; p.a_rows[1] = 256; p.b_rows[0] = 64; p.a_off = j * 128;
    run("T3b A SBO=3072 start row 5", p, mA256, mB64, m4,
        [&](int m, int k) { return fA[(j + (m / 8) * 24 + m % 8) * 64 + k]; },
        [&](int n, int k) { return fB[n * 64 + k]; });
  }
  // ---- T4: MN-major A (M=128 = 2 x 64 cols, second half LBO = 128 rows away), MN-major B (N=64). K = 64 rows.
  {
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_mn = 1; p.b_mn = 1;
    p.a_lbo = 128 * 128; p.a_sbo = 1024; p.b_lbo = 0; p.b_sbo = 1024;
    p.a_kstep = 2048; p.b_kstep = 2048;  // 16 K rows per step
    p.a_rows[0] = 256; p.b_rows[0] = 64;
    run("T4 MN-major A(LBO=16K) B", p, mA256, mB64, m4,
        [&](int m, int k) { return fA[((m / 64) * 128 + k) * 64 + m % 64]; },
        [&](int n, int k) { return fB[k * 64 + n]; });
  }
  // ---- T5: MN-major A with LBO = 128 B (two taps one pixel apart), start row 3, SBO = 16 rows
  {
    int j = 3;
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_mn = 1; p.b_mn = 1;
    p.a_lbo = 128; p.a_sbo = 2048; p.b_sbo = 1024; p.a_off = j * 128;
    p.a_kstep = 4096; p.b_kstep = 2048;
    p.a_rows[0] = 256; p.b_rows[0] = 64;
    run("T5 MN-major A LBO=128B SBO=2048 start 3", p, mA256, mB64, m4,
        [&](int m, int k) { return fA[(j + (m / 64) + (k / 8) * 16 + k % 8) * 64 + m % 64]; },
        [&](int n, int k) { return fB[k * 64 + n]; });
  }
  // ---- T6: MN-major B with N = 256 (4 x 64, LBO = 64 rows = 8192 B), A K-major? no: A MN-major M=128
  {
    Probe p{};
    p.M = 128; p.N = 256; p.ksteps = 4; p.a_mn = 1; p.b_mn = 1;
    p.a_lbo = 64 * 128; p.a_sbo = 1024; p.b_lbo = 64 * 128; p.b_sbo = 1024;
    p.a_kstep = 2048; p.b_kstep = 2048;
    p.a_rows[0] = 128; p.b_rows[0] = 256;
    run("T6 MN-major N=256", p, mA128, mB256, m4,
        [&](int m, int k) { return fA[((m / 64) * 64 + k) * 64 + m % 64]; },
        [&](int n, int k) { return fB[((n / 64) * 64 + k) * 64 + n % 64]; });
  }
  // ---- T7: K-major N=256 B (256 rows), A K-major
  {
    Probe p{};
    p.M = 128; p.N = 256; p.ksteps = 4; p.a_sbo = 1024; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.a_rows[0] = 128; p.b_rows[0] = 256;
    run("T7 K-major N=256", p, mA128, mB256, m4, [&](int m, int k) { return fA[m * 64 + k]; },
        [&](int n, int k) { return fB[n * 64 + k]; });
  }
  // ---- T8: 4D TMA box {64, 8, 8, 2} at (0, -2, -1, 0): smem row m = (nb*8 + ty)*8 + tx ; OOB -> 0
  {
    Probe p{};
    p.M = 128; p.N = 64; p.ksteps = 4; p.a_sbo = 1024; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.use4d = 1; p.rows4d = 128; p.c4[0] = 0; p.c4[1] = -2; p.c4[2] = -1; p.c4[3] = 0;
    p.b_rows[0] = 64;
    auto A4 = [&](int m, int k) -> float {
      int tx = m % 8, ty = (m / 8) % 8, nb = m / 64;
      int x = tx - 2, y = ty - 1;
      if (x < 0 || y < 0 || x >= W4 || y >= H4) return 0.f;
      return fA[((nb * H4 + y) * W4 + x) * 64 + k];
    };
    run("T8 4D box negative coords", p, mA128, mB64, m4, A4, [&](int n, int k) { return fB[n * 64 + k]; });
    p.c4[1] = 28; p.c4[2] = 8; p.c4[3] = 1;  // runs off the far edges and past the last image
    auto A4b = [&](int m, int k) -> float {
      int tx = m % 8, ty = (m / 8) % 8, nb = m / 64 + 1;
      int x = tx + 28, y = ty + 8;
      if (x >= W4 || y >= H4 || nb >= N4) return 0.f;
      return fA[((nb * H4 + y) * W4 + x) * 64 + k];
    };
    run("T8b 4D box far-edge OOB", p, mA128, mB64, m4, A4b, [&](int n, int k) { return fB[n * 64 + k]; });
  }
  // ---- T9: N=16 (Cout=1 head padded), N=128
  for (int n : {16, 128}) {
    Probe p{};
    p.M = 128; p.N = n; p.ksteps = 4; p.a_sbo = 1024; p.b_sbo = 1024; p.a_kstep = 32; p.b_kstep = 32;
    p.a_rows[0] = 128; p.b_rows[0] = 128;
    char nm[32]; snprintf(nm, 32, "T9 K-major N=%d", n);
    run(nm, p, mA128, mB128, m4, [&](int m, int k) { return fA[m * 64 + k]; },
        [&](int nn, int k) { return fB[nn * 64 + k]; });
  }
  printf("probe done\n");
  return 0;
}
