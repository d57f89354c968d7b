// tcgen05.mma issue-rate probe: how many SM cycles does one M=128 x N x K=16 BF16 MMA cost in SS mode
// (both operands from shared memory) for N = 64 / 128 / 256, with aligned and shifted-window A descriptors?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_mma_rate tools/probe_mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../gdn_pytorch_b200/csrc/sm100_ptx.cuh"
using namespace gdn;

__global__ void __launch_bounds__(128, 1)
rate_kernel(int N, int iters, uint32_t a_sbo, uint32_t a_shift, int nacc, int mn_major, long long* out, int whole_warp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (whole_warp && threadIdx.x < 32) {
    // the WHOLE warp runs the (warp-uniform) loop; one elected lane issues a whole tap's MMAs at once, with
    // descriptors formed as 64-bit base + immediate so that everything stays in uniform registers
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 128 * 1024;
    const uint64_t a_hi = make_smem_desc_sw128(0, mn_major ? 128 : 0, a_sbo);
    const uint64_t b_hi = make_smem_desc_sw128(0, 8192, 1024);
    const uint32_t tb = tmem_base;
    const uint32_t accstep = (nacc > 1) ? (uint32_t)N : 0u;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t tap = (it % 9) * 128u * a_shift;
      const uint64_t a0 = a_hi + ((sA + tap) >> 4);
      const uint64_t b0 = b_hi + (sB >> 4);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) umma_bf16(tb + j * accstep, a0 + (j * 64 + k4 * 2), b0 + k4 * 2, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  } else if (!whole_warp && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mn_major, mn_major);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 128 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t tap = (it % 9) * 128u * a_shift;
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
          uint64_t ad = make_smem_desc_sw128(sA + tap + j * 1024 + k4 * 32, mn_major ? 128 : 0, a_sbo);
          uint64_t bd = make_smem_desc_sw128(sB + k4 * 32, 8192, 1024);
          umma_bf16(tmem_base + (j % nacc) * N, ad, bd, idesc, 1);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
  const int iters = 2000;
  struct Cfg { int N; uint32_t sbo, shift; int nacc; int mn; const char* name; };
  Cfg cfgs[] = {
      {64, 1024, 0, 4, 0, "N=64  aligned A, 4 accumulators"},
      {64, 5120, 1, 4, 0, "N=64  halo window (SBO 5120, shifted), 4 acc"},
      {64, 5120, 1, 1, 0, "N=64  halo window, 1 accumulator (dependent)"},
      {128, 1024, 0, 2, 0, "N=128 aligned A"},
      {128, 2816, 1, 2, 0, "N=128 halo window (SBO 2816)"},
      {256, 1024, 0, 2, 0, "N=256 aligned A"},
      {256, 2560, 1, 2, 0, "N=256 halo window (SBO 2560)"},
      {64, 3072, 1, 4, 1, "N=64  MN-major both (wgrad), LBO 128"},
      {16, 5120, 1, 4, 0, "N=16  halo window (head)"},
  };
  for (auto& c : cfgs) {
    for (int ww = 0; ww < 2; ww++) {
      const int grid = 148;
      rate_kernel<<<grid, 128, 201 * 1024 + 1024>>>(c.N, iters, c.sbo, c.shift, c.nacc, c.mn, d, ww);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < grid; i++) avg += (double)h[i];
      avg /= grid;
      const double per = avg / (iters * 16.0);
      printf("%-48s %s: %7.1f cycles / MMA  (ideal %5.1f) -> %.0f%% of the tensor pipe\n", c.name, ww ? "warp+elect" : "lane0     ", per,
             128.0 * c.N / 256.0, 100.0 * (128.0 * c.N / 256.0) / per);
    }
  }
  return 0;
}
