// Micro-benchmark 3: the K-I2T k-block issue pattern (12 N=64 MMAs into S, then 4 x {hi, lo} N=16 residual MMAs
// into four different column blocks of O), against variants, to find what a change of accumulator costs.
#include "common.cuh"
#include <cstdio>
using namespace csam;

__global__ void __launch_bounds__(128, 1) pat_kernel(int variant, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t id64 = umma_idesc_f16(128, 64, 0, 0), id16 = umma_idesc_f16(128, 16, 0, 0);
      const uint32_t ah = umma_desc_lo(smem_u32(smem), 16), al = umma_desc_lo(smem_u32(smem + 16384), 16);
      const uint32_t bh = umma_desc_lo(smem_u32(smem + 65536), 16), bl = umma_desc_lo(smem_u32(smem + 65536 + 8192), 16);
      const uint32_t idd = umma_desc_lo(smem_u32(smem + 98304), 16);
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          if (variant < 2 || variant == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_w(tm, ah + 2 * k, bh + 2 * k, id64, 1u);
              umma_f16_w(tm, al + 2 * k, bh + 2 * k, id64, 1u);
              umma_f16_w(tm, ah + 2 * k, bl + 2 * k, id64, 1u);
            }
          }
          if (variant == 1 || variant == 2) {          // residual, N = 16, four accumulator blocks
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t dr = tm + 128 + (i & 3) * 64 + k * 16;
              umma_f16_w(dr, ah + 2 * k, idd + 130 * k, id16, 0u);
              umma_f16_w(dr, al + 2 * k, idd + 130 * k, id16, 1u);
            }
          }
          if (variant == 3) {                          // residual, N = 64, one accumulator block
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t dr = tm + 128 + (i & 3) * 64;
              umma_f16_w(dr, ah + 2 * k, idd + 2 * k, id64, k ? 1u : 0u);
              umma_f16_w(dr, al + 2 * k, idd + 2 * k, id64, 1u);
            }
          }
          if (variant == 5) {                          // 12 x N=64, A_hi used by two CONSECUTIVE MMAs (operand reuse?)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_w(tm, ah + 2 * k, bh + 2 * k, id64, 1u);
              umma_f16_w(tm, ah + 2 * k, bl + 2 * k, id64, 1u);
              umma_f16_w(tm, al + 2 * k, bh + 2 * k, id64, 1u);
            }
          }
          if (variant == 6) {                          // 8 MMAs: A_hi x [B_hi | B_lo] as N = 128, then A_lo x B_hi (N = 64)
            constexpr uint32_t id128 = umma_idesc_f16(128, 128, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_w(tm, ah + 2 * k, bh + 2 * k, id128, 1u);
              umma_f16_w(tm, al + 2 * k, bh + 2 * k, id64, 1u);
            }
          }
          if (variant == 4) {                          // S alternating between two buffers every MMA
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_w(tm + 64 * (k & 1), ah + 2 * k, bh + 2 * k, id64, 1u);
              umma_f16_w(tm + 64 * ((k + 1) & 1), al + 2 * k, bh + 2 * k, id64, 1u);
            }
          }
        }
        umma_commit(&bar);
        mbar_wait(&bar, rep & 1);
        const long long t1 = clock64();
        if (rep == 1 && blockIdx.x == 0) out[0] = t1 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

int main() {
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(pat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512;
  const char* names[] = {"12 x N=64 (S only)", "12 x N=64 + 4 x {2 x N=16} residual", "4 x {2 x N=16} residual only",
                         "12 x N=64 + 8 x N=64 residual (one block)", "8 x N=64 alternating two accumulators",
                         "12 x N=64, hh hl lh order (A_hi twice in a row)", "4 x {N=128 [Bh|Bl], N=64} = the same products in 8 MMAs"};
  for (int v = 0; v < 7; ++v) {
    pat_kernel<<<148, 128, 200 * 1024>>>(v, iters, out);
    long long t = 0; cudaMemcpy(&t, out, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("%-46s %7.1f clk per k-block %s\n", names[v], (double)t / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
