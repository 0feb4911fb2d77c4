// Micro-benchmark: tcgen05.mma issue/execute rate by shape, operand majorness and accumulator reuse (SS mode),
// one CTA per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I crowdsam_b200/csrc scripts/micro/mma_rate.cu -o mma_rate
#include "common.cuh"
#include <cstdio>
using namespace csam;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int a_mn, int b_mn, int nd, int a_tmem, int iters, long long* out) {
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
      const uint32_t idesc = umma_idesc_f16(128, N, a_mn, b_mn);
      const uint32_t ad = umma_desc_lo(smem_u32(smem), a_mn ? 16384 : 16);
      const uint32_t bd = umma_desc_lo(smem_u32(smem + 65536), b_mn ? 16384 : 16);
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = tm + ((i * 4 + k) % nd) * 256 / (nd > 1 ? nd : 1) * 0 + (nd > 1 ? ((i * 4 + k) % nd) * N : 0);
            const uint32_t step = (a_mn ? 128 : 2) * k, bstep = (b_mn ? 128 : 2) * k;
            if (a_tmem) umma_f16_ts(d, tm + 384 + 8 * k, bd + bstep, idesc, 1u);
            else umma_f16_w(d, ad + step, bd + bstep, idesc, 1u);
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
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 256;   // x4 MMAs
  struct { int N, a_mn, b_mn, nd, a_tmem; const char* name; } cfg[] = {
    {64, 0, 0, 1, 0, "N=64  K-major A,B  one D"},   {64, 0, 0, 2, 0, "N=64  K-major A,B  two D"},
    {64, 1, 1, 1, 0, "N=64  MN-major A,B one D"},   {64, 0, 0, 1, 1, "N=64  A in TMEM    one D"},
    {128, 0, 0, 1, 0, "N=128 K-major A,B  one D"},  {128, 0, 0, 1, 1, "N=128 A in TMEM    one D"},
    {256, 0, 0, 1, 0, "N=256 K-major A,B  one D"},  {256, 0, 0, 1, 1, "N=256 A in TMEM    one D"},
    {16, 0, 0, 1, 0, "N=16  K-major A,B  one D"},   {32, 0, 0, 1, 0, "N=32  K-major A,B  one D"},
  };
  for (int grid : {1, 148}) {
    for (auto& c : cfg) {
      rate_kernel<<<grid, 128, 200 * 1024>>>(c.N, c.a_mn, c.b_mn, c.nd, c.a_tmem, iters, out);
      long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("grid %3d  %-28s %7.1f clk / MMA  (floor %d)%s\n", grid, c.name, (double)h / (iters * 4), 128 * c.N / 256,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
