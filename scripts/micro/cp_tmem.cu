// Micro-benchmark 2: tcgen05.cp (shared -> tensor memory) of one K = 16 step of a K-major, 128B-swizzled A tile,
// (a) layout check against tcgen05.ld, (b) rate of [cp hi, cp lo, 3 x MMA with A from TMEM] against 3 x SS MMAs.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace csam;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint32_t desc_lo) {
  asm volatile("{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}"
               ::"r"(taddr), "r"(desc_lo), "r"(UMMA_DESC_HI_SW128) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

// mode 0: layout check.  smem A tile: 128 rows x 64 fp16 (K-major, SW128), element (r, k) = r * 64 + k as 16-bit int.
__global__ void __launch_bounds__(128, 1) check_kernel(uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const int chunk = (k >> 3) ^ (r & 7);
    reinterpret_cast<uint16_t*>(smem)[r * 64 + chunk * 8 + (k & 7)] = (uint16_t)(r * 64 + k);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<64>(&slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t ad = umma_desc_lo(smem_u32(smem), 16);
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tm + 8 * k, ad + 2 * k);
      umma_commit(&bar);
    }
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[8];
  for (int k = 0; k < 4; ++k) {
    tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + 8 * k, r);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[(threadIdx.x * 4 + k) * 8 + j] = r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tm);
}

// mode 1: rates.  variant 0 = SS x3, 1 = cp + TS x3 (double-buffered A in TMEM), 2 = TS x3 only (no cp)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int variant, int iters, long long* out) {
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
      const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
      const uint32_t ah = umma_desc_lo(smem_u32(smem), 16), al = umma_desc_lo(smem_u32(smem + 16384), 16);
      const uint32_t bh = umma_desc_lo(smem_u32(smem + 65536), 16), bl = umma_desc_lo(smem_u32(smem + 65536 + 32768), 16);
      const uint32_t ta = tm + 256;      // A staging: [buf][hi 8 cols | lo 8 cols]
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (variant == 0) {
              umma_f16_w(tm, ah + 2 * k, bh + 2 * k, idesc, 1u);
              umma_f16_w(tm, al + 2 * k, bh + 2 * k, idesc, 1u);
              umma_f16_w(tm, ah + 2 * k, bl + 2 * k, idesc, 1u);
            } else {
              const uint32_t tb = ta + (k & 1) * 16;
              if (variant == 1) { tmem_cp_128x256b(tb, ah + 2 * k); tmem_cp_128x256b(tb + 8, al + 2 * k); }
              umma_f16_ts(tm, tb, bh + 2 * k, idesc, 1u);
              umma_f16_ts(tm, tb + 8, bh + 2 * k, idesc, 1u);
              umma_f16_ts(tm, tb, bl + 2 * k, idesc, 1u);
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
  uint32_t* dout; cudaMalloc(&dout, 128 * 32 * 4);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  check_kernel<<<1, 128, 32 * 1024>>>(dout);
  uint32_t* h = (uint32_t*)malloc(128 * 32 * 4);
  cudaError_t e = cudaMemcpy(h, dout, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  printf("check: %s\n", cudaGetErrorString(e));
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 32; ++c) {
      const uint32_t want = (uint32_t)(r * 64 + 2 * c) | ((uint32_t)(r * 64 + 2 * c + 1) << 16);
      if (h[r * 32 + c] != want && bad++ < 8) printf("  row %d col %d: got %08x want %08x\n", r, c, h[r * 32 + c], want);
    }
  printf("tcgen05.cp.128x256b of a K-major SW128 tile -> lane = row, 8 columns per K step: %s (%d mismatches)\n", bad ? "MISMATCH" : "OK", bad);
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 256;
  for (int N : {64, 128, 256})
    for (int v = 0; v < 3; ++v) {
      rate_kernel<<<148, 128, 200 * 1024>>>(N, v, iters, out);
      long long t = 0; cudaMemcpy(&t, out, 8, cudaMemcpyDeviceToHost);
      e = cudaGetLastError();
      printf("N=%3d  %-34s %7.1f clk per k-step (3 MMAs), floor %d %s\n", N,
             v == 0 ? "SS x3" : v == 1 ? "cp hi + cp lo + TS x3" : "TS x3 (no cp)", (double)t / (iters * 4), 3 * 128 * N / 256,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
