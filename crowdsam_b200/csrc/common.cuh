// common.cuh — shared helpers for libcsam_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/csam.h"

namespace csam {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(const char* fmt, const char* a = "", long long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return 1;
}
inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
// ---- per-device state ---------------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per DEVICE: a process that calls the
// library on a second GPU must opt in again there.  One bit per device ordinal in a per-call-site atomic.
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
template <typename Kern>
inline int ensure_dyn_smem(Kern kern, int bytes, std::atomic<unsigned long long>& done, const char* what) {
  const int dev = current_device();
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(max dynamic smem = %d) failed for %s on device %d", bytes, what, dev);
    return 1;
  }
  done.fetch_or(bit, std::memory_order_release);
  return 0;
}
#define CSAM_DYN_SMEM(kern, bytes, what)                                    \
  do {                                                                      \
    static std::atomic<unsigned long long> csam_done_{0};                   \
    if (::csam::ensure_dyn_smem(kern, bytes, csam_done_, what)) return 1;   \
  } while (0)
int num_sms();   // SM count of the current device (gemm.cu)

#define CSAM_REQUIRE(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) return ::csam::fail("%s (%lld)", msg, __LINE__); \
  } while (0)

// ---- h16 pair (hi + lo split) ---------------------------------------------------------
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  v = clamp_h(v);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ void store_pair(__half* hi, __half* lo, size_t i, float v) {
  __half h, l;
  split_h(v, h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}
__device__ __forceinline__ float load_pair(const __half* hi, const __half* lo, size_t i) {
  float v = __half2float(hi[i]);
  if (lo) v += __half2float(lo[i]);
  return v;
}
// Packed split of two values: hi = fp16x2(a,b), lo = fp16x2(a - hi.x, b - hi.y).  One F2FP per pair instead of
// one conversion per element; the subtraction is exact in fp32.  Inputs are clamped to the fp16 range.
__device__ __forceinline__ void split_h2(float a, float b, __half2& hi, __half2& lo) {
  a = clamp_h(a);
  b = clamp_h(b);
  hi = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(a - hf.x, b - hf.y);
}
// same without the fp16-range clamp, for values known to be small (probabilities, LayerNorm outputs): saves two
// FMNMX per element in epilogues that are bound by their instruction count
__device__ __forceinline__ void split_h2_nc(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(a - hf.x, b - hf.y);
}
// 4 consecutive values -> two 8-byte stores
__device__ __forceinline__ void store_pair4(__half* hi, __half* lo, size_t i, const float* v) {
  __align__(8) __half2 h[2];
  __align__(8) __half2 l[2];
  split_h2(v[0], v[1], h[0], l[0]);
  split_h2(v[2], v[3], h[1], l[1]);
  *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
  if (lo) *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
}
// 8 consecutive values -> two 16-byte stores
__device__ __forceinline__ void store_pair8(__half* hi, __half* lo, size_t i, const float* v) {
  __align__(16) __half2 h[4];
  __align__(16) __half2 l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_h2(v[2 * j], v[2 * j + 1], h[j], l[j]);
  *reinterpret_cast<uint4*>(hi + i) = *reinterpret_cast<const uint4*>(h);
  if (lo) *reinterpret_cast<uint4*>(lo + i) = *reinterpret_cast<const uint4*>(l);
}

// 32-byte (one full sector) global accesses, sm_100+: a lane that owns a whole row moves complete sectors
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ldg256f(const float* p, float* r) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256f(float* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]),
               "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
// 16 consecutive values -> two 32-byte stores (i must be a multiple of 16 elements from a 32-byte aligned base)
__device__ __forceinline__ void store_pair16(__half* hi, __half* lo, size_t i, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half2 hh, ll;
    split_h2(v[2 * j], v[2 * j + 1], hh, ll);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  stg256(hi + i, h);
  if (lo) stg256(lo + i, l);
}

// streaming variant: the stores bypass L1 and are first in line for L2 eviction (outputs nobody re-reads soon)
__device__ __forceinline__ void stg256_stream(void* p, const uint32_t* r) {
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void store_pair16_stream(__half* hi, __half* lo, size_t i, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half2 hh, ll;
    split_h2_nc(v[2 * j], v[2 * j + 1], hh, ll);      // callers: LayerNorm outputs (bounded by |gamma| * 16 + |beta|)
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  stg256_stream(hi + i, h);
  if (lo) stg256_stream(lo + i, l);
}

// Branch-free erf: odd rational x*P(x^2)/Q(x^2) on [-4, 4] (|erf| rounds to 1 beyond), max abs error 4.5e-7
// against a double-precision erf over [-6, 6] (checked on the host).  17 instructions
// instead of libdevice erff's two divergent branches: GELU is what bounds the ConvTranspose / fc1 epilogues.
__device__ __forceinline__ float erf_rational(float x) {
  x = fminf(fmaxf(x, -4.f), 4.f);
  const float x2 = x * x;
  float p = fmaf(x2, -2.72614225801306e-10f, 2.77068142495902e-08f);
  p = fmaf(p, x2, -2.10102402082508e-06f);
  p = fmaf(p, x2, -5.69250639462346e-05f);
  p = fmaf(p, x2, -7.34990630326855e-04f);
  p = fmaf(p, x2, -2.95459980854025e-03f);
  p = fmaf(p, x2, -1.60960333262415e-02f);
  p *= x;
  float q = fmaf(x2, -1.45660718464996e-05f, -2.13374055278905e-04f);
  q = fmaf(q, x2, -1.68282697438203e-03f);
  q = fmaf(q, x2, -7.37332916720468e-03f);
  q = fmaf(q, x2, -1.42647390514189e-02f);
  return __fdividef(p, q);
}
// Exact (erf) GELU in 12 instructions: gelu(x) = x Phi(x) = relu(x) - |x| Q(|x|) with Q(t) = Phi(-t) = erfc(t / sqrt2) / 2,
// and Q(t) = 2^p(t) with p a degree-8 polynomial fitted to log2 Q on [0, 5.6] (Chebyshev nodes; beyond 5.6 the
// term |x| Q is < 6e-8).  One FMNMX + 8 FFMA + MUFU.EX2 + FMNMX + FFMA against ~21 for the rational erf above:
// the ConvTranspose epilogues execute 3.2 G GELUs per 1024 prompts and are bound by exactly this count.
// Max abs error against a double-precision erf GELU over [-10, 10] (4 M points, fp32 evaluation): 4.8e-7
// (the rational-erf form: 1.6e-6); relative error of gelu(x) as x -> 0: 2.7e-6.
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(fabsf(x), 5.6f);
  float p = fmaf(-7.6606284746e-08f, t, -2.0927212226e-07f);
  p = fmaf(p, t, 4.7745383604e-05f);
  p = fmaf(p, t, -8.6993596824e-04f);
  p = fmaf(p, t, 8.3637992435e-03f);
  p = fmaf(p, t, -5.3778513192e-02f);
  p = fmaf(p, t, -4.5857301888e-01f);
  p = fmaf(p, t, -1.1512274409e+00f);
  p = fmaf(p, t, -9.9999607805e-01f);
  float q;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(p));
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CSAM_ACT_GELU) return gelu_erf(v);
  if (act == CSAM_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- PTX wrappers: mbarrier / TMA / tcgen05 -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Watchdog budget in SM clocks.  A lost arrival is a protocol bug of this library, never a data-dependent event;
// the budget only has to be far above any legitimate wait.  ~60 s at 2 GHz leaves room for compute-sanitizer /
// debugger single-stepping and time-sliced (MPS) sharing, where a 2 s budget could kill a healthy process
// (round-1 ADVICE); build with -DCSAM_MBAR_BUDGET=... to change it, -DCSAM_MBAR_DEBUG to print the waiter.
#ifndef CSAM_MBAR_BUDGET
#define CSAM_MBAR_BUDGET 120000000000LL
#endif
// Spin on the phase parity; a watchdog turns a lost arrival into a trap instead of a hang.  The trap carries no
// printf: a (never taken) vprintf call inside the wait loop made ptxas spill every live register around each wait
// -- 25 STL + 25 LDL per tile in the fused decoder epilogues (ncu: local memory traffic in the hot path).
// Build with -DCSAM_MBAR_DEBUG to get the block / thread of a lost arrival printed.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > CSAM_MBAR_BUDGET) {
#ifdef CSAM_MBAR_DEBUG
      printf("csam: mbarrier watchdog block %d thread %d\n", blockIdx.x, threadIdx.x);
#endif
      __trap();
    }
  }
}

// non-blocking probe of a phase (for a thread that has other useful work to issue meanwhile)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority hints: a stream that is read again soon (evict_last) next to one that is written once
// and never read by this kernel (evict_first stores below)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
// pull one box of a tensor map into L2 (no shared memory, no barrier): decouples HBM latency from the smem ring
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// One lane of a CONVERGED warp, chosen by the hardware.  Role branches must use this and not `lane == 0`: behind
// `elect.sync` ptxas knows that exactly one thread is active and issues the uniform-datapath instructions
// (UTCHMMA, UTMALDG, UTCBAR) back to back with their descriptors precomputed in uniform registers; behind a
// divergent `lane == 0` it wraps EVERY such instruction in an ELECT / BRA.U.ANY loop of ~8 dependent instructions
// (cuobjdump: 329 ELECT for 156 UTCHMMA in gemm.o), and the issuing thread needed ~140 clocks per MMA where a
// 128 x 64 x 16 MMA executes in 32 (timeline of K-T2I, scripts/trace_dec.py).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all tcgen05 ops previously issued by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive 32-bit columns, registers -> TMEM (thread i of the warp writes lane base+i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major or MN-major operand stored as rows of 128 bytes with
// the 128B swizzle (what TMA SWIZZLE_128B writes): 8-row atoms of 1024 B, SBO = 1024.
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same descriptor as two 32-bit words.  The high word is a constant of the layout (SBO = 1024, version 1,
// SWIZZLE_128B); the low word is (addr >> 4) | (LBO >> 4) << 16, so stepping an operand by `b` bytes inside a tile
// is ONE integer add of (b >> 4) (shared-memory addresses are < 256 KB: the 14-bit field cannot overflow).
// The single MMA-issuing thread is a real bottleneck for narrow MMAs (N = 64 executes in 32 cycles): rebuilding
// the 64-bit descriptor with shifts / masks for every instruction cost ~50 cycles per MMA.
constexpr uint32_t UMMA_DESC_HI_SW128 = 0x40004040u;
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_f16_w(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
// Same, with the A operand in TENSOR MEMORY (lane = row, two consecutive K elements per 32-bit column, i.e. 8 columns
// per K = 16 step): the MMA then reads only B from shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
// Instruction descriptor for kind::f16, fp16 A/B, fp32 accumulate (InstrDescriptor bit layout).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // c_format = F32
         | (0u << 7) | (0u << 10)           // a/b format = F16
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- host: TMA descriptor creation through the driver entry point (no -lcuda at build time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// 2-D fp16 tensor [rows, cols] with row stride ld (elements); box = [box_rows, box_cols]; 128B swizzle.
int make_tmap_2d_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, uint32_t box_cols);

}  // namespace csam
