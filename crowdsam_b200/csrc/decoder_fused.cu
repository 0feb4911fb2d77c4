// decoder_fused.cu — K-I2T: the image->token half of a two-way transformer layer as ONE kernel.
//
// Reference (transformer.py:184-190, Attention.forward :228-254), per prompt p and image token x (4096 per prompt):
//     q = q_proj(x + pe);  a = softmax(q k_t^T / 4) v_t  over the 7 prompt tokens, 8 heads x 16;
//     x' = LayerNorm(x + out_proj(a))
// The 7 tokens are tiny, so both projections are folded into per-prompt operands (csam_dec_fold_i2t):
//     scores[x, (h,j)] = x . B1[(h,j), 0:256] + peq[x] . B1[(h,j), 256:384]      B1 = log2e/4 * (Wq_h^T k_t[j,h] | blockdiag k_t)
//     out_proj(a)[x]   = P[x, (h,j)] . B2[:, (h,j)] + b_o                        B2 = Wo_h v_t[j,h]
// with peq = pe Wq^T + b_q a constant of the weights.  One CTA then does, per 128-row tile:
//     MMA1  S[128,64]  = [X | PEQ] (K = 384) * B1^T          (tcgen05, 3 MMAs per k-step: hi/lo split operands)
//     E1    P = softmax over the 7 tokens of each head       (fp32, four threads per row; P goes to TENSOR MEMORY as an
//                                                             h16 pair and enters MMA2 as its A operand -- no shared-
//                                                             memory P buffer, which pays for a third ring stage)
//     MMA2  O[128,256] += P (K = 64) * B2^T
//     MMAr  O[128,256] = X * I                               (the residual enters the accumulator THROUGH the tensor core:
//                                                             while k-block kb of X is resident for MMA1, hi * I + lo * I
//                                                             with a 64 x 64 identity writes x (exactly: hi + lo fits fp32)
//                                                             into columns 64 kb .. 64 kb + 63 of O; MMA2 then accumulates)
//     E2    x' = LayerNorm(O + b_o) -> h16 pair
// so the [P*4096,128] q and attention-output streams and the separate out_proj GEMM of the unfused path never
// exist: the layer reads X once (1 KB per row) and writes X' once (1 KB per row).
//   warp 0       TMA producer  (A / B1 k-blocks through a 3-stage ring; B2 once per prompt)
//   warp 1       MMA issuer    (MMA2 of tile i, then MMA1 of tile i+1 while the epilogue normalises tile i)
//   warp 2       TMEM allocator (S0 S1 O P = 64 + 64 + 256 + 64 columns)
//   warps 4..19  epilogue: residual fetch (i) | drain O(i) -> normalise + store (i) -> E1(i+1);
//                four threads per row (64 columns each): 16 warps hide the load / TMEM / barrier latencies that
//                8 warps could not (ncu: issue slots 22 % active, long-scoreboard + barrier stalls dominant)
#include "common.cuh"

namespace csam {

// Timeline instrumentation (builds with -DCSAM_TRACE only, scripts/trace_dec.py): CTA 0 stamps clock64 at the
// hand-over points of the pipelines below; the product build compiles the macro away.
#ifdef CSAM_TRACE
// one slot per (tag, tile): a plain store, so the stamping thread never waits (an atomic slot counter made the
// MMA-issuing thread wait for an L2 round trip per stamp and showed up as ~1500 clocks per k-block)
constexpr int TRACE_TAGS = 160, TRACE_TILES = 64;
__device__ unsigned long long g_trace[TRACE_TAGS * TRACE_TILES];
#define CSAM_TR(tag, val)                                                                      \
  do {                                                                                         \
    if (blockIdx.x == 0 && (unsigned)(val) < (unsigned)TRACE_TILES)                            \
      g_trace[(tag) * TRACE_TILES + (val)] = (unsigned long long)clock64();                    \
  } while (0)
#else
#define CSAM_TR(tag, val) do { } while (0)
#endif

constexpr int I2T_BM = 128;
constexpr int I2T_THREADS = 640;                 // 4 control warps + 16 epilogue warps (4 per TMEM lane quadrant)
constexpr int I2T_KB1 = 6;                       // 4 k-blocks of X (256) + 2 of PEQ (128)
#ifndef I2T_STAGES_N
#define I2T_STAGES_N 3
#endif
constexpr int I2T_STAGES = I2T_STAGES_N;
constexpr int I2T_A_BYTES = 128 * 64 * 2;        // 16 KB per operand half
constexpr int I2T_B1_BYTES = 64 * 64 * 2;        // 8 KB
constexpr int I2T_STAGE_BYTES = 2 * I2T_A_BYTES + 2 * I2T_B1_BYTES;   // 48 KB
constexpr int I2T_B2_BYTES = 256 * 64 * 2;       // 32 KB per half
constexpr int I2T_OFF_B2 = I2T_STAGES * I2T_STAGE_BYTES;
constexpr int I2T_OFF_ID = I2T_OFF_B2 + 2 * I2T_B2_BYTES;    // 64 x 64 fp16 identity, K-major, 128B swizzle (8 KB)
constexpr int I2T_TM_O = 128, I2T_TM_PH = 384, I2T_TM_PL = 416;   // TMEM columns: S (hi | lo part) | O | P hi | P lo
constexpr int I2T_OFF_BAR = I2T_OFF_ID + 64 * 64 * 2;
constexpr int I2T_OFF_EPI = I2T_OFF_BAR + 256;
constexpr int I2T_SMEM_BYTES = I2T_OFF_EPI + (8 * 128 + 3 * 256) * 4;
static_assert(I2T_SMEM_BYTES <= 227 * 1024, "i2t layer shared memory budget");

struct I2TBars {
  uint64_t full[I2T_STAGES], empty[I2T_STAGES];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full, o_full, o_empty, b2_full, b2_empty;
  uint32_t tmem_slot;
};

struct I2TParams {
  int x_shared;              // 1: the same 4096 key rows for every prompt (layer 0)
  int tiles;                 // P * 32
  int pf;                    // L2 prefetch of the X stream: 0 none (default, fastest), 1 rest of this tile, 2 next tile
  const float* bias; const float* gamma; const float* beta; float eps;
  __half* out_hi; __half* out_lo;
};

__device__ __forceinline__ float ex2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// k-block order of MMA1: the two positional blocks first -- they carry no residual, so they may run while the
// epilogue still drains the previous tile's accumulator -- then the four X blocks
__device__ __forceinline__ int i2t_kb(int i) { return i < 2 ? 4 + i : i - 2; }

__global__ void __launch_bounds__(I2T_THREADS, 1)
dec_i2t_layer_kernel(const __grid_constant__ CUtensorMap tx_hi, const __grid_constant__ CUtensorMap tx_lo,
                     const __grid_constant__ CUtensorMap tq_hi, const __grid_constant__ CUtensorMap tq_lo,
                     const __grid_constant__ CUtensorMap tb1_hi, const __grid_constant__ CUtensorMap tb1_lo,
                     const __grid_constant__ CUtensorMap tb2_hi, const __grid_constant__ CUtensorMap tb2_lo,
                     I2TParams a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  I2TBars* bars = reinterpret_cast<I2TBars*>(smem + I2T_OFF_BAR);
  float* epi = reinterpret_cast<float*>(smem + I2T_OFF_EPI);   // ex_sum[4][128] ex_sq[4][128] gamma beta bias
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous tile range per CTA: consecutive tiles share a prompt, so B2 is loaded about once per 32 tiles
  const int t0 = (int)((long long)a.tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)a.tiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tx_hi); tma_prefetch_desc(&tx_lo); tma_prefetch_desc(&tq_hi); tma_prefetch_desc(&tq_lo);
    tma_prefetch_desc(&tb1_hi); tma_prefetch_desc(&tb1_lo); tma_prefetch_desc(&tb2_hi); tma_prefetch_desc(&tb2_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < I2T_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->s_full[b], 1); mbar_init(&bars->s_empty[b], 16); }
    mbar_init(&bars->p_full, 16);
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, 16);
    mbar_init(&bars->b2_full, 1);
    mbar_init(&bars->b2_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(&bars->tmem_slot);
  {
    // identity operand of the residual MMAs: row n holds K elements 8c .. 8c+7 in 16-byte chunk c ^ (n & 7)
    uint4* idm = reinterpret_cast<uint4*>(smem + I2T_OFF_ID);
    for (int i = threadIdx.x; i < 512; i += I2T_THREADS) {
      const int n = i >> 3, c = (i & 7) ^ (n & 7);             // physical chunk i & 7 of row n holds logical chunk c
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (c == (n >> 3)) {
        const uint32_t one = 0x3C00u << (16 * (n & 1));         // fp16 1.0 at element n & 7 of the chunk
        const int w = (n & 7) >> 1;
        v.x = w == 0 ? one : 0u; v.y = w == 1 ? one : 0u; v.z = w == 2 ? one : 0u; v.w = w == 3 ? one : 0u;
      }
      idm[i] = v;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;      // S hi part [0,64) lo part [64,128) | O [128,384) | P hi [384,416) lo [416,448)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0; int b2_cnt = 0;
      for (int t = t0; t < t1; ++t) {
        const int p = t >> 5, mrow = (t & 31) * I2T_BM;
        const int arow = a.x_shared ? mrow : t * I2T_BM;
        if (!a.x_shared && a.pf > 0) {
          // the k-blocks of this tile beyond the ring depth are requested from HBM right away, so that the ring's
          // loads find them in L2.  Looking further ahead (whole tiles) was measured SLOWER (4.1 against 4.9 TB/s)
          for (int kb = (a.pf > 1 ? 0 : I2T_STAGES); kb < 4; ++kb) {
            const int tp = a.pf > 1 ? t + 1 : t;
            if (tp < t1) {
              tma_prefetch_l2_2d(&tx_hi, kb * 64, tp * I2T_BM);
              tma_prefetch_l2_2d(&tx_lo, kb * 64, tp * I2T_BM);
            }
          }
        }
        for (int i = 0; i < I2T_KB1; ++i) {
          const int kb = i2t_kb(i);
          mbar_wait(&bars->empty[stage], phase ^ 1);
          CSAM_TR(100 + kb, t - t0);
          uint8_t* sa = smem + stage * I2T_STAGE_BYTES;
          uint8_t* sb = sa + 2 * I2T_A_BYTES;
          mbar_expect_tx(&bars->full[stage], I2T_STAGE_BYTES);
          if (kb < 4) {
            tma_load_2d(sa, &tx_hi, &bars->full[stage], kb * 64, arow);
            tma_load_2d(sa + I2T_A_BYTES, &tx_lo, &bars->full[stage], kb * 64, arow);
          } else {
            tma_load_2d(sa, &tq_hi, &bars->full[stage], (kb - 4) * 64, mrow);
            tma_load_2d(sa + I2T_A_BYTES, &tq_lo, &bars->full[stage], (kb - 4) * 64, mrow);
          }
          tma_load_2d(sb, &tb1_hi, &bars->full[stage], kb * 64, p * 64);
          tma_load_2d(sb + I2T_B1_BYTES, &tb1_lo, &bars->full[stage], kb * 64, p * 64);
          if (++stage == I2T_STAGES) { stage = 0; phase ^= 1; }
        }
        if (t == t0 || (t & 31) == 0) {
          // first tile of a prompt in this CTA: its B2 (needed by MMA2 only) replaces the previous prompt's
          mbar_wait(&bars->b2_empty, (b2_cnt & 1) ^ 1);
          mbar_expect_tx(&bars->b2_full, 2 * I2T_B2_BYTES);
          tma_load_2d(smem + I2T_OFF_B2, &tb2_hi, &bars->b2_full, 0, p * 256);
          tma_load_2d(smem + I2T_OFF_B2 + I2T_B2_BYTES, &tb2_lo, &bars->b2_full, 0, p * 256);
          ++b2_cnt;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_f16(I2T_BM, 64, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(I2T_BM, 256, 0, 0);
      constexpr uint32_t idesc_r = umma_idesc_f16(I2T_BM, 16, 0, 0);
      constexpr uint32_t idesc1w = umma_idesc_f16(I2T_BM, 128, 0, 0);
      const uint32_t idd = umma_desc_lo(smem_u32(smem + I2T_OFF_ID), 16);
      int stage = 0; uint32_t phase = 0; int b2_cnt = 0;
      auto mma1 = [&](int li) {
        // ONE 128-column score buffer (hi part | lo part, see below): MMA1(li) is issued after MMA2(li-1), i.e. after
        // the softmax of tile li-1 has long read it, so the second buffer of the earlier design bought nothing
        mbar_wait(&bars->s_empty[0], (li & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base;
        for (int i = 0; i < I2T_KB1; ++i) {
          const int kb = i2t_kb(i);
          mbar_wait(&bars->full[stage], phase);
          CSAM_TR(110 + kb, li);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * I2T_STAGE_BYTES);
          const uint32_t ad = umma_desc_lo(sa, 16);
          const uint32_t bd = umma_desc_lo(sa + 2 * I2T_A_BYTES, 16);
          // two MMAs per k-step: X_hi x [B1_hi | B1_lo] as ONE N = 128 operand (the two halves of a stage's B1 block are
          // contiguous) into S columns [0,64) | [64,128), then X_lo x B1_hi into [0,64); E1 adds the halves
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_w(d, ad + 2 * k, bd + 2 * k, idesc1w, (i | k) ? 1u : 0u);
            umma_f16_w(d, ad + (I2T_A_BYTES >> 4) + 2 * k, bd + 2 * k, idesc1, 1u);
          }
          if (kb == 0) {
            // the residual MMAs overwrite O: tile li-1 must be drained.  Everything issued above (the positional part
            // of the scores and the first X block) did not need O and ran under MMA2(li-1) / the drain.
            mbar_wait(&bars->o_empty, (li & 1) ^ 1);
            CSAM_TR(121, li);
            tc_fence_after();
          }
          if (kb < 4) {
            // residual: O[:, 64 kb + 16 k ..+16) = X_hi * I + X_lo * I over the 16 features of k-step k (N = 16 MMAs
            // against rows 16 k ..+16 of the identity: 2048 k bytes further, plus the usual 32 k bytes along K)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t dr = tmem_base + 128 + kb * 64 + k * 16;
              umma_f16_w(dr, ad + 2 * k, idd + 130 * k, idesc_r, 0u);
              umma_f16_w(dr, ad + (I2T_A_BYTES >> 4) + 2 * k, idd + 130 * k, idesc_r, 1u);
            }
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == I2T_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->s_full[0]);
      };
      if (t0 < t1) mma1(0);
      int li = 0;
      for (int t = t0; t < t1; ++t, ++li) {
        if (t == t0 || (t & 31) == 0) { mbar_wait(&bars->b2_full, b2_cnt & 1); ++b2_cnt; }
        mbar_wait(&bars->p_full, li & 1);
        CSAM_TR(120, li);
        tc_fence_after();
        const uint32_t bd = umma_desc_lo(smem_u32(smem + I2T_OFF_B2), 16);
        const uint32_t d = tmem_base + I2T_TM_O;
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // A = P from tensor memory: 8 columns per K = 16 step
          umma_f16_ts(d, tmem_base + I2T_TM_PH + 8 * k, bd + 2 * k, idesc2, 1u);     // on top of the residual MMA1 left there
          umma_f16_ts(d, tmem_base + I2T_TM_PL + 8 * k, bd + 2 * k, idesc2, 1u);
          umma_f16_ts(d, tmem_base + I2T_TM_PH + 8 * k, bd + (I2T_B2_BYTES >> 4) + 2 * k, idesc2, 1u);
        }
        umma_commit(&bars->o_full);
        if (t + 1 == t1 || ((t + 1) & 31) == 0) umma_commit(&bars->b2_empty);   // last tile of this prompt here
        if (t + 1 < t1) mma1(li + 1);
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    // 640 threads start at 96 registers (61440 for the CTA -- setmaxnreg only redistributes WITHIN that launch
    // allocation); the 4 control warps hand 40 each back so that the 16 epilogue warps run at 104
    // (128 x 56 + 512 x 104 = 60416)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int ew = warp - 4;
    const int q = ew & 3;                        // TMEM lane quadrant == warp % 4
    const int cq = ew >> 2;                      // column quarter (S: 16 of 64 columns, O: 64 of 256)
    const int r = q * 32 + lane;                 // row of the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int et = threadIdx.x - 128;            // 0..511
    float* ex_sum = epi;                         // [4][128]
    float* ex_sq = epi + 512;                    // [4][128]
    float* s_gamma = epi + 1024;
    float* s_beta = s_gamma + 256;
    float* s_bias = s_gamma + 512;
    if (et < 256) {
      s_gamma[et] = a.gamma[et];
      s_beta[et] = a.beta[et];
      s_bias[et] = a.bias ? a.bias[et] : 0.f;
    }
    asm volatile("bar.sync 5, 512;" ::: "memory");
    const bool has_bias = a.bias != nullptr;

    // E1: softmax over the 7 tokens of each of this thread's 2 heads -> P (hi/lo) in tensor memory
    auto softmax_tile = [&](int li) {
      if (warp == 4 && lane == 0) CSAM_TR(130, li);
      mbar_wait(&bars->s_full[0], li & 1);
      if (warp == 4 && lane == 0) CSAM_TR(131, li);
      tc_fence_after();
      uint32_t raw[16];
      {
        uint32_t raw2[16];
        tmem_ld16(lane_addr + cq * 16, raw);
        tmem_ld16(lane_addr + 64 + cq * 16, raw2);         // the X_hi x B1_lo part of the scores
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) raw[c] = __float_as_uint(__uint_as_float(raw[c]) + __uint_as_float(raw2[c]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty[0]);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float s[7];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 7; ++j) { s[j] = __uint_as_float(raw[hh * 8 + j]); m = fmaxf(m, s[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) { s[j] = ex2f_approx(s[j] - m); sum += s[j]; }
        const float inv = 1.0f / sum;
        float pr[8];
#pragma unroll
        for (int j = 0; j < 7; ++j) pr[j] = s[j] * inv;
        pr[7] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          __half2 h2, l2;
          split_h2_nc(pr[j], pr[j + 1], h2, l2);
          hi[hh * 4 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[hh * 4 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);
        }
      }
      // this thread's 16 (head, token) columns are K elements 16 cq .. 16 cq + 15 of P = K step cq of MMA2's A operand
      tmem_st8(lane_addr + I2T_TM_PH + 8 * cq, hi);
      tmem_st8(lane_addr + I2T_TM_PL + 8 * cq, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    };

    if (t0 < t1) softmax_tile(0);
    int li = 0;
    for (int t = t0; t < t1; ++t, ++li) {
      float x[64];
      if (warp == 4 && lane == 0) CSAM_TR(132, li);
      mbar_wait(&bars->o_full, li & 1);
      if (warp == 4 && lane == 0) CSAM_TR(133, li);
      tc_fence_after();
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t raw[16];
        tmem_ld16(lane_addr + 128 + cq * 64 + c, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          x[c + j + 0] = __uint_as_float(raw[j + 0]);      // residual + out_proj(attention), both from the tensor core
          x[c + j + 1] = __uint_as_float(raw[j + 1]);
          x[c + j + 2] = __uint_as_float(raw[j + 2]);
          x[c + j + 3] = __uint_as_float(raw[j + 3]);
          if (has_bias) {        // only when the caller did not fold out_proj's bias into B2
            const float4 bb = *reinterpret_cast<const float4*>(s_bias + cq * 64 + c + j);
            x[c + j + 0] += bb.x; x[c + j + 1] += bb.y; x[c + j + 2] += bb.z; x[c + j + 3] += bb.w;
          }
          sum += (x[c + j] + x[c + j + 1]) + (x[c + j + 2] + x[c + j + 3]);
        }
      }
      // O drained (and MMA2(li) has finished reading P): release both before the long normalise + store phase
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_empty);
      if (warp == 4 && lane == 0) CSAM_TR(134, li);
      ex_sum[cq * 128 + r] = sum;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      const float mean = ((ex_sum[r] + ex_sum[128 + r]) + (ex_sum[256 + r] + ex_sum[384 + r])) * (1.0f / 256.0f);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) { const float d = x[j] - mean; sq = fmaf(d, d, sq); }
      ex_sq[cq * 128 + r] = sq;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      const float rstd = 1.0f / sqrtf(((ex_sq[r] + ex_sq[128 + r]) + (ex_sq[256 + r] + ex_sq[384 + r])) * (1.0f / 256.0f) + a.eps);
      const size_t orow = (size_t)t * I2T_BM + r;
      const float nmr = -mean * rstd;
      auto norm_store = [&](int c) {
        const int col = cq * 64 + c;
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 g = *reinterpret_cast<const float4*>(s_gamma + col + j);
          const float4 bt = *reinterpret_cast<const float4*>(s_beta + col + j);
          y[j + 0] = fmaf(fmaf(x[c + j + 0], rstd, nmr), g.x, bt.x);     // nmr = -mean * rstd
          y[j + 1] = fmaf(fmaf(x[c + j + 1], rstd, nmr), g.y, bt.y);
          y[j + 2] = fmaf(fmaf(x[c + j + 2], rstd, nmr), g.z, bt.z);
          y[j + 3] = fmaf(fmaf(x[c + j + 3], rstd, nmr), g.w, bt.w);
        }
        store_pair16_stream(a.out_hi, a.out_lo, orow * 256 + col, y);
      };
      norm_store(0);
      norm_store(16);
      norm_store(32);
      // E1 of the next tile before the last quarter of the stores: MMA1(li+1) had the first three quarters to finish,
      // and MMA2(li+1) -- which this warp would otherwise wait for idle -- runs under the last quarter
      if (warp == 4 && lane == 0) CSAM_TR(135, li);
      if (t + 1 < t1) softmax_tile(li + 1);
      norm_store(48);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---- per-prompt folded operands -------------------------------------------------------------------------
// B1 [P*64, 384], row n = h*8 + j (j = 7 is a zero row): cols 0..255 = s * sum_d Wq[h*16+d][c] k_t[j][h*16+d],
// cols 256 + h'*16 + d = (h' == h) ? s * k_t[j][h*16+d] : 0, s = log2(e) / sqrt(16).
// B2 [P*256, 64], row c: col n = h*8 + j = sum_d Wo[c][h*16+d] v_t[j][h*16+d] (zero for j = 7).
__global__ void __launch_bounds__(256)
dec_fold_i2t_kernel(const float* __restrict__ kt, const float* __restrict__ vt, const float* __restrict__ wq,
                    const float* __restrict__ wo, const float* __restrict__ bo, __half* b1_hi, __half* b1_lo,
                    __half* b2_hi, __half* b2_lo) {
  __shared__ float ks[7][128];
  __shared__ float vs[7][128];
  const int p = blockIdx.x, c = threadIdx.x;
  constexpr float SC = 0.25f * 1.4426950408889634f;
  for (int i = c; i < 7 * 128; i += 256) {
    ks[i >> 7][i & 127] = kt[(size_t)p * 896 + i] * SC;
    vs[i >> 7][i & 127] = vt ? vt[(size_t)p * 896 + i] : 0.f;
  }
  __syncthreads();
  __half* r1h = b1_hi + (size_t)p * 64 * 384;
  __half* r1l = b1_lo + (size_t)p * 64 * 384;
  for (int h = 0; h < 8; ++h) {
    float w[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) w[d] = wq[(size_t)(h * 16 + d) * 256 + c];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
      if (j < 7) {
#pragma unroll
        for (int d = 0; d < 16; ++d) acc = fmaf(w[d], ks[j][h * 16 + d], acc);
      }
      store_pair(r1h, r1l, (size_t)(h * 8 + j) * 384 + c, acc);
      if (c < 128) {
        const float v = (j < 7 && (c >> 4) == h) ? ks[j][c] : 0.f;
        store_pair(r1h, r1l, (size_t)(h * 8 + j) * 384 + 256 + c, v);
      }
    }
  }
  if (!vt) return;                     // token -> image folding needs B1 only
  // out_proj's bias rides along: the probabilities of every head sum to 1, so adding b_o[c] / 8 to each of the 7
  // columns of each of the 8 heads adds exactly b_o[c] to the product (one add per element less in the epilogue)
  const float bshare = bo ? bo[c] * 0.125f : 0.f;
  float o[64];
  const float* wrow = wo + (size_t)c * 128;
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    float w[16];
#pragma unroll
    for (int d = 0; d < 16; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(wrow + h * 16 + d);
      w[d] = t.x; w[d + 1] = t.y; w[d + 2] = t.z; w[d + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
      if (j < 7) {
        acc = bshare;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc = fmaf(w[d], vs[j][h * 16 + d], acc);
      }
      o[h * 8 + j] = acc;
    }
  }
  const size_t ob = ((size_t)p * 256 + c) * 64;
#pragma unroll
  for (int i = 0; i < 64; i += 8) store_pair8(b2_hi, b2_lo, ob + i, o + i);
}


// =====================================================================================================
// K-T2I: token -> image cross attention with the k / v projections folded away (transformer.py:171-176,
// 104-112; Attention.forward :228-254).  Reference, per prompt: k = k_proj(x + pe), v = v_proj(x) for the 4096
// image tokens x; the 7 prompt tokens attend over them with 8 heads x 16.  Here
//     scores[x, (h,i)] = x . B1[(h,i), 0:256] + pek[x] . B1[(h,i), 256:384]     (B1 from csam_dec_fold_t2i: q tokens, Wk)
//     xbar[(h,i), :]   = sum_x softmax_x(scores)[x, (h,i)] * x                  (a 256-vector per head and token)
//     out[i, h*16+d]   = Wv[h*16+d, :] . xbar[(h,i), :] + b_v                   (csam_dec_t2i_out, tiny)
// so the kernel reads the keys once (1 KB per row) and the [P*4096, 256] k | v stream of the unfused path
// (4 KB per row written + read) never exists.  One CTA owns a prompt; per 128-key tile:
//     MMA1  S[128 keys, 64] = [PEK | X] (K = 384, the two positional k-blocks first) * B1^T
//     softmax over KEYS: two threads per key row (32 columns each); the column maxima live in shared memory and may lag by
//           2^8, so the common path is exp2(s - m[c]) and a block-wide vote; a violated bound triggers
//           an exact column maximum (warp shuffles) and a rescale of the accumulator in TMEM
//     MMA2  XBAR^T[256, 64] += X^T (MN-major A straight from the resident X tile) * P (MN-major B, h16 pair)
// The X tile (128 KB as hi + lo) stays in shared memory from MMA1 to MMA2, so tiles are processed one
// at a time; the next tile is prefetched into L2 meanwhile.
constexpr int T2I_THREADS = 640;                 // 4 control warps + 16 softmax warps (four threads per key row)
constexpr int T2I_SLOT = 32768;                  // one X / PEK k-block: hi 16 KB | lo 16 KB
constexpr int T2I_OFF_PEK = 4 * T2I_SLOT;
constexpr int T2I_OFF_B1 = T2I_OFF_PEK + T2I_SLOT;             // 2-stage ring of B1 k-blocks (hi 8 KB | lo 8 KB)
constexpr int T2I_OFF_P = T2I_OFF_B1 + 2 * 2 * I2T_B1_BYTES;   // P [128 keys][64] fp16, hi then lo
constexpr int T2I_OFF_BAR = T2I_OFF_P + 2 * 16384;
constexpr int T2I_OFF_ST = T2I_OFF_BAR + 256;                  // m[64] l[64] alpha[64] wmax[4][64]
constexpr int T2I_SMEM_BYTES = T2I_OFF_ST + (3 * 64 + 4 * 64) * 4;
static_assert(T2I_SMEM_BYTES <= 227 * 1024, "t2i shared memory budget");
constexpr float T2I_TAU = 8.f;

struct T2IBars {
  uint64_t x_full[4], x_empty[2], pek_full, pek_empty, b_full[2], b_empty[2], v_full[2];
  uint64_t s_full, s_empty, p_full, pv_done;
  uint32_t tmem_slot;
};

struct T2IParams {
  int x_shared, P, pf;
  float* xbar;               // [P, 64, 256]
};

// k-block order of MMA1: the two PEK blocks first (they do not wait for the X slots, so they run under the previous
// tile's softmax), then the four X blocks
__device__ __forceinline__ int t2i_kb(int i) { return i < 2 ? 4 + i : i - 2; }

__global__ void __launch_bounds__(T2I_THREADS, 1)
dec_t2i_kernel(const __grid_constant__ CUtensorMap tx_hi, const __grid_constant__ CUtensorMap tx_lo,
               const __grid_constant__ CUtensorMap tk_hi, const __grid_constant__ CUtensorMap tk_lo,
               const __grid_constant__ CUtensorMap tb1_hi, const __grid_constant__ CUtensorMap tb1_lo, T2IParams a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  T2IBars* bars = reinterpret_cast<T2IBars*>(smem + T2I_OFF_BAR);
  float* st_m = reinterpret_cast<float*>(smem + T2I_OFF_ST);
  float* st_l = st_m + 64;
  float* st_alpha = st_m + 128;
  float* st_wmax = st_m + 192;                     // [4][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tx_hi); tma_prefetch_desc(&tx_lo); tma_prefetch_desc(&tk_hi); tma_prefetch_desc(&tk_lo);
    tma_prefetch_desc(&tb1_hi); tma_prefetch_desc(&tb1_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars->x_full[i], 1);
    mbar_init(&bars->x_empty[0], 1); mbar_init(&bars->x_empty[1], 1);
    mbar_init(&bars->pek_full, 1); mbar_init(&bars->pek_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); mbar_init(&bars->v_full[i], 1); }
    mbar_init(&bars->s_full, 1); mbar_init(&bars->s_empty, 16);
    mbar_init(&bars->p_full, 16); mbar_init(&bars->pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(&bars->tmem_slot);
  if (threadIdx.x < 64) { st_m[threadIdx.x] = -INFINITY; st_l[threadIdx.x] = 0.f; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // S: hi part [0,64) + lo part [64,128) | XBAR^T features 0..127: A [128,192) B [192,256) | features 128..255: A [256,320) B [320,384)
  // (A collects X_hi^T P_hi + X_lo^T P_hi, B collects X_hi^T P_lo; their sum is formed once per prompt)
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer #1: the key tiles
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");       // 128 x 56 + 512 x 104 <= 640 x 96 (as in K-I2T)
    // Two producers, so that the X loads of a tile are requested the moment MMA2 of the previous tile lets go of
    // the slots.  (One thread used to walk X / PEK / B1 in k-block order: the X load of k-block kb + 2 then sat
    // behind a wait on the 2-deep B1 ring, i.e. behind MMA1 of k-block kb, and the four load latencies of a tile
    // were paid one after the other -- the softmax warps spent 69 % of their time waiting for S.)
    if (elect_one()) {
      int tl = 0;
      for (int p = blockIdx.x; p < a.P; p += gridDim.x) {
        for (int ti = 0; ti < 32; ++ti, ++tl) {
          const int arow = (a.x_shared ? 0 : p * 4096) + ti * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // slots 0,1 / 2,3 are released separately, as soon as MMA2 of the previous tile is done with them
            mbar_wait(&bars->x_empty[half], (tl & 1) ^ 1);
            CSAM_TR(10 + half, tl);
#pragma unroll
            for (int kb = 2 * half; kb < 2 * half + 2; ++kb) {
              uint8_t* sx = smem + kb * T2I_SLOT;
              mbar_expect_tx(&bars->x_full[kb], T2I_SLOT);
              tma_load_2d(sx, &tx_hi, &bars->x_full[kb], kb * 64, arow);
              tma_load_2d(sx + 16384, &tx_lo, &bars->x_full[kb], kb * 64, arow);
            }
            if (half == (a.pf == 2 ? 1 : 0)) {
              // (experiments, CSAM_T2I_PF) 1: the NEXT tile of X goes to L2 now, so that its loads (which must wait for
              // this tile's MMA2) are L2 hits; 2: only its first two k-blocks, the ones MMA1 waits for, requested after
              // this tile's loads.  Default 0: no look-ahead measured fastest (1134 us; 1: 1184 us)
              int nrow = -1;
              if (ti + 1 < 32) nrow = arow + 128;
              else if (!a.x_shared && p + (int)gridDim.x < a.P) nrow = (p + (int)gridDim.x) * 4096;
              else if (a.x_shared && p + (int)gridDim.x < a.P) nrow = 0;
              if (nrow >= 0 && a.pf) {
                for (int k2 = 0; k2 < (a.pf == 2 ? 2 : 4); ++k2) {
                  tma_prefetch_l2_2d(&tx_hi, k2 * 64, nrow);
                  tma_prefetch_l2_2d(&tx_lo, k2 * 64, nrow);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer #2: PEK tiles and B1
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // B1 (96 KB per prompt) does not fit next to the key tile, so its six k-blocks are streamed per tile.  A 2-slot
    // ring for all six exposed one L2 latency per k-block (a slot is free only after its MMAs have COMPLETED): the
    // timeline showed 1750 clocks per k-block, 10.5 k per tile.  Now the two dedicated slots serve the blocks whose
    // loads hide under other work (PE blocks 4,5 under the previous softmax, X blocks 0,1 under softmax / MMA2), and
    // X blocks 2,3 land in the P buffer, which is dead from MMA2(t-1) to softmax(t) -- exactly when the key slots
    // 2,3 are refilled, so both arrive together.
    if (elect_one()) {
      int bcnt = 0, pcnt = 0, tl = 0;
      for (int p = blockIdx.x; p < a.P; p += gridDim.x) {
        for (int ti = 0; ti < 32; ++ti, ++tl) {
          const int mrow = ti * 128;
          for (int i = 0; i < 4; ++i) {
            const int kb = t2i_kb(i);
            if (kb >= 4) {
              mbar_wait(&bars->pek_empty, (pcnt & 1) ^ 1);
              uint8_t* sx = smem + T2I_OFF_PEK;
              mbar_expect_tx(&bars->pek_full, T2I_SLOT);
              tma_load_2d(sx, &tk_hi, &bars->pek_full, (kb - 4) * 64, mrow);
              tma_load_2d(sx + 16384, &tk_lo, &bars->pek_full, (kb - 4) * 64, mrow);
              ++pcnt;
            }
            const int bs = bcnt & 1;
            mbar_wait(&bars->b_empty[bs], ((bcnt >> 1) & 1) ^ 1);
            uint8_t* sb = smem + T2I_OFF_B1 + bs * 2 * I2T_B1_BYTES;
            mbar_expect_tx(&bars->b_full[bs], 2 * I2T_B1_BYTES);
            tma_load_2d(sb, &tb1_hi, &bars->b_full[bs], kb * 64, p * 64);
            tma_load_2d(sb + I2T_B1_BYTES, &tb1_lo, &bars->b_full[bs], kb * 64, p * 64);
            ++bcnt;
          }
          if (tl > 0) mbar_wait(&bars->pv_done, (tl - 1) & 1);       // MMA2 of the previous tile has read P
          for (int v = 0; v < 2; ++v) {
            uint8_t* sb = smem + T2I_OFF_P + v * 2 * I2T_B1_BYTES;
            mbar_expect_tx(&bars->v_full[v], 2 * I2T_B1_BYTES);
            tma_load_2d(sb, &tb1_hi, &bars->v_full[v], (2 + v) * 64, p * 64);
            tma_load_2d(sb + I2T_B1_BYTES, &tb1_lo, &bars->v_full[v], (2 + v) * 64, p * 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_f16(128, 64, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(128, 64, 1, 1);     // A = X^T and B = P, both MN-major
      constexpr uint32_t idesc2w = umma_idesc_f16(128, 128, 1, 1);
      constexpr uint32_t idesc1w = umma_idesc_f16(128, 128, 0, 0);
      int bcnt = 0, pcnt = 0;
      int n_tiles_total = 0;
      for (int p = blockIdx.x; p < a.P; p += gridDim.x) n_tiles_total += 32;
      // MMA1 of tile tl, k-blocks [i0, i1) in the order of t2i_kb
      auto mma1 = [&](int tl, int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
          const int kb = t2i_kb(i);
          uint32_t sa;
          if (kb < 4) {
            mbar_wait(&bars->x_full[kb], tl & 1);
            CSAM_TR(20 + kb, tl);
            sa = smem_u32(smem + kb * T2I_SLOT);
          } else {
            mbar_wait(&bars->pek_full, pcnt & 1);
            sa = smem_u32(smem + T2I_OFF_PEK);
          }
          const int bs = bcnt & 1;
          uint32_t sbb;
          if (i < 4) {
            mbar_wait(&bars->b_full[bs], (bcnt >> 1) & 1);
            sbb = smem_u32(smem + T2I_OFF_B1 + bs * 2 * I2T_B1_BYTES);
          } else {
            mbar_wait(&bars->v_full[i - 4], tl & 1);
            sbb = smem_u32(smem + T2I_OFF_P + (i - 4) * 2 * I2T_B1_BYTES);
          }
          tc_fence_after();
          const uint32_t ad = umma_desc_lo(sa, 16);
          const uint32_t bd = umma_desc_lo(sbb, 16);
          // two MMAs per k-step: X_hi x [B1_hi | B1_lo] as ONE N = 128 operand (a B1 block is 64 hi rows followed by
          // 64 lo rows) into S columns [0,64) | [64,128), then X_lo x B1_hi into [0,64); the softmax adds the two
          // halves.  X_hi is fetched once instead of twice: 14 KB instead of 18 KB of shared-memory operand reads per
          // step for a kernel that is bound by exactly that (449 against 576 clocks per k-block in isolation)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_w(tmem_base, ad + 2 * k, bd + 2 * k, idesc1w, (i | k) ? 1u : 0u);
            umma_f16_w(tmem_base, ad + (16384 >> 4) + 2 * k, bd + 2 * k, idesc1, 1u);
          }
          if (i < 4) { umma_commit(&bars->b_empty[bs]); ++bcnt; }      // blocks 2,3 sit in the P buffer: freed by pv_done
          if (kb >= 4) { umma_commit(&bars->pek_empty); ++pcnt; }
        }
        if (i1 == 6) umma_commit(&bars->s_full);
      };
      if (n_tiles_total > 0) mma1(0, 0, 6);
      for (int tl = 0; tl < n_tiles_total; ++tl) {
        const int ti = tl & 31;
        // the PEK part of the next tile's scores runs under this tile's softmax (S is free once it has been read).
        // Its second block goes through the same PEK slot as the first, i.e. one L2 round trip later: it is issued
        // when it happens to be ready before P is, and otherwise after MMA2 -- a blocking wait here held MMA2 back.
        // (Always deferring it was measured slower, 1280 against 1240 us: its refill traffic then lands in the
        // middle of the X part of MMA1, whose SS-mode MMAs are bound by shared-memory bandwidth.)
        bool pe1_pending = false;
        if (tl + 1 < n_tiles_total) {
          mbar_wait(&bars->s_empty, tl & 1);
          tc_fence_after();
          mma1(tl + 1, 0, 1);
          pe1_pending = true;
        }
        {
          const long long tw0 = clock64();
          while (!mbar_test(&bars->p_full, tl & 1)) {      // P stored, accumulator rescaled if the maxima moved
            if (pe1_pending && mbar_test(&bars->pek_full, pcnt & 1) && mbar_test(&bars->b_full[bcnt & 1], (bcnt >> 1) & 1)) {
              mma1(tl + 1, 1, 2);
              pe1_pending = false;
            }
            if (clock64() - tw0 > CSAM_MBAR_BUDGET) __trap();
          }
        }
        CSAM_TR(30, tl);
        tc_fence_after();
        // B = [P_hi | P_lo] as ONE MN-major operand of N = 128 (the lo half is the next 64-wide atom, 16 KB further):
        // X_hi^T is fetched once for both products, 2 MMAs (64 + 48 clocks) per step instead of 3 x 48
        const uint32_t pd2 = umma_desc_lo(smem_u32(smem + T2I_OFF_P), 16384);
        const uint32_t pd1 = umma_desc_lo(smem_u32(smem + T2I_OFF_P), 8192);
#pragma unroll
        for (int fb = 0; fb < 2; ++fb) {
          const uint32_t xd = umma_desc_lo(smem_u32(smem + 2 * fb * T2I_SLOT), T2I_SLOT);   // LBO: next 64 features
          const uint32_t d = tmem_base + 128 + fb * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k) {           // 16 keys per step = 16 rows of 128 B
            umma_f16_w(d, xd + 128 * k, pd2 + 128 * k, idesc2w, (ti | k) ? 1u : 0u);
            umma_f16_w(d, xd + (16384 >> 4) + 128 * k, pd1 + 128 * k, idesc2, 1u);
          }
          umma_commit(&bars->x_empty[fb]);
        }
        umma_commit(&bars->pv_done);
        if (pe1_pending) mma1(tl + 1, 1, 2);
        if (tl + 1 < n_tiles_total) mma1(tl + 1, 2, 6);
      }
    }
  } else if (warp == 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax over keys (16 warps, four threads per key row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // 16 columns per thread: with 8 warps (32 columns each) every scheduler had two warps to cover the tcgen05.ld,
    // MUFU and shared-store latencies with, and the softmax took 3.3 k of the 9.2 k clocks of a tile (timeline)
    const int wq = (warp - 4) & 3;                 // TMEM lane quadrant == warp % 4
    const int cq = (warp - 4) >> 2;                // which 16 of the 64 (head, token) columns
    const int r = wq * 32 + lane;
    const int et = threadIdx.x - 128;              // 0..511
    const uint32_t lane_addr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float* my_m = st_m + cq * 16;
    // accumulator work (rescale, final read-out): this warp owns lanes 32 wq.. of feature half cq >> 1
    const uint32_t acc_addr = lane_addr + 128 + (cq >> 1) * 128;
    int tl = 0;
    for (int p = blockIdx.x; p < a.P; p += gridDim.x) {
      float lsum[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) lsum[c] = 0.f;
      for (int ti = 0; ti < 32; ++ti, ++tl) {
        if (warp == 4 && lane == 0) CSAM_TR(40, tl);
        mbar_wait(&bars->s_full, tl & 1);
        if (warp == 4 && lane == 0) CSAM_TR(41, tl);
        tc_fence_after();
        uint32_t raw[16];
        {
          uint32_t raw2[16];
          tmem_ld16(lane_addr + cq * 16, raw);
          tmem_ld16(lane_addr + 64 + cq * 16, raw2);       // the X_hi x B1_lo part of the scores
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) raw[c] = __float_as_uint(__uint_as_float(raw[c]) + __uint_as_float(raw2[c]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_empty);
        bool waited_pv = false;
        while (true) {
          // cheap pass first: does any score exceed its column's (stale) maximum by more than 2^TAU ?
          int viol = 0;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float2 mm = *reinterpret_cast<const float2*>(my_m + c);
            viol |= (__uint_as_float(raw[c]) > mm.x + T2I_TAU) | (__uint_as_float(raw[c + 1]) > mm.y + T2I_TAU);
          }
          int any;
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\tbar.red.or.pred q, 2, 512, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                       : "=r"(any) : "r"(viol) : "memory");
          if (!any) break;
          // some column outgrew its (stale) maximum: exact column maxima of this tile, new m, rescale factors
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float v = warp_max(__uint_as_float(raw[c]));
            if (lane == 0) st_wmax[wq * 64 + cq * 16 + c] = v;
          }
          asm volatile("bar.sync 2, 512;" ::: "memory");
          if (et < 64) {
            const float mo = st_m[et];
            const float mn = fmaxf(fmaxf(mo, fmaxf(st_wmax[et], st_wmax[64 + et])), fmaxf(st_wmax[128 + et], st_wmax[192 + et]));
            st_alpha[et] = ex2f_approx(mo - mn);     // first tile of a prompt: exp2(-inf) = 0
            st_m[et] = mn;
          }
          asm volatile("bar.sync 2, 512;" ::: "memory");
          if (ti > 0) {
            if (!waited_pv) { mbar_wait(&bars->pv_done, (tl - 1) & 1); tc_fence_after(); waited_pv = true; }
            // this warp's 32 lanes x one of the two 64-column blocks (A / B) of its feature half
#pragma unroll 1
            for (int hh = 0; hh < 4; ++hh) {         // 16 columns at a time: this rare path must not set the register budget
              uint32_t o[16];
              const uint32_t ad = acc_addr + (cq & 1) * 64 + hh * 16;
              tmem_ld16(ad, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * st_alpha[hh * 16 + c]);
              tmem_st16(ad, o);
            }
            tmem_st_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) lsum[c] *= st_alpha[cq * 16 + c];
          }
        }
        if (warp == 4 && lane == 0) CSAM_TR(42, tl);
        if (tl > 0 && !waited_pv) { mbar_wait(&bars->pv_done, (tl - 1) & 1); tc_fence_after(); }   // P buffer free
        if (warp == 4 && lane == 0) CSAM_TR(43, tl);
        uint8_t* pb = smem + T2I_OFF_P + r * 128;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = u * 8 + 2 * k;
            const float2 mm = *reinterpret_cast<const float2*>(my_m + c);
            const float p0 = ex2f_approx(__uint_as_float(raw[c]) - mm.x);
            const float p1 = ex2f_approx(__uint_as_float(raw[c + 1]) - mm.y);
            lsum[c] += p0;
            lsum[c + 1] += p1;
            __half2 h2, l2;
            split_h2_nc(p0, p1, h2, l2);
            ph[k] = *reinterpret_cast<const uint32_t*>(&h2);
            pl[k] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          const int off = ((cq * 2 + u) ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(pb + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(pb + 16384 + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        if (lane == 0) CSAM_TR(70 + (warp - 4), tl);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bars->p_full); CSAM_TR(50 + (warp - 4), tl); }
      }
      // ---- end of prompt: column sums over all keys, then XBAR = accumulator / l
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        // per-warp partial sums, added in a FIXED order below: a shared-memory atomicAdd here made the denominators
        // (and with them every score downstream) depend on the arrival order of the four warps, i.e. differ in the
        // last bit from run to run.  st_wmax is free here: its readers passed the last barrier of the final tile.
        const float v = warp_sum(lsum[c]);
        if (lane == 0) st_wmax[wq * 64 + cq * 16 + c] = v;
      }
      mbar_wait(&bars->pv_done, (tl - 1) & 1);
      tc_fence_after();
      asm volatile("bar.sync 2, 512;" ::: "memory");
      float* xo = a.xbar + (size_t)p * 64 * 256;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        // feature row r of half cq >> 1, (head, token) columns 32 (cq & 1) + 16 hh ..+16: block A + block B
        uint32_t oa[16], ob[16];
        const int c0 = (cq & 1) * 32 + hh * 16;
        tmem_ld16(acc_addr + c0, oa);
        tmem_ld16(acc_addr + 64 + c0, ob);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int col = c0 + c;
          const float l = ((st_wmax[col] + st_wmax[64 + col]) + st_wmax[128 + col]) + st_wmax[192 + col];
          xo[(size_t)col * 256 + (cq >> 1) * 128 + r] = (__uint_as_float(oa[c]) + __uint_as_float(ob[c])) / l;
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (et < 64) { st_m[et] = -INFINITY; st_l[et] = 0.f; }
      asm volatile("bar.sync 2, 512;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// out[p, i, o] = Wv[o, :] . xbar[p, (o/16)*8 + i, :] + bv[o]   (the v projection applied to the pooled keys)
__global__ void __launch_bounds__(128)
dec_t2i_out_kernel(const float* __restrict__ xbar, const float* __restrict__ wv_t /*[256,128]*/, const float* __restrict__ bv,
                   float* out_f32, __half* out_hi, __half* out_lo) {
  const int p = blockIdx.x, o = threadIdx.x, h = o >> 4;
  const float* xb = xbar + ((size_t)p * 64 + h * 8) * 256;
  float acc[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = 0.f;
  for (int c = 0; c < 256; ++c) {
    const float w = wv_t[c * 128 + o];
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[i] = fmaf(w, xb[i * 256 + c], acc[i]);
  }
  const float b = bv ? bv[o] : 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const size_t oo = ((size_t)p * 7 + i) * 128 + o;
    const float v = acc[i] + b;
    if (out_f32) out_f32[oo] = v;
    if (out_hi) store_pair(out_hi, out_lo, oo, v);
  }
}

}  // namespace csam

using namespace csam;

#ifdef CSAM_TRACE
// copies the [tag][tile] stamp table (160 x 64 clock64 values, 0 = not stamped) to `host` and clears it
extern "C" __attribute__((visibility("default"))) int csam_debug_trace(unsigned long long* host, int max_entries) {
  cudaDeviceSynchronize();
  const int n = TRACE_TAGS * TRACE_TILES < max_entries ? TRACE_TAGS * TRACE_TILES : max_entries;
  cudaMemcpyFromSymbol(host, g_trace, (size_t)n * 8);
  static unsigned long long zeros[TRACE_TAGS * TRACE_TILES];
  cudaMemcpyToSymbol(g_trace, zeros, sizeof(zeros));
  return n;
}
#endif

extern "C" int csam_dec_fold_i2t(const float* kt, const float* vt, int P, const float* wq, const float* wo, const float* bo,
                                 void* b1_hi, void* b1_lo, void* b2_hi, void* b2_lo, void* stream) {
  CSAM_REQUIRE(kt && vt && wq && wo && b1_hi && b1_lo && b2_hi && b2_lo && P > 0, "csam_dec_fold_i2t: bad args");
  CSAM_REQUIRE((reinterpret_cast<uintptr_t>(wo) & 15) == 0 && (reinterpret_cast<uintptr_t>(b2_hi) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(b2_lo) & 15) == 0,
               "csam_dec_fold_i2t: 16-byte alignment");
  dec_fold_i2t_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(kt, vt, wq, wo, bo, static_cast<__half*>(b1_hi),
                                                           static_cast<__half*>(b1_lo), static_cast<__half*>(b2_hi),
                                                           static_cast<__half*>(b2_lo));
  return check_launch("dec_fold_i2t_kernel");
}

extern "C" int csam_dec_i2t_layer(const csam_i2t_layer_args* a, void* stream) {
  CSAM_REQUIRE(a && a->x_hi && a->x_lo && a->peq_hi && a->peq_lo && a->b1_hi && a->b1_lo && a->b2_hi && a->b2_lo &&
                   a->gamma && a->beta && a->out_hi && a->out_lo,
               "csam_dec_i2t_layer: null operand (h16 pairs with both halves are required)");
  CSAM_REQUIRE(a->P > 0 && a->P <= (1 << 16), "csam_dec_i2t_layer: prompt count");
  auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  CSAM_REQUIRE(al32(a->x_hi) && al32(a->x_lo) && al32(a->out_hi) && al32(a->out_lo) && al32(a->peq_hi) &&
                   al32(a->peq_lo) && al32(a->b1_hi) && al32(a->b1_lo) && al32(a->b2_hi) && al32(a->b2_lo),
               "csam_dec_i2t_layer: 32-byte alignment");
  const uint64_t xrows = a->x_shared ? 4096ull : (uint64_t)a->P * 4096ull;
  CUtensorMap tx_hi, tx_lo, tq_hi, tq_lo, tb1_hi, tb1_lo, tb2_hi, tb2_lo;
  if (make_tmap_2d_f16(&tx_hi, a->x_hi, xrows, 256, 256, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tx_lo, a->x_lo, xrows, 256, 256, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tq_hi, a->peq_hi, 4096, 128, 128, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tq_lo, a->peq_lo, 4096, 128, 128, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tb1_hi, a->b1_hi, (uint64_t)a->P * 64, 384, 384, 64, 64)) return 1;
  if (make_tmap_2d_f16(&tb1_lo, a->b1_lo, (uint64_t)a->P * 64, 384, 384, 64, 64)) return 1;
  if (make_tmap_2d_f16(&tb2_hi, a->b2_hi, (uint64_t)a->P * 256, 64, 64, 256, 64)) return 1;
  if (make_tmap_2d_f16(&tb2_lo, a->b2_lo, (uint64_t)a->P * 256, 64, 64, 256, 64)) return 1;
  CSAM_DYN_SMEM(dec_i2t_layer_kernel, I2T_SMEM_BYTES, "dec_i2t_layer_kernel");
  I2TParams p;
  p.x_shared = a->x_shared ? 1 : 0;
  p.tiles = a->P * 32;
  p.pf = 0;
  if (const char* env = getenv("CSAM_I2T_PF")) p.pf = atoi(env);      // experiments
  p.bias = a->bias; p.gamma = a->gamma; p.beta = a->beta; p.eps = a->eps;
  p.out_hi = static_cast<__half*>(a->out_hi); p.out_lo = static_cast<__half*>(a->out_lo);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = p.tiles < sms ? p.tiles : sms;
  if (const char* env = getenv("CSAM_I2T_GRID")) grid = atoi(env) > 0 && atoi(env) < grid ? atoi(env) : grid;   // experiments
  dec_i2t_layer_kernel<<<grid, I2T_THREADS, I2T_SMEM_BYTES, (cudaStream_t)stream>>>(tx_hi, tx_lo, tq_hi, tq_lo, tb1_hi,
                                                                                   tb1_lo, tb2_hi, tb2_lo, p);
  return check_launch("dec_i2t_layer_kernel");
}

extern "C" int csam_dec_fold_t2i(const float* qt, int P, const float* wk, void* b1_hi, void* b1_lo, void* stream) {
  CSAM_REQUIRE(qt && wk && b1_hi && b1_lo && P > 0, "csam_dec_fold_t2i: bad args");
  dec_fold_i2t_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(qt, nullptr, wk, nullptr, nullptr, static_cast<__half*>(b1_hi),
                                                           static_cast<__half*>(b1_lo), nullptr, nullptr);
  return check_launch("dec_fold_i2t_kernel");
}

extern "C" int csam_dec_t2i(const csam_t2i_args* a, void* stream) {
  CSAM_REQUIRE(a && a->x_hi && a->x_lo && a->pek_hi && a->pek_lo && a->b1_hi && a->b1_lo && a->xbar && a->wv_t &&
                   (a->out_f32 || a->out_hi),
               "csam_dec_t2i: null operand (h16 pairs with both halves are required)");
  CSAM_REQUIRE(a->P > 0 && a->P <= (1 << 16), "csam_dec_t2i: prompt count");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  CSAM_REQUIRE(al16(a->x_hi) && al16(a->x_lo) && al16(a->pek_hi) && al16(a->pek_lo) && al16(a->b1_hi) && al16(a->b1_lo),
               "csam_dec_t2i: 16-byte alignment");
  const uint64_t xrows = a->x_shared ? 4096ull : (uint64_t)a->P * 4096ull;
  CUtensorMap tx_hi, tx_lo, tk_hi, tk_lo, tb1_hi, tb1_lo;
  if (make_tmap_2d_f16(&tx_hi, a->x_hi, xrows, 256, 256, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tx_lo, a->x_lo, xrows, 256, 256, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tk_hi, a->pek_hi, 4096, 128, 128, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tk_lo, a->pek_lo, 4096, 128, 128, 128, 64)) return 1;
  if (make_tmap_2d_f16(&tb1_hi, a->b1_hi, (uint64_t)a->P * 64, 384, 384, 64, 64)) return 1;
  if (make_tmap_2d_f16(&tb1_lo, a->b1_lo, (uint64_t)a->P * 64, 384, 384, 64, 64)) return 1;
  CSAM_DYN_SMEM(dec_t2i_kernel, T2I_SMEM_BYTES, "dec_t2i_kernel");
  T2IParams p;
  p.x_shared = a->x_shared ? 1 : 0;
  p.P = a->P;
  p.pf = 0;        // measured: 1320 us without, 1338 us with the next-tile L2 prefetch at P = 1024
  if (const char* env = getenv("CSAM_T2I_PF")) p.pf = atoi(env);      // experiments
  p.xbar = a->xbar;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a->P < sms ? a->P : sms;
  dec_t2i_kernel<<<grid, T2I_THREADS, T2I_SMEM_BYTES, (cudaStream_t)stream>>>(tx_hi, tx_lo, tk_hi, tk_lo, tb1_hi, tb1_lo, p);
  if (check_launch("dec_t2i_kernel")) return 1;
  dec_t2i_out_kernel<<<a->P, 128, 0, (cudaStream_t)stream>>>(a->xbar, a->wv_t, a->bv, a->out_f32,
                                                             static_cast<__half*>(a->out_hi), static_cast<__half*>(a->out_lo));
  return check_launch("dec_t2i_out_kernel");
}
