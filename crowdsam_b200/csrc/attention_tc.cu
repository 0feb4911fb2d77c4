// attention_tc.cu — K-WATTN / K-GATTN: flash attention on tcgen05 tensor cores (head dim 64).
//
// One CTA = 128 queries of one (group, head); keys stream through in tiles of 64.
//   warp 0      TMA producer   Q once, then K/V tiles into a 3-stage ring (128B-swizzled boxes)
//   warp 1      MMA issuer     S = Q K^T  (M128 N64 K64, both operands K-major)  -> TMEM S[2]
//                              O_j = P V  (M128 N64 K64, V is the MN-major B operand) -> TMEM O[2]
//   warp 2      TMEM allocator
//   warps 4..7  softmax        one thread per query row: tcgen05.ld S, scale + decomposed rel-pos
//                              bias, online softmax in fp32, P written to shared memory as an fp16
//                              hi/lo pair in the UMMA K-major swizzled layout, O accumulated in
//                              registers with the usual rescale.
// S(j+1) is issued before softmax(j) finishes (two S buffers), so the tensor pipe overlaps the
// exponentials.  With the hi/lo split every product is 3 MMAs (fp32-level accuracy).
// Reference: image_encoder.py:224-240,325-361; dinov2/layers/attention.py:56-69.
#include "common.cuh"

namespace csam {

int compute_relpos(const csam_attn_args* a, cudaStream_t st);   // attention_simt.cu

constexpr int AT_BM = 128, AT_BN = 64, AT_HD = 64;
constexpr int AT_REL_LD = 29;   // 28 rel-pos values per query (S = 14) padded to an odd stride

// PLO: also split the probabilities P into hi + lo (3 MMAs for P V); without it P is a single fp16
// (relative error 2^-12 per probability, measured 2e-5 at the encoder output) and P V needs 2 MMAs.
// NQ: query tiles (of 128 rows) per CTA, each with its own softmax warpgroup and TMEM S/O buffers, sharing
// the K/V ring.  One softmax warp per scheduler cannot issue fast enough to keep the tensor pipe busy
// (ncu: issue slots 33 % active, tensor 26 %, XU 21 %); two warpgroups on two query tiles double that.
template <int SPLIT, int BIAS, bool PLO, int NQ>
struct AttnCfg {
  static constexpr int NOPS = (SPLIT == 3) ? 2 : 1;
  static constexpr int PNOPS = (SPLIT == 3 && PLO) ? 2 : 1;
  static constexpr int THREADS = 128 + 128 * NQ;
  static constexpr int Q_BYTES = AT_BM * AT_HD * 2;          // 16 KB per operand half per query tile
  static constexpr int KV_TILE = AT_BN * AT_HD * 2;          // 8 KB
  static constexpr int STAGE_BYTES = NOPS * 2 * KV_TILE;     // K(hi,lo) then V(hi,lo)
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;          // 16 KB per operand half
  // K/V ring depth = whatever shared memory is left (<= 8).  A stage is released when P V of its tile has
  // retired and S of the NEXT tile is issued one tile early, so a ring of n stages gives the TMA only n - 2
  // tile times to land: with 3 stages the kernel ran at TMA latency (0.92 us per key tile with all math removed).
  static constexpr int FIXED_BYTES = NQ * NOPS * Q_BYTES + NQ * 2 * PNOPS * P_BYTES +
                                     (BIAS == 1 ? NQ * AT_BM * AT_REL_LD * 4 : 0) + 512 + 1024;
  static constexpr int STAGES_FIT = (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 2, "attention K/V ring");
  static constexpr int OFF_KV = NQ * NOPS * Q_BYTES;
  static constexpr int OFF_P = OFF_KV + STAGES * STAGE_BYTES;
  static constexpr int OFF_REL = OFF_P + NQ * 2 * PNOPS * P_BYTES;
  static constexpr int OFF_BAR = OFF_REL + (BIAS == 1 ? NQ * AT_BM * AT_REL_LD * 4 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_COLS = 256 * NQ;                 // per query tile: S0 S1 O0 O1, 64 columns each
  static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
};

struct AttnBars {
  uint64_t q_full;
  uint64_t kv_full[8], kv_empty[8];
  uint64_t s_full[2][2], s_empty[2][2], p_full[2][2], o_full[2][2], o_empty[2][2];   // [query tile][buffer]
  uint32_t tmem_slot;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// BIAS: 0 none, 1 window (S = 14, table in shared memory), 2 global (S = 64 == key tile, registers)
template <int SPLIT, int BIAS, bool PLO, int NQ>
__global__ void __launch_bounds__(128 + 128 * NQ, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap t_hi, const __grid_constant__ CUtensorMap t_lo,
                        csam_attn_args a, const float* __restrict__ rel, int dbg) {
  using Cfg = AttnCfg<SPLIT, BIAS, PLO, NQ>;
  constexpr int STAGES = Cfg::STAGES;
  // Dynamic shared memory is the only shared allocation of this kernel, so it starts at the (1024-byte
  // aligned) base of the CTA's window; keeping `smem` a plain __shared__ array (no integer round-trip) lets
  // the compiler emit LDS/STS instead of generic LD/ST for every staging access.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + Cfg::OFF_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + Cfg::OFF_REL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (AT_BM * NQ);
  const int D = a.heads * AT_HD;
  const int row_base = g * a.tokens;
  const int n_tiles = (a.tokens + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&t_hi);
    if (SPLIT == 3) tma_prefetch_desc(&t_lo);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
    for (int t = 0; t < NQ; ++t)
      for (int b = 0; b < 2; ++b) {
        // softmax-side arrivals are one elected lane per warp (4 per warpgroup): 128 threads arriving on one
        // mbarrier serialise as shared-memory atomics -- three of those per key tile cost 0.9 us, more than
        // the math (measured: removing ALL softmax math did not change the kernel time)
        mbar_init(&bars->s_full[t][b], 1); mbar_init(&bars->s_empty[t][b], 4);
        mbar_init(&bars->p_full[t][b], 4);
        mbar_init(&bars->o_full[t][b], 1); mbar_init(&bars->o_empty[t][b], 4);
      }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;   // per query tile t: S0 [0,64) S1 [64,128) O0 [128,192) O1 [192,256) + 256 t

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(&bars->q_full, NQ * Cfg::NOPS * Cfg::Q_BYTES);
      for (int t = 0; t < NQ; ++t)
        for (int half = 0; half < 2; ++half) {
          uint8_t* sq = smem + t * Cfg::NOPS * Cfg::Q_BYTES + half * 8192;
          const int row = row_base + q0 + t * AT_BM + half * 64;
          tma_load_2d(sq, &t_hi, &bars->q_full, h * AT_HD, row);
          if (SPLIT == 3) tma_load_2d(sq + Cfg::Q_BYTES, &t_lo, &bars->q_full, h * AT_HD, row);
        }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(&bars->kv_empty[st], ph ^ 1);
        uint8_t* sk = smem + Cfg::OFF_KV + st * Cfg::STAGE_BYTES;
        uint8_t* sv = sk + Cfg::NOPS * Cfg::KV_TILE;
        mbar_expect_tx(&bars->kv_full[st], Cfg::STAGE_BYTES);
        const int row = row_base + j * AT_BN;
        tma_load_2d(sk, &t_hi, &bars->kv_full[st], D + h * AT_HD, row);
        tma_load_2d(sv, &t_hi, &bars->kv_full[st], 2 * D + h * AT_HD, row);
        if (SPLIT == 3) {
          tma_load_2d(sk + Cfg::KV_TILE, &t_lo, &bars->kv_full[st], D + h * AT_HD, row);
          tma_load_2d(sv + Cfg::KV_TILE, &t_lo, &bars->kv_full[st], 2 * D + h * AT_HD, row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, AT_HD, 0, 1);
      auto issue_s = [&](int j) {          // S(j) of every query tile of this CTA
        const int st = j % STAGES;
        mbar_wait(&bars->kv_full[st], (j / STAGES) & 1);
        const uint32_t sk = smem_u32(smem + Cfg::OFF_KV + st * Cfg::STAGE_BYTES);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          mbar_wait(&bars->s_empty[t][j & 1], ((j >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t sq = smem_u32(smem + t * Cfg::NOPS * Cfg::Q_BYTES);
          const uint32_t d = tmem_base + t * 256 + (j & 1) * AT_BN;
#pragma unroll
          for (int k = 0; k < ((dbg & 64) && j > 1 ? 0 : AT_HD / 16); ++k) {
            const uint64_t q_hi = umma_desc_sw128(sq + k * 32, 16, 1024);
            const uint64_t k_hi = umma_desc_sw128(sk + k * 32, 16, 1024);
            umma_f16(d, q_hi, k_hi, idesc_s, k ? 1u : 0u);
            if (SPLIT == 3) {
              const uint64_t q_lo = umma_desc_sw128(sq + Cfg::Q_BYTES + k * 32, 16, 1024);
              const uint64_t k_lo = umma_desc_sw128(sk + Cfg::KV_TILE + k * 32, 16, 1024);
              umma_f16(d, q_lo, k_hi, idesc_s, 1u);
              umma_f16(d, q_hi, k_lo, idesc_s, 1u);
            }
          }
          umma_commit(&bars->s_full[t][j & 1]);
        }
      };
      mbar_wait(&bars->q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int b = j & 1;
        const uint32_t use = (j >> 1) & 1;
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int st = j % STAGES;
        const uint32_t sv = smem_u32(smem + Cfg::OFF_KV + st * Cfg::STAGE_BYTES + Cfg::NOPS * Cfg::KV_TILE);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          mbar_wait(&bars->p_full[t][b], use);
          mbar_wait(&bars->o_empty[t][b], use ^ 1);
          tc_fence_after();
          const uint32_t sp = smem_u32(smem + Cfg::OFF_P + (t * 2 + b) * Cfg::PNOPS * Cfg::P_BYTES);
          const uint32_t d = tmem_base + t * 256 + 128 + b * AT_HD;
#pragma unroll
          for (int k = 0; k < ((dbg & 32) && j > 1 ? 0 : AT_BN / 16); ++k) {
            const uint64_t p_hi = umma_desc_sw128(sp + k * 32, 16, 1024);
            const uint64_t v_hi = umma_desc_sw128(sv + k * 2048, 8192, 1024);
            umma_f16(d, p_hi, v_hi, idesc_pv, k ? 1u : 0u);
            if (SPLIT == 3) {
              const uint64_t v_lo = umma_desc_sw128(sv + Cfg::KV_TILE + k * 2048, 8192, 1024);
              umma_f16(d, p_hi, v_lo, idesc_pv, 1u);
              if (PLO) {
                const uint64_t p_lo = umma_desc_sw128(sp + Cfg::P_BYTES + k * 32, 16, 1024);
                umma_f16(d, p_lo, v_hi, idesc_pv, 1u);
              }
            }
          }
          umma_commit(&bars->o_full[t][b]);
        }
        umma_commit(&bars->kv_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / accumulate
    const int qt = (warp - 4) >> 2;                // query tile of this warpgroup
    const int wq = (warp - 4) & 3;                 // TMEM lane quadrant == warp % 4
    const int r = wq * 32 + lane;                  // query row in the tile == TMEM lane
    const int q = q0 + qt * AT_BM + r;
    const int qc = min(q, a.tokens - 1);
    const uint32_t lane_addr = tmem_base + qt * 256 + ((uint32_t)(wq * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const float scale2 = a.scale * LOG2E;
    float relw[BIAS == 2 ? 64 : 1];
    const float* relq = nullptr;
    float* rel_r = rel_s + (qt * AT_BM + r) * AT_REL_LD;
    if (BIAS != 0) relq = rel + (((size_t)g * a.heads + h) * a.tokens + qc) * 2 * a.S;
    if (BIAS == 2) {
#pragma unroll
      for (int c = 0; c < 64; ++c) relw[c] = relq[64 + c] * LOG2E;
    }
    if (BIAS == 1) {
      for (int i = 0; i < 28; ++i) rel_r[i] = relq[i] * LOG2E;
    }
    float acc[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

    auto accumulate_o = [&](int j, float alpha) {
      const int b = j & 1;
      mbar_wait(&bars->o_full[qt][b], (j >> 1) & 1);
      tc_fence_after();
      if (dbg & 4) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&bars->o_empty[qt][b]); acc[0] += alpha; return; }
      uint32_t o[64];
      tmem_ld32(lane_addr + 128 + b * AT_HD, o);
      tmem_ld32(lane_addr + 128 + b * AT_HD + 32, o + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_empty[qt][b]);
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) acc[d] = fmaf(acc[d], alpha, __uint_as_float(o[d]));
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      mbar_wait(&bars->s_full[qt][b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t raw[64];
      if (dbg & 8) {
#pragma unroll
        for (int c = 0; c < 64; ++c) raw[c] = __float_as_uint(0.01f * (float)(c + j));
      } else {
        tmem_ld32(lane_addr + b * AT_BN, raw);
        tmem_ld32(lane_addr + b * AT_BN + 32, raw + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty[qt][b]);
      // s[c] = logit * log2(e) (+ bias); without a bias the scale is folded into the exponent's FFMA
      float s[64];
      const int key0 = j * AT_BN;
      if (BIAS == 2) {
        const float bh = relq[j] * LOG2E;           // key tile j == key row kh = j (S == 64)
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = fmaf(__uint_as_float(raw[c]), scale2, bh + relw[c]);
      } else if (BIAS == 1) {
        int kh = key0 / 14, kw = key0 % 14;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const int khc = min(kh, 13);
          s[c] = fmaf(__uint_as_float(raw[c]), scale2, rel_r[khc] + rel_r[14 + kw]);
          if (++kw == 14) { kw = 0; ++kh; }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = __uint_as_float(raw[c]);
      }
      if (key0 + AT_BN > a.tokens) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (key0 + c >= a.tokens) s[c] = -INFINITY;
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) tmax = fmaxf(tmax, s[c]);
      if (BIAS == 0) tmax *= scale2;                 // scale2 > 0: max commutes with the scaling
      const float m_new = fmaxf(m, tmax);
      const float alpha = ex2_approx(m - m_new);
      float psum = 0.f;
      uint8_t* pb = smem + Cfg::OFF_P + (qt * 2 + b) * Cfg::PNOPS * Cfg::P_BYTES + r * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float pv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          pv[t] = (BIAS == 0) ? fmaf(s[u * 8 + t], scale2, -m_new) : s[u * 8 + t] - m_new;
          if (!(dbg & 1)) pv[t] = ex2_approx(pv[t]);
          psum += pv[t];
        }
        __align__(16) __half2 hi2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) hi2[t] = __floats2half2_rn(pv[2 * t], pv[2 * t + 1]);   // one packed cvt per pair
        const int off = (u ^ (r & 7)) << 4;
        if (!(dbg & 2) || u == (j & 7)) *reinterpret_cast<uint4*>(pb + off) = *reinterpret_cast<const uint4*>(hi2);
        if (SPLIT == 3 && PLO) {
          __align__(16) __half2 lo2[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 hf = __half22float2(hi2[t]);
            lo2[t] = __floats2half2_rn(pv[2 * t] - hf.x, pv[2 * t + 1] - hf.y);
          }
          *reinterpret_cast<uint4*>(pb + Cfg::P_BYTES + off) = *reinterpret_cast<const uint4*>(lo2);
        }
      }
      l = fmaf(l, alpha, psum);
      m = m_new;
      fence_proxy_async();                 // generic-proxy writes of P -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[qt][b]);
      if (j > 0) accumulate_o(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate_o(n_tiles - 1, alpha_prev);
    if (q < a.tokens) {
      const float inv = 1.0f / l;
      __half* ohi = static_cast<__half*>(a.out_hi);
      __half* olo = static_cast<__half*>(a.out_lo);
      const size_t oo = ((size_t)row_base + q) * a.ld_out + (size_t)h * AT_HD;
#pragma unroll
      for (int d = 0; d < AT_HD; d += 8) {
        float v8[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v8[t] = acc[d + t] * inv;
        store_pair8(ohi, olo, oo + d, v8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int SPLIT, int BIAS, bool PLO, int NQ>
static int launch_attn_tc(const csam_attn_args* a, const float* rel, cudaStream_t st) {
  using Cfg = AttnCfg<SPLIT, BIAS, PLO, NQ>;
  CUtensorMap t_hi, t_lo;
  const uint64_t rows = (uint64_t)a->groups * a->tokens;
  const uint64_t cols = 3ull * a->heads * a->hd;
  if (make_tmap_2d_f16(&t_hi, a->qkv_hi, rows, cols, a->ld_qkv, 64, 64)) return 1;
  t_lo = t_hi;
  if (SPLIT == 3 && make_tmap_2d_f16(&t_lo, a->qkv_lo, rows, cols, a->ld_qkv, 64, 64)) return 1;
  auto kern = vit_attention_tc_kernel<SPLIT, BIAS, PLO, NQ>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail("%s", "cudaFuncSetAttribute(smem) failed for vit_attention_tc_kernel");
    attr = true;
  }
  dim3 grid((a->tokens + AT_BM * NQ - 1) / (AT_BM * NQ), a->heads, a->groups);
  static const int dbg = getenv("CSAM_ATTN_DBG") ? atoi(getenv("CSAM_ATTN_DBG")) : 0;   // timing experiments only
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(t_hi, t_lo, *a, rel, dbg);
  return check_launch("vit_attention_tc_kernel");
}

// two query tiles per CTA when there is more than one tile and the kernel variant fits the register file
static bool use_nq2(const csam_attn_args* a) {
  if (const char* env = getenv("CSAM_ATTN_NQ")) return atoi(env) == 2 && a->tokens > AT_BM;
  return false;   // measured on B200 (DINOv2 shape): 723 us with two tiles vs 528 us with one -- see DESIGN.md section 7
}

int vit_attention_tc(const csam_attn_args* a, cudaStream_t st) {
  CSAM_REQUIRE(a->hd == 64, "csam_vit_attention(tcgen05): head dim 64 only (use impl=1 for others)");
  CSAM_REQUIRE((a->ld_qkv % 8) == 0 && (a->ld_out % 8) == 0, "csam_vit_attention: strides must be multiples of 8");
  CSAM_REQUIRE((reinterpret_cast<uintptr_t>(a->qkv_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_hi) & 15) == 0,
               "csam_vit_attention: 16-byte alignment");
  const bool split = a->qkv_lo != nullptr;
  CSAM_REQUIRE(!split || a->out_lo, "csam_vit_attention: split input needs split output");
  int bias = 0;
  const float* rel = nullptr;
  if (a->rel_h) {
    CSAM_REQUIRE(a->S == 14 || a->S == 64, "csam_vit_attention(tcgen05): S must be 14 (window) or 64 (global)");
    if (compute_relpos(a, st)) return 1;
    rel = a->scratch;
    bias = a->S == 14 ? 1 : 2;
  }
  if (split && a->p_split) {
    if (bias == 0) return launch_attn_tc<3, 0, true, 1>(a, rel, st);
    if (bias == 1) return launch_attn_tc<3, 1, true, 1>(a, rel, st);
    return launch_attn_tc<3, 2, true, 1>(a, rel, st);
  }
  if (split) {
    if (bias == 0) return use_nq2(a) ? launch_attn_tc<3, 0, false, 2>(a, rel, st) : launch_attn_tc<3, 0, false, 1>(a, rel, st);
    if (bias == 1) return launch_attn_tc<3, 1, false, 1>(a, rel, st);
    return launch_attn_tc<3, 2, false, 1>(a, rel, st);
  }
  if (bias == 0) return use_nq2(a) ? launch_attn_tc<1, 0, false, 2>(a, rel, st) : launch_attn_tc<1, 0, false, 1>(a, rel, st);
  if (bias == 1) return launch_attn_tc<1, 1, false, 1>(a, rel, st);
  return launch_attn_tc<1, 2, false, 1>(a, rel, st);
}

}  // namespace csam
