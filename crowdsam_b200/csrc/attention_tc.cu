// attention_tc.cu — K-WATTN / K-GATTN: flash attention on tcgen05 tensor cores (head dim 64).
//
// One CTA = 128 queries of one (group, head); keys stream through in tiles of 64.
//   warp 0      TMA producer   Q once, then K/V tiles into a 3-stage ring (128B-swizzled boxes)
//   warp 3      MMA issuer #1  S = Q K^T  (M128 N64 K64, both operands K-major)  -> TMEM S[2]
//   warp 1      MMA issuer #2  O += P V   (M128 N64 K64, V is the MN-major B operand) -> TMEM O
//   warp 2      TMEM allocator
//   warps 4..   softmax        one warpgroup per query tile (NQ = 1 or 2 tiles per CTA share the K/V ring);
//                              one thread per query row: tcgen05.ld S, scale + decomposed rel-pos bias, online
//                              softmax in fp32, P written to shared memory as fp16 (optionally hi/lo) in the UMMA
//                              K-major swizzled layout.  O accumulates in TMEM across key tiles; the row maximum
//                              may lag by 2^8 and O is rescaled in place (tcgen05.ld -> * alpha -> tcgen05.st) only
//                              when a row really outgrows it.
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
  // P buffers per query tile.  With two query tiles per CTA the probabilities are single-buffered (the
  // softmax warpgroup folds O(j-1) into its accumulator -- which proves P V(j-1) has retired -- BEFORE it
  // stores P(j)), which leaves room for a 4-stage K/V ring next to two Q tiles.
  static constexpr int PBUF = (NQ == 2) ? 1 : 2;
  static constexpr int THREADS = 128 + 128 * NQ;
  static constexpr int Q_BYTES = AT_BM * AT_HD * 2;          // 16 KB per operand half per query tile
  static constexpr int KV_TILE = AT_BN * AT_HD * 2;          // 8 KB
  static constexpr int STAGE_BYTES = NOPS * 2 * KV_TILE;     // K(hi,lo) then V(hi,lo)
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;          // 16 KB per operand half
  // K/V ring depth = whatever shared memory is left (<= 8).  A stage is released when P V of its tile has
  // retired and S of the NEXT tile is issued one tile early, so a ring of n stages gives the TMA only n - 2
  // tile times to land: with 3 stages the kernel ran at TMA latency (0.92 us per key tile with all math removed).
  static constexpr int FIXED_BYTES = NQ * NOPS * Q_BYTES + NQ * PBUF * PNOPS * P_BYTES +
                                     (BIAS == 1 ? NQ * AT_BM * AT_REL_LD * 4 : 0) + 512 + 1024;
  static constexpr int STAGES_FIT = (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 2, "attention K/V ring");
  static constexpr int OFF_KV = NQ * NOPS * Q_BYTES;
  static constexpr int OFF_P = OFF_KV + STAGES * STAGE_BYTES;
  static constexpr int OFF_REL = OFF_P + NQ * PBUF * PNOPS * P_BYTES;
  static constexpr int OFF_BAR = OFF_REL + (BIAS == 1 ? NQ * AT_BM * AT_REL_LD * 4 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_COLS = 256 * NQ;                 // per query tile: S0 S1 O0 O1, 64 columns each
  static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
};

struct AttnBars {
  uint64_t q_full;
  uint64_t kv_full[8], kv_empty[8];
  uint64_t s_full[2][2], s_empty[2][2], p_full[2][2], pv_done[2][2];   // [query tile][key tile parity]
  uint32_t tmem_slot;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// The running maximum is allowed to lag the true row maximum by up to 2^AT_TAU (probabilities stay <= 256, well
// inside fp16): O then accumulates IN TMEM across key tiles and is only read back when a row's maximum really
// grows -- a handful of times per row instead of once per key tile.  TMEM reads were the bound of the previous
// version (S and O read back every tile: 64 KB per key tile per SM at 64 B/clk = 1024 of ~2400 cycles).
constexpr float AT_TAU = 8.f;

// BIAS: 0 none, 1 window (S = 14, table in shared memory), 2 global (S = 64 == key tile, registers)
template <int SPLIT, int BIAS, bool PLO, int NQ>
__global__ void __launch_bounds__(128 + 128 * NQ, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap t_hi, const __grid_constant__ CUtensorMap t_lo,
                        csam_attn_args a, const float* __restrict__ rel) {
  using Cfg = AttnCfg<SPLIT, BIAS, PLO, NQ>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int PBUF = Cfg::PBUF;
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
        // mbarrier serialise as shared-memory atomics (measured 0.3 us per barrier)
        mbar_init(&bars->s_full[t][b], 1); mbar_init(&bars->s_empty[t][b], 4);
        mbar_init(&bars->p_full[t][b], 4);
        mbar_init(&bars->pv_done[t][b], 1);
      }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(&bars->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;   // per query tile t: S0 [0,64) S1 [64,128) O [128,192), + 256 t

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
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
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer #1: S = Q K^T
    // Two issuing threads (this one and warp 1 for P V): a narrow MMA (N = 64) executes in 32 cycles, less than
    // one thread needs to issue it, so a single issuer held the tensor pipe back (x3: 20 MMAs per key tile).
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN, 0, 0);
      mbar_wait(&bars->q_full, 0);
      tc_fence_after();
      uint32_t qd[NQ];
#pragma unroll
      for (int t = 0; t < NQ; ++t) qd[t] = umma_desc_lo(smem_u32(smem + t * Cfg::NOPS * Cfg::Q_BYTES), 16);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        mbar_wait(&bars->kv_full[st], (j / STAGES) & 1);
        const uint32_t kd = umma_desc_lo(smem_u32(smem + Cfg::OFF_KV + st * Cfg::STAGE_BYTES), 16);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          mbar_wait(&bars->s_empty[t][j & 1], ((j >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + t * 256 + (j & 1) * AT_BN;
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) {
            umma_f16_w(d, qd[t] + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
            if (SPLIT == 3) {
              umma_f16_w(d, qd[t] + (Cfg::Q_BYTES >> 4) + 2 * k, kd + 2 * k, idesc_s, 1u);
              umma_f16_w(d, qd[t] + 2 * k, kd + (Cfg::KV_TILE >> 4) + 2 * k, idesc_s, 1u);
            }
          }
          umma_commit(&bars->s_full[t][j & 1]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer #2: O += P V
    if (elect_one()) {
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, AT_HD, 0, 1);
      for (int j = 0; j < n_tiles; ++j) {
        const int b = j & 1;
        const uint32_t use = (j >> 1) & 1;
        const int st = j % STAGES;
        const uint32_t vd = umma_desc_lo(smem_u32(smem + Cfg::OFF_KV + st * Cfg::STAGE_BYTES + Cfg::NOPS * Cfg::KV_TILE), 8192);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          mbar_wait(&bars->p_full[t][b], use);      // P(j) stored (so S(j), K(j), V(j) have landed), any rescale of O finished
          tc_fence_after();
          const uint32_t pd = umma_desc_lo(smem_u32(smem + Cfg::OFF_P + (t * PBUF + (PBUF == 2 ? b : 0)) * Cfg::PNOPS * Cfg::P_BYTES), 16);
          const uint32_t d = tmem_base + t * 256 + 128;      // O accumulates across key tiles
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            umma_f16_w(d, pd + 2 * k, vd + 128 * k, idesc_pv, (j | k) ? 1u : 0u);
            if (SPLIT == 3) {
              umma_f16_w(d, pd + 2 * k, vd + (Cfg::KV_TILE >> 4) + 128 * k, idesc_pv, 1u);
              if (PLO) umma_f16_w(d, pd + (Cfg::P_BYTES >> 4) + 2 * k, vd + 128 * k, idesc_pv, 1u);
            }
          }
          umma_commit(&bars->pv_done[t][b]);
        }
        // K(j) was consumed by S(j), which completed before softmax(j) could publish P(j): this thread's commit
        // (tracking only its own P V MMAs) therefore also covers the K half of the stage
        umma_commit(&bars->kv_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax
    const int qt = (warp - 4) >> 2;                // query tile of this warpgroup
    const int wq = (warp - 4) & 3;                 // TMEM lane quadrant == warp % 4
    const int r = wq * 32 + lane;                  // query row in the tile == TMEM lane
    const int q = q0 + qt * AT_BM + r;
    const int qc = min(q, a.tokens - 1);
    const uint32_t lane_addr = tmem_base + qt * 256 + ((uint32_t)(wq * 32) << 16);
    const uint32_t o_addr = lane_addr + 128;
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
      // 28 floats per query = 7 float4 (rows are 112 B apart, 16-byte aligned): all loads in flight at once instead
      // of 28 dependent scalar round trips in the prologue of a CTA that only has 4 key tiles of work
      float4 t4[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) t4[i] = *reinterpret_cast<const float4*>(relq + 4 * i);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        rel_r[4 * i + 0] = t4[i].x * LOG2E; rel_r[4 * i + 1] = t4[i].y * LOG2E;
        rel_r[4 * i + 2] = t4[i].z * LOG2E; rel_r[4 * i + 3] = t4[i].w * LOG2E;
      }
    }
    float m = -INFINITY, l = 0.f;      // m: the (possibly stale) maximum the probabilities are taken against
    auto wait_pv = [&](int j) {        // P V(j) retired: O holds tiles 0..j, the P buffer of tile j is free
      mbar_wait(&bars->pv_done[qt][j & 1], (j >> 1) & 1);
      tc_fence_after();
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      mbar_wait(&bars->s_full[qt][b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t raw[64];
      tmem_ld32(lane_addr + b * AT_BN, raw);
      tmem_ld32(lane_addr + b * AT_BN + 32, raw + 32);
      tmem_ld_wait();
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
      if (j == 0) {
        m = tmax;
      } else if (__any_sync(0xffffffffu, tmax > m + AT_TAU)) {
        // some row of this warp outgrew its stale maximum: every row of the warp moves to its true maximum and
        // rescales its O row in TMEM (tcgen05.ld / .st are warp-wide).  O must hold all of tiles 0..j-1 first.
        const float m_new = fmaxf(m, tmax);
        const float alpha = ex2_approx(m - m_new);
        wait_pv(j - 1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          tmem_ld32(o_addr + hh * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 32; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * alpha);
          tmem_st32(o_addr + hh * 32, o);
        }
        tmem_st_wait();
        l *= alpha;
        m = m_new;
      }
      float psum = 0.f;
      uint32_t ph[32];
      uint32_t pl[(SPLIT == 3 && PLO) ? 32 : 1];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        float p0 = (BIAS == 0) ? fmaf(s[c], scale2, -m) : s[c] - m;
        float p1 = (BIAS == 0) ? fmaf(s[c + 1], scale2, -m) : s[c + 1] - m;
        p0 = ex2_approx(p0);
        p1 = ex2_approx(p1);
        psum += p0 + p1;
        const __half2 h2 = __floats2half2_rn(p0, p1);   // one packed cvt per pair
        ph[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        if (SPLIT == 3 && PLO) {
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
          pl[c >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
      }
      l += psum;
      if (j >= PBUF) wait_pv(j - PBUF);              // the P buffer about to be overwritten has been consumed
      uint8_t* pb = smem + Cfg::OFF_P + (qt * PBUF + (PBUF == 2 ? b : 0)) * Cfg::PNOPS * Cfg::P_BYTES + r * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int off = (u ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(pb + off) = make_uint4(ph[4 * u], ph[4 * u + 1], ph[4 * u + 2], ph[4 * u + 3]);
        if (SPLIT == 3 && PLO)
          *reinterpret_cast<uint4*>(pb + Cfg::P_BYTES + off) = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
      }
      fence_proxy_async();                 // generic-proxy writes of P -> visible to the tensor core
      tc_fence_before();                   // orders this warp's tcgen05.st (rescale) before the issuer's MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[qt][b]);
    }
    wait_pv(n_tiles - 1);
    const float inv = 1.0f / l;
    __half* ohi = static_cast<__half*>(a.out_hi);
    __half* olo = static_cast<__half*>(a.out_lo);
    const size_t oo = ((size_t)row_base + q) * a.ld_out + (size_t)h * AT_HD;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o[32];
      tmem_ld32(o_addr + hh * 32, o);
      tmem_ld_wait();
      if (q < a.tokens) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          float v8[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v8[t] = __uint_as_float(o[d + t]) * inv;
          store_pair8(ohi, olo, oo + hh * 32 + d, v8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// =====================================================================================================
// TS variant: Q and P live in TENSOR MEMORY and enter the MMAs as the A operand.
// With both operands in shared memory a 128 x 64 x 16 MMA reads 4 KB (A) + 2 KB (B) per 32 cycles = 192 B/clk, more
// than the 128 B/clk a CTA gets from shared memory: the SS kernel above is shared-memory-bandwidth bound (measured
// 2540 cycles per key tile pair against 1280 cycles of tensor work).  Here the MMAs read only K / V (64 B/clk):
//   TMEM per query tile (256 columns):  [0,64) S0 | P0   [64,128) S1 | P1   [128,192) O   [192,224) Q hi   [224,256) Q lo
//   Q: each softmax thread copies its own row from global memory into TMEM once (tcgen05.st);
//   P(j): fp16 pairs written over the first 32 columns of S(j) after S(j) has been read (no shared memory, no proxy
//         fence); S(j+2) may only be issued once P V(j) has retired.
// Shared memory holds nothing but the K/V ring (6-7 stages).
// G2 (global blocks, head dim 64, V as one fp16): the 64 per-query column terms of the decomposed rel-pos bias live in
// shared memory instead of 64 registers per thread, the stage holds K hi | K lo | V hi only (24 KB, 3 stages), and TWO
// CTAs share an SM (<= 128 registers, 256 TMEM columns, ~108 KB each): one CTA's load -> S -> softmax -> P V chain runs
// under the other's, which is what the two query tiles per CTA do for the bias-free kernel (the round-1 review asked for
// the bias out of the registers: one tile per CTA at 186 registers ran at 32-39 % tensor pipe).
template <int SPLIT, int BIAS, int NQ, int HD = 64, int G2 = 0>
struct AttnTsCfg {
  static constexpr int NOPS = (SPLIT == 3) ? 2 : 1;
  static constexpr int THREADS = 128 + 128 * NQ;
  static_assert(HD == 64 || (HD == 80 && NQ == 1), "head dim 64, or 80 with one query tile per CTA");
  static_assert(G2 == 0 || (BIAS == 2 && NQ == 1 && HD == 64 && SPLIT == 3), "G2: split global blocks, head dim 64");
  static constexpr int NA = (HD + 63) / 64;                  // 64-wide swizzle atoms per K / V tile (80 -> 2, the second 1/4 used)
  static constexpr int KV_TILE = NA * AT_BN * 64 * 2;        // 8 KB per atom
  static constexpr int V_OPS = G2 ? 1 : NOPS;                // V tiles per stage (hi | lo, or hi only)
  static constexpr int STAGE_BYTES = (NOPS + V_OPS) * KV_TILE;
  static constexpr int REL_LD = G2 ? 68 : AT_REL_LD;         // 68 floats = 272 B: 16-byte aligned rows, conflict-free LDS.128
  static constexpr int REL_BYTES = (BIAS == 1 || G2) ? NQ * AT_BM * REL_LD * 4 : 0;
  // window blocks, head dim 64: the decomposed rel-pos terms q . Rh[qh - kh + 13], q . Rw[qw - kw + 13] are computed IN the
  // kernel, as one small TS-mode MMA per query tile against both tables (27 + 27 rows of 64, hi | lo, K-major SW128)
  static constexpr int TAB_BYTES = (BIAS == 1 && HD == 64) ? 2 * 8192 : 0;
  static constexpr int STAGES_FIT = (227 * 1024 - REL_BYTES - TAB_BYTES - 512 - 1024) / STAGE_BYTES;
  // window blocks with ONE query tile per CTA (the default, see use_nq2): 4 key tiles in all, so a 2-stage ring is
  // enough and two CTAs (256 TMEM columns, ~82 KB, <= 128 registers each) share an SM
  static constexpr int STAGES = (BIAS == 1 && NQ == 1 && HD == 64) ? 2 : G2 ? 3 : (STAGES_FIT > 8 ? 8 : STAGES_FIT);
  static constexpr int OFF_TAB = STAGES * STAGE_BYTES;        // 1024-byte aligned (stages are multiples of 16 KB)
  static constexpr int OFF_REL = OFF_TAB + TAB_BYTES;
  static constexpr int OFF_BAR = OFF_REL + REL_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  // TMEM columns per query tile: S0 | P0 [0,64)  S1 | P1 [64,128)  O [128,128+HD)  Q hi  Q lo (HD/2 columns each)
  static constexpr int COL_O = 128, COL_QH = 128 + HD, COL_QL = COL_QH + HD / 2;
  static constexpr int TSTRIDE = (HD == 64) ? 256 : 512;
  static constexpr int TMEM_COLS = TSTRIDE * NQ;
};

struct AttnTsBars {
  uint64_t kv_full[8], kv_empty[8];
  uint64_t q_ready[2], s_full[2][2], p_full[2][2], pv_done[2][2];   // [query tile]([key tile parity])
  uint64_t rel_full[2], rel_done[2];                                // in-kernel rel-pos terms: computed / read back
  uint32_t tmem_slot;
};

template <int SPLIT, int BIAS, int NQ, int HD, int G2>
__global__ void __launch_bounds__(128 + 128 * NQ, ((BIAS == 1 && NQ == 1 && HD == 64) || G2) ? 2 : 1)
vit_attention_ts_kernel(const __grid_constant__ CUtensorMap t_hi, const __grid_constant__ CUtensorMap t_lo,
                        csam_attn_args a, const float* __restrict__ rel, int n_full, int n_single) {
  using Cfg = AttnTsCfg<SPLIT, BIAS, NQ, HD, G2>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TS = Cfg::TSTRIDE;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  AttnTsBars* bars = reinterpret_cast<AttnTsBars*>(smem + Cfg::OFF_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + Cfg::OFF_REL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items in launch order: first every (query-tile pair, head, group) item, then the single-tile items.  With
  // 336 equal pair items on 148 SMs a third of the last wave's SMs idle; sized as 2 full waves of pairs plus a
  // wave of (shorter) single-tile CTAs the tail is one short wave instead (DINOv2: 288 + 96 CTAs).
  const int HG = a.heads * a.groups;
  const int lin = blockIdx.x;
  int nqa, q0, hg;
  if (lin < n_full * HG) {
    nqa = NQ; q0 = (lin % n_full) * (AT_BM * NQ); hg = lin / n_full;
  } else {
    const int l2 = lin - n_full * HG;
    nqa = 1; q0 = n_full * (AT_BM * NQ) + (l2 % n_single) * AT_BM; hg = l2 / n_single;
  }
  const int h = hg % a.heads, g = hg / a.heads;
  const int D = a.heads * HD;
  const int row_base = g * a.tokens;
  const int n_tiles = (a.tokens + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&t_hi);
    if (SPLIT == 3) tma_prefetch_desc(&t_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
    for (int t = 0; t < NQ; ++t) {
      mbar_init(&bars->q_ready[t], 4);
      mbar_init(&bars->rel_full[t], 1);
      mbar_init(&bars->rel_done[t], 4);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[t][b], 1);
        mbar_init(&bars->p_full[t][b], 4);
        mbar_init(&bars->pv_done[t][b], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(&bars->tmem_slot);
  // rel == nullptr on a window launch: the rel-pos terms are computed here (see MMA issuer #1) instead of being read
  // from a precomputed [groups * heads * tokens, 28] array (relpos_win14_kernel: 39 us per layer, 20 layers)
  const bool rel_inkernel = (BIAS == 1) && Cfg::TAB_BYTES > 0 && rel == nullptr;
  // p_split < 0: V enters P V as ONE fp16 (no V_lo load, one MMA per k-step instead of two).  P is a single fp16 already,
  // so the product carries 2^-12 relative per factor instead of per probability only
  const bool v_lo = (SPLIT == 3) && a.p_split >= 0 && !G2;      // (G2 stages have no room for V_lo)
  if (rel_inkernel) {
    // B operand [64 rows = 27 of Rh | 27 of Rw | 10 zero][K = 64], hi then lo, 128B-swizzled K-major
    uint8_t* tab = smem + Cfg::OFF_TAB;
    for (int i = threadIdx.x; i < 64 * 64; i += Cfg::THREADS) {
      const int n = i >> 6, k = i & 63;
      float v = 0.f;
      if (n < 27) v = a.rel_h[n * 64 + k];
      else if (n < 54) v = a.rel_w[(n - 27) * 64 + k];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const int off = n * 128 + ((((k >> 3) ^ (n & 7))) << 4) + (k & 7) * 2;
      *reinterpret_cast<__half*>(tab + off) = hi;
      *reinterpret_cast<__half*>(tab + 8192 + off) = lo;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (K / V only)
    if (elect_one()) {
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        mbar_wait(&bars->kv_empty[st], ((j / STAGES) & 1) ^ 1);
        uint8_t* sk = smem + st * Cfg::STAGE_BYTES;
        uint8_t* sv = sk + Cfg::NOPS * Cfg::KV_TILE;
        mbar_expect_tx(&bars->kv_full[st], (Cfg::NOPS + (v_lo ? Cfg::NOPS : 1)) * Cfg::KV_TILE);
        const int row = row_base + j * AT_BN;
#pragma unroll
        for (int at = 0; at < Cfg::NA; ++at) {      // head dim 80: a second 64-wide box of which 16 columns are used
          tma_load_2d(sk + at * 8192, &t_hi, &bars->kv_full[st], D + h * HD + at * 64, row);
          tma_load_2d(sv + at * 8192, &t_hi, &bars->kv_full[st], 2 * D + h * HD + at * 64, row);
          if (SPLIT == 3) {
            tma_load_2d(sk + Cfg::KV_TILE + at * 8192, &t_lo, &bars->kv_full[st], D + h * HD + at * 64, row);
            if (v_lo) tma_load_2d(sv + Cfg::KV_TILE + at * 8192, &t_lo, &bars->kv_full[st], 2 * D + h * HD + at * 64, row);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer #1: S = Q K^T, Q from TMEM
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN, 0, 0);
#pragma unroll
      for (int t = 0; t < NQ; ++t)
        if (t < nqa) mbar_wait(&bars->q_ready[t], 0);
      tc_fence_after();
      if (rel_inkernel) {
        // T[128 queries, 64] = Q (tensor memory) x [Rh ; Rw]^T into the S1 columns, which the main loop does not touch
        // before S(1): the softmax threads pick their 14 + 14 terms out of it and release it through rel_done
        const uint32_t td = umma_desc_lo(smem_u32(smem + Cfg::OFF_TAB), 16);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          if (t >= nqa) break;
          const uint32_t d = tmem_base + t * TS + AT_BN;
          const uint32_t qa = tmem_base + t * TS + Cfg::COL_QH;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ts(d, qa + 8 * k, td + 2 * k, idesc_s, k ? 1u : 0u);
            if (SPLIT == 3) {
              umma_f16_ts(d, qa + HD / 2 + 8 * k, td + 2 * k, idesc_s, 1u);
              umma_f16_ts(d, qa + 8 * k, td + (8192 >> 4) + 2 * k, idesc_s, 1u);
            }
          }
          umma_commit(&bars->rel_full[t]);
        }
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        mbar_wait(&bars->kv_full[st], (j / STAGES) & 1);
        const uint32_t kd = umma_desc_lo(smem_u32(smem + st * Cfg::STAGE_BYTES), 16);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          if (t >= nqa) break;
          // S(j) overwrites the buffer that held S(j-2) and then P(j-2): P V(j-2) must have retired
          if (j >= 2) mbar_wait(&bars->pv_done[t][j & 1], ((j - 2) >> 1) & 1);
          if (j == 1 && rel_inkernel) mbar_wait(&bars->rel_done[t], 0);     // the rel-pos terms have been read out of S1
          tc_fence_after();
          const uint32_t d = tmem_base + t * TS + (j & 1) * AT_BN;
          const uint32_t qa = tmem_base + t * TS + Cfg::COL_QH;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t kk = kd + (k >> 2) * (8192 >> 4) + 2 * (k & 3);      // swizzle atom k/4, 32-byte step k%4
            umma_f16_ts(d, qa + 8 * k, kk, idesc_s, k ? 1u : 0u);
            if (SPLIT == 3) {
              umma_f16_ts(d, qa + HD / 2 + 8 * k, kk, idesc_s, 1u);
              umma_f16_ts(d, qa + 8 * k, kk + (Cfg::KV_TILE >> 4), idesc_s, 1u);
            }
          }
          umma_commit(&bars->s_full[t][j & 1]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer #2: O += P V, P from TMEM
    if (elect_one()) {
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, HD, 0, 1);
      for (int j = 0; j < n_tiles; ++j) {
        const int b = j & 1;
        const int st = j % STAGES;
        const uint32_t vd = umma_desc_lo(smem_u32(smem + st * Cfg::STAGE_BYTES + Cfg::NOPS * Cfg::KV_TILE), 8192);
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          if (t >= nqa) break;
          mbar_wait(&bars->p_full[t][b], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t pa = tmem_base + t * TS + b * AT_BN;      // P(j): 32 columns of fp16 pairs over S(j)
          const uint32_t d = tmem_base + t * TS + Cfg::COL_O;
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            umma_f16_ts(d, pa + 8 * k, vd + 128 * k, idesc_pv, (j | k) ? 1u : 0u);
            if (SPLIT == 3 && v_lo) umma_f16_ts(d, pa + 8 * k, vd + (Cfg::KV_TILE >> 4) + 128 * k, idesc_pv, 1u);
          }
          umma_commit(&bars->pv_done[t][b]);
        }
        umma_commit(&bars->kv_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax
    const int qt = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;
    const int r = wq * 32 + lane;
    const int q = q0 + qt * AT_BM + r;
    const int qc = min(q, a.tokens - 1);
    const uint32_t lane_addr = tmem_base + qt * TS + ((uint32_t)(wq * 32) << 16);
    const uint32_t o_addr = lane_addr + Cfg::COL_O;
    constexpr float LOG2E = 1.4426950408889634f;
    const float scale2 = a.scale * LOG2E;
    if (qt < nqa) {                      // the second warpgroup of a single-tile CTA has nothing to do
    {
      // this thread's Q row -> TMEM (A operand of every S MMA of the CTA)
      const __half* qh = static_cast<const __half*>(a.qkv_hi) + ((size_t)row_base + qc) * a.ld_qkv + (size_t)h * HD;
#pragma unroll
      for (int i = 0; i < HD / 16; ++i) {          // 16 elements = 32 bytes = 8 TMEM columns at a time
        uint32_t w[8];
        ldg256(qh + i * 16, w);
        tmem_st8(lane_addr + Cfg::COL_QH + i * 8, w);
      }
      if (SPLIT == 3) {
        const __half* ql = static_cast<const __half*>(a.qkv_lo) + ((size_t)row_base + qc) * a.ld_qkv + (size_t)h * HD;
#pragma unroll
        for (int i = 0; i < HD / 16; ++i) {
          uint32_t w[8];
          ldg256(ql + i * 16, w);
          tmem_st8(lane_addr + Cfg::COL_QL + i * 8, w);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->q_ready[qt]);
    }
    float relw[(BIAS == 2 && !G2) ? 64 : 1];
    const float* relq = nullptr;
    float* rel_r = rel_s + (qt * AT_BM + r) * Cfg::REL_LD;
    if (BIAS != 0 && !rel_inkernel) relq = rel + (((size_t)g * a.heads + h) * a.tokens + qc) * 2 * a.S;
    if (BIAS == 2 && G2) {
      // this row's 64 column terms -> its own shared-memory row (read back by the same thread: no barrier needed)
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(relq + 64 + c);
        *reinterpret_cast<float4*>(rel_r + c) = make_float4(t4.x * LOG2E, t4.y * LOG2E, t4.z * LOG2E, t4.w * LOG2E);
      }
    } else if (BIAS == 2) {
#pragma unroll
      for (int c = 0; c < 64; ++c) relw[c] = relq[64 + c] * LOG2E;
    }
    if (BIAS == 1 && rel_inkernel) {
      // this row's 54 products q . Rh[n], q . Rw[n] out of tensor memory; the 14 + 14 it needs are
      // Rh[qh - kh + 13] and Rw[qw - kw + 13] (image_encoder.py get_rel_pos / add_decomposed_rel_pos)
      float tv[64];
      {
        uint32_t t0[32], t1[32];
        mbar_wait(&bars->rel_full[qt], 0);
        tc_fence_after();
        tmem_ld32(lane_addr + AT_BN, t0);
        tmem_ld32(lane_addr + AT_BN + 32, t1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->rel_done[qt]);
#pragma unroll
        for (int c = 0; c < 32; ++c) { tv[c] = __uint_as_float(t0[c]); tv[32 + c] = __uint_as_float(t1[c]); }
      }
      const int qh = qc / 14, qw = qc % 14;
#pragma unroll 1
      for (int k = 0; k < 14; ++k) {
        rel_r[k] = tv[qh - k + 13] * LOG2E;             // dynamic index: tv lives in local memory, once per CTA
        rel_r[14 + k] = tv[27 + qw - k + 13] * LOG2E;
      }
    } else if (BIAS == 1) {
      float4 t4[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) t4[i] = *reinterpret_cast<const float4*>(relq + 4 * i);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        rel_r[4 * i + 0] = t4[i].x * LOG2E; rel_r[4 * i + 1] = t4[i].y * LOG2E;
        rel_r[4 * i + 2] = t4[i].z * LOG2E; rel_r[4 * i + 3] = t4[i].w * LOG2E;
      }
    }
    float m = -INFINITY, l = 0.f;
    auto wait_pv = [&](int j) {
      mbar_wait(&bars->pv_done[qt][j & 1], (j >> 1) & 1);
      tc_fence_after();
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      mbar_wait(&bars->s_full[qt][b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t raw[64];
      tmem_ld32(lane_addr + b * AT_BN, raw);
      tmem_ld32(lane_addr + b * AT_BN + 32, raw + 32);
      tmem_ld_wait();
      float s[64];
      const int key0 = j * AT_BN;
      if (BIAS == 2) {
        const float bh = relq[j] * LOG2E;
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = fmaf(__uint_as_float(raw[c]), scale2, bh + (G2 ? rel_r[c] : relw[c]));
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
      // four independent running maxima: one serial chain of 32 dependent FMNMX3 (~150 clocks) sat in front of the
      // first exponential of every key tile
      float tm0 = -INFINITY, tm1 = -INFINITY, tm2 = -INFINITY, tm3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        tm0 = fmaxf(tm0, fmaxf(s[c], s[c + 1]));
        tm1 = fmaxf(tm1, fmaxf(s[c + 2], s[c + 3]));
        tm2 = fmaxf(tm2, fmaxf(s[c + 4], s[c + 5]));
        tm3 = fmaxf(tm3, fmaxf(s[c + 6], s[c + 7]));
      }
      float tmax = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
      if (BIAS == 0) tmax *= scale2;
      if (j == 0) {
        m = tmax;
      } else if (__any_sync(0xffffffffu, tmax > m + AT_TAU)) {
        const float m_new = fmaxf(m, tmax);
        const float alpha = ex2_approx(m - m_new);
        wait_pv(j - 1);
#pragma unroll
        for (int hh = 0; hh < HD / 16; ++hh) {
          uint32_t o[16];
          tmem_ld16(o_addr + hh * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * alpha);
          tmem_st16(o_addr + hh * 16, o);
        }
        l *= alpha;
        m = m_new;
      }
      // (Measured, round 2: evaluating every fourth exponential as a polynomial on the FMA pipe -- the FlashAttention-4
      //  trick against the MUFU.EX2 limit -- made this kernel 5 % SLOWER, and so did nothing for it shortening the
      //  dependency chains above: the key-tile period is the MMAs' 1280 clocks PLUS the 1024 clocks the two softmax
      //  warpgroups need to read S back at the TMEM read rate of 64 B/clk, which do not overlap.)
      float psum = 0.f;
      uint32_t ph[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        float p0 = (BIAS == 0) ? fmaf(s[c], scale2, -m) : s[c] - m;
        float p1 = (BIAS == 0) ? fmaf(s[c + 1], scale2, -m) : s[c + 1] - m;
        p0 = ex2_approx(p0);
        p1 = ex2_approx(p1);
        psum += p0 + p1;
        const __half2 h2 = __floats2half2_rn(p0, p1);
        ph[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      l += psum;
      tmem_st32(lane_addr + b * AT_BN, ph);          // P(j) over the first half of S(j)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[qt][b]);
    }
    wait_pv(n_tiles - 1);
    const float inv = 1.0f / l;
    __half* ohi = static_cast<__half*>(a.out_hi);
    __half* olo = static_cast<__half*>(a.out_lo);
    const size_t oo = ((size_t)row_base + q) * a.ld_out + (size_t)h * HD;
#pragma unroll
    for (int hh = 0; hh < HD / 16; ++hh) {
      uint32_t o[16];
      tmem_ld16(o_addr + hh * 16, o);
      tmem_ld_wait();
      if (q < a.tokens) {
#pragma unroll
        for (int d = 0; d < 16; d += 8) {
          float v8[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v8[t] = __uint_as_float(o[d + t]) * inv;
          store_pair8(ohi, olo, oo + hh * 16 + d, v8);
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int SPLIT, int BIAS, int NQ, int HD = 64, int G2 = 0>
static int launch_attn_ts(const csam_attn_args* a, const float* rel, cudaStream_t st) {
  using Cfg = AttnTsCfg<SPLIT, BIAS, NQ, HD, G2>;
  CUtensorMap t_hi, t_lo;
  const uint64_t rows = (uint64_t)a->groups * a->tokens;
  const uint64_t cols = 3ull * a->heads * a->hd;
  if (make_tmap_2d_f16(&t_hi, a->qkv_hi, rows, cols, a->ld_qkv, 64, 64)) return 1;
  t_lo = t_hi;
  if (SPLIT == 3 && make_tmap_2d_f16(&t_lo, a->qkv_lo, rows, cols, a->ld_qkv, 64, 64)) return 1;
  auto kern = vit_attention_ts_kernel<SPLIT, BIAS, NQ, HD, G2>;
  CSAM_DYN_SMEM(kern, Cfg::SMEM_BYTES, "vit_attention_ts_kernel");
  // query tiles -> n_full items of NQ tiles + n_single items of one tile (per head and group)
  const int tiles = (a->tokens + AT_BM - 1) / AT_BM;
  const int HG = a->heads * a->groups;
  int n_full = (tiles + NQ - 1) / NQ, n_single = 0;
  if (NQ == 2 && n_full * HG > 148) {
    const int waves = (n_full * HG) / 148;               // full waves of pair items
    const int p = (waves * 148) / HG;                    // pairs per (head, group) that fit those waves
    const int s1 = tiles - 2 * p;
    static const int tail_env = getenv("CSAM_ATTN_TAIL") ? atoi(getenv("CSAM_ATTN_TAIL")) : 1;
    if (tail_env && p > 0 && s1 > 0 && s1 * HG <= 148 && (n_full * HG) % 148 != 0) { n_full = p; n_single = s1; }
  }
  dim3 grid((n_full + n_single) * HG);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(t_hi, t_lo, *a, rel, n_full, n_single > 0 ? n_single : 1);
  return check_launch("vit_attention_ts_kernel");
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
  CSAM_DYN_SMEM(kern, Cfg::SMEM_BYTES, "vit_attention_tc_kernel");
  dim3 grid((a->tokens + AT_BM * NQ - 1) / (AT_BM * NQ), a->heads, a->groups);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(t_hi, t_lo, *a, rel);
  return check_launch("vit_attention_tc_kernel");
}

// Two query tiles per CTA (two softmax warpgroups sharing one K/V ring) when there is more than one tile.
// One softmax warp per scheduler is latency-bound (~2400 cycles per key tile against 640 cycles of MMA); a second
// warpgroup overlaps those chains.  CSAM_ATTN_NQ=1 forces the single-tile kernel (measurement only).
static bool use_nq2(const csam_attn_args* a, int bias) {
  static const int env = getenv("CSAM_ATTN_NQ") ? atoi(getenv("CSAM_ATTN_NQ")) : 0;
  static const int env_win = getenv("CSAM_ATTN_WIN_NQ") ? atoi(getenv("CSAM_ATTN_WIN_NQ")) : 0;
  if (a->tokens <= AT_BM) return false;
  // window blocks (196 tokens = two query tiles): ONE tile per CTA with a 2-stage ring, so that two CTAs share an SM
  // and overlap each other's load -> S -> softmax -> PV chains (4 key tiles per CTA: the chain, not the MMAs, is what a
  // window CTA spends its time on).  Measured: attention class 9.08 / 9.07 against 9.23 / 9.43 ms per step for
  // two-tile CTAs (CSAM_ATTN_WIN_NQ=2 selects those).
  if (bias == 1) return env_win == 2;
  return env != 1;
}

int vit_attention_tc(const csam_attn_args* a, cudaStream_t st) {
  CSAM_REQUIRE(a->hd == 64 || a->hd == 80, "csam_vit_attention(tcgen05): head dim 64 or 80 (use impl=1 for others)");
  CSAM_REQUIRE((a->ld_qkv % 8) == 0 && (a->ld_out % 8) == 0, "csam_vit_attention: strides must be multiples of 8");
  CSAM_REQUIRE((reinterpret_cast<uintptr_t>(a->qkv_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_hi) & 15) == 0,
               "csam_vit_attention: 16-byte alignment");
  const bool split = a->qkv_lo != nullptr;
  CSAM_REQUIRE(!split || a->out_lo, "csam_vit_attention: split input needs split output");
  int bias = 0;
  const float* rel = nullptr;
  if (a->rel_h) {
    CSAM_REQUIRE(a->S == 14 || a->S == 64, "csam_vit_attention(tcgen05): S must be 14 (window) or 64 (global)");
    bias = a->S == 14 ? 1 : 2;
    // window blocks with head dim 64 on the TS kernel compute their rel-pos terms themselves (rel stays nullptr)
    static const int rel_env = getenv("CSAM_ATTN_REL_INKERNEL") ? atoi(getenv("CSAM_ATTN_REL_INKERNEL")) : 1;
    const bool al32q = (reinterpret_cast<uintptr_t>(a->qkv_hi) & 31) == 0 && (a->ld_qkv % 16) == 0 &&
                       (!split || (reinterpret_cast<uintptr_t>(a->qkv_lo) & 31) == 0);
    static const int ts_env0 = getenv("CSAM_ATTN_TS") ? atoi(getenv("CSAM_ATTN_TS")) : 1;
    const bool inkernel = rel_env && bias == 1 && a->hd == 64 && a->tokens == 196 && ts_env0 && al32q && !(split && a->p_split > 0);
    if (!inkernel) {
      if (compute_relpos(a, st)) return 1;
      rel = a->scratch;
    }
  }
  {
    // Q / P in tensor memory (default): needs 32-byte aligned Q rows for the row copy; p_split keeps the SS kernel
    static const int ts_env = getenv("CSAM_ATTN_TS") ? atoi(getenv("CSAM_ATTN_TS")) : 1;
    const bool al32 = (reinterpret_cast<uintptr_t>(a->qkv_hi) & 31) == 0 && (a->ld_qkv % 16) == 0 &&
                      (!split || (reinterpret_cast<uintptr_t>(a->qkv_lo) & 31) == 0);
    if (a->hd == 80) {
      // ViT-H: one query tile per CTA (TMEM: 128 + 80 + 80 columns), K / V tiles of two swizzle atoms
      CSAM_REQUIRE(al32 && !(split && a->p_split > 0), "csam_vit_attention(tcgen05): head dim 80 needs 32-byte aligned rows, no p_split");
      if (split) {
        if (bias == 0) return launch_attn_ts<3, 0, 1, 80>(a, rel, st);
        if (bias == 1) return launch_attn_ts<3, 1, 1, 80>(a, rel, st);
        return launch_attn_ts<3, 2, 1, 80>(a, rel, st);
      }
      if (bias == 0) return launch_attn_ts<1, 0, 1, 80>(a, rel, st);
      if (bias == 1) return launch_attn_ts<1, 1, 1, 80>(a, rel, st);
      return launch_attn_ts<1, 2, 1, 80>(a, rel, st);
    }
    if (ts_env && al32 && !(split && a->p_split > 0)) {
      const bool nq2 = use_nq2(a, bias) && bias != 2;
      if (split) {
        if (bias == 0) return nq2 ? launch_attn_ts<3, 0, 2>(a, rel, st) : launch_attn_ts<3, 0, 1>(a, rel, st);
        if (bias == 1) return nq2 ? launch_attn_ts<3, 1, 2>(a, rel, st) : launch_attn_ts<3, 1, 1>(a, rel, st);
        // global blocks with V as one fp16 (the default): two CTAs per SM, column bias terms in shared memory
        static const int g2_env = getenv("CSAM_ATTN_G2") ? atoi(getenv("CSAM_ATTN_G2")) : 1;
        if (g2_env && a->p_split < 0 && a->S == 64) return launch_attn_ts<3, 2, 1, 64, 1>(a, rel, st);
        return launch_attn_ts<3, 2, 1>(a, rel, st);
      }
      if (bias == 0) return nq2 ? launch_attn_ts<1, 0, 2>(a, rel, st) : launch_attn_ts<1, 0, 1>(a, rel, st);
      if (bias == 1) return nq2 ? launch_attn_ts<1, 1, 2>(a, rel, st) : launch_attn_ts<1, 1, 1>(a, rel, st);
      return launch_attn_ts<1, 2, 1>(a, rel, st);
    }
  }
  if (split && a->p_split > 0) {
    if (bias == 0) return launch_attn_tc<3, 0, true, 1>(a, rel, st);
    if (bias == 1) return launch_attn_tc<3, 1, true, 1>(a, rel, st);
    return launch_attn_tc<3, 2, true, 1>(a, rel, st);
  }
  if (split) {
    if (bias == 0) return use_nq2(a, 0) ? launch_attn_tc<3, 0, false, 2>(a, rel, st) : launch_attn_tc<3, 0, false, 1>(a, rel, st);
    if (bias == 1) return use_nq2(a, 1) ? launch_attn_tc<3, 1, false, 2>(a, rel, st) : launch_attn_tc<3, 1, false, 1>(a, rel, st);
    return launch_attn_tc<3, 2, false, 1>(a, rel, st);
  }
  if (bias == 0) return use_nq2(a, 0) ? launch_attn_tc<1, 0, false, 2>(a, rel, st) : launch_attn_tc<1, 0, false, 1>(a, rel, st);
  if (bias == 1) return use_nq2(a, 1) ? launch_attn_tc<1, 1, false, 2>(a, rel, st) : launch_attn_tc<1, 1, false, 1>(a, rel, st);
  return launch_attn_tc<1, 2, false, 1>(a, rel, st);
}

}  // namespace csam
