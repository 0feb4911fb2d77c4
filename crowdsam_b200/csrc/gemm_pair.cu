// gemm_pair.cu — K-GEMM for the large encoder shapes on CTA PAIRS: tcgen05.mma.cta_group::2, 256 x 256 pair tiles.
//
// Why: the single-CTA 128x128 tile of gemm.cu moves (128 + 128) rows x 64 k x 2 B x 2 (hi, lo) = 64 KB of operands
// into shared memory per k-block and spends 3 (hi*hi, lo*hi, hi*lo) x 4 x 64 = 768 tensor-pipe cycles on them:
// 85 B/clk per SM, twice what the L2 delivers chip-wide (~6300 B/clk = 42.6 B/clk per SM, B300_MICROARCH.md "LTS
// throughput cap"), which is why ncu shows the tensor pipe 52-57 % active on those launches.  A CTA pair computing a
// 256 x 256 tile needs per CTA only ITS 128 rows of A and ITS 128-row half of B per k-block (the MMA reads the other
// half of B from the peer's shared memory): the same 64 KB now feed 1536 cycles = 42.6 B/clk.
//
// Protocol (one CTA pair = cluster of 2; rank 0 is the leader):
//   producer warp (both CTAs)  waits its OWN empty[s]; TMA-loads its A / B boxes with .cta_group::2, completing
//                              the transaction bytes on the LEADER's full[s]; the leader arms expect_tx for both
//   MMA thread (leader only)   waits full[s], issues tcgen05.mma.cta_group::2 (M = 256, N = 256), then
//                              tcgen05.commit ... multicast::cluster to empty[s] of BOTH CTAs; after the last k-block
//                              commits tfull[buf] of both CTAs
//   epilogue (8 warps per CTA) each CTA drains its own 128 TMEM lanes x 256 columns; every warp then arrives on the
//                              LEADER's tempty[buf] (count 16) through shared::cluster
//   TMEM                       tcgen05.alloc.cta_group::2 by warp 2 of both CTAs (512 columns = 2 accumulator buffers);
//                              cluster barrier before the first remote access and before dealloc
// Epilogue = the row-per-lane ("direct") EPI_STD form of gemm.cu: bias / GELU / ReLU / LayerScale / row scale /
// fp32 residual (in place allowed) / row scatter / fp32 and h16-pair outputs.
// Replaces the same nn.Linear call sites as K-GEMM (image_encoder.py:212-213,236; common.py:21-26;
// dinov2/layers/attention.py:56-69, mlp.py:34-40) whenever M x N gives at least half a wave of pair tiles.
#include "gemm_shared.cuh"

namespace csam {

namespace {

constexpr int PM = 128;             // rows of A per CTA
constexpr int PBK = 64;
constexpr int P_A_BYTES = PM * PBK * 2;            // 16 KB
constexpr int P_EPI_WARPS = 8;
constexpr int P_THREADS = 128 + 32 * P_EPI_WARPS;
// PBN = pair tile N (256 or 128); each CTA stages PBN / 2 rows of W.  256: 64 KB stages, 3 of them; 128: 48 KB, 4.
template <int PBN>
struct PairCfg {
  static constexpr int W_BYTES = (PBN / 2) * PBK * 2;
  static constexpr int STAGE_BYTES = 2 * (P_A_BYTES + W_BYTES);   // hi + lo of both operands
  static constexpr int STAGES = PBN == 256 ? 3 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};
constexpr uint64_t TMA_EVICT_NORMAL = 0x1000000000000000ull;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// TMA load issued by either CTA of the pair into its OWN shared memory; the bytes are accounted on the LEADER's
// barrier (`leader_bar` = shared::cluster address of the barrier in the CTA of rank 0)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(TMA_EVICT_NORMAL)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair_512(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair_512(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}

template <int PBN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
                 const __grid_constant__ CUtensorMap tw_hi, const __grid_constant__ CUtensorMap tw_lo,
                 GemmEpi e, int K, int tiles_m, int tiles_n) {
  using Cfg = PairCfg<PBN>;
  constexpr int P_W_BYTES = Cfg::W_BYTES, P_STAGE_BYTES = Cfg::STAGE_BYTES, P_STAGES = Cfg::STAGES;
  constexpr int EC = PBN / 2;                    // columns per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + P_STAGES * P_STAGE_BYTES);   // used in the leader
  uint64_t* empty_bar = full_bar + P_STAGES;     // one set per CTA, signalled by the multicast commit
  uint64_t* tfull_bar = empty_bar + P_STAGES;    // [2] per CTA
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] used in the leader: 2 x P_EPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / PBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ta_hi); tma_prefetch_desc(&ta_lo);
    tma_prefetch_desc(&tw_hi); tma_prefetch_desc(&tw_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * P_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair_512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m0 = (t / tiles_n) * (2 * PM) + rank * PM;            // this CTA's 128 rows of A
        const int n0 = (t % tiles_n) * PBN + rank * (PBN / 2);          // this CTA's half of W's rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * P_STAGE_BYTES;
          uint8_t* sw = sa + 2 * P_A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);    // both CTAs' bytes
          const int k0 = kb * PBK;
          const uint32_t lbar = map_to_rank(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, &ta_hi, lbar, k0, m0);
          tma_load_2d_pair(sa + P_A_BYTES, &ta_lo, lbar, k0, m0);
          tma_load_2d_pair(sw, &tw_hi, lbar, k0, n0);
          tma_load_2d_pair(sw + P_W_BYTES, &tw_lo, lbar, k0, n0);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * PM, PBN, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
        const int buf = local & 1;
        const uint32_t bphase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + buf * PBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * P_STAGE_BYTES);
          const uint32_t sw = sa + 2 * P_A_BYTES;
          const uint32_t ad = umma_desc_lo(sa, 16);
          const uint32_t wd = umma_desc_lo(sw, 16);
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k) {
            umma_f16_pair(d_addr, ad + 2 * k, wd + 2 * k, idesc, (kb | k) ? 1u : 0u);                  // hi * hi
            umma_f16_pair(d_addr, ad + (P_A_BYTES >> 4) + 2 * k, wd + 2 * k, idesc, 1u);               // lo * hi
            umma_f16_pair(d_addr, ad + 2 * k, wd + (P_W_BYTES >> 4) + 2 * k, idesc, 1u);               // hi * lo
          }
          umma_commit_pair(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[buf]);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (each CTA: its 128 rows x 256 columns)
    const int ew = warp - 4;
    const int q = ew & 3;                        // TMEM lane quadrant == warp % 4
    const int ch = ew >> 2;                      // which half of the tile's columns
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int local = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      const int m0 = (t / tiles_n) * (2 * PM) + rank * PM;
      const int n0 = (t % tiles_n) * PBN;
      const int r = m0 + q * 32 + lane;
      const float rs = (e.row_scale && r < e.M) ? e.row_scale[r] : 1.f;
      int orow = -1;
      if (r < e.M) orow = e.row_map ? e.row_map[r] : r;
      const float* pres = nullptr;
      if (orow >= 0 && e.residual) pres = e.residual + (size_t)(e.res_mod > 0 ? (orow % e.res_mod) : orow) * e.ldr;
#pragma unroll 1
      for (int part = 0; part < EC / 64; ++part) {     // EC columns per warp in passes of 64 (register budget)
        const int cbase = n0 + ch * EC + part * 64;
        float res[64];
        if (pres) {
#pragma unroll
          for (int c = 0; c < 64; c += 8)
            if (cbase + c < e.N) ldg256f(pres + cbase + c, res + c);
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) res[c] = 0.f;
        }
        if (part == 0) {
          mbar_wait(&tfull_bar[buf], bphase);
          tc_fence_after();
        }
        const uint32_t col_addr = lane_addr + buf * PBN + ch * EC + part * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          const int col0 = cbase + c;
          if (col0 < e.N) {                      // warp-uniform (N is a multiple of 16)
            uint32_t raw[16];
            tmem_ld16(col_addr + c, raw);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f), cs = make_float4(1.f, 1.f, 1.f, 1.f);
              if (e.bias) b = *reinterpret_cast<const float4*>(e.bias + col0 + j);
              if (e.col_scale) cs = *reinterpret_cast<const float4*>(e.col_scale + col0 + j);
              v[j + 0] = apply_act(__uint_as_float(raw[j + 0]) * rs + b.x, e.act) * cs.x + res[c + j + 0];
              v[j + 1] = apply_act(__uint_as_float(raw[j + 1]) * rs + b.y, e.act) * cs.y + res[c + j + 1];
              v[j + 2] = apply_act(__uint_as_float(raw[j + 2]) * rs + b.z, e.act) * cs.z + res[c + j + 2];
              v[j + 3] = apply_act(__uint_as_float(raw[j + 3]) * rs + b.w, e.act) * cs.w + res[c + j + 3];
            }
            if (orow >= 0) {
              if (e.out_f32) {
                float* po = e.out_f32 + (size_t)orow * e.ldo + col0;
                stg256f(po, v);
                stg256f(po + 8, v + 8);
              }
              if (e.out_hi) store_pair16(e.out_hi, e.out_lo, (size_t)orow * e.ldh + col0, v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_rank(smem_u32(&tempty_bar[buf]), 0));
    }
  }
  // teardown: nobody may leave (or free tensor memory) while the pair can still touch its shared memory / TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair_512(tmem_base);
}

}  // namespace

template <int PBN>
static int launch_pair_t(const csam_gemm_args* a, const GemmEpi& e, cudaStream_t st, int tiles_m) {
  using Cfg = PairCfg<PBN>;
  const int tiles_n = a->N / PBN;
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  if (make_tmap_2d_f16(&ta_hi, a->a_hi, a->M, a->K, a->lda, PM, PBK)) return 1;
  if (make_tmap_2d_f16(&ta_lo, a->a_lo, a->M, a->K, a->lda, PM, PBK)) return 1;
  if (make_tmap_2d_f16(&tw_hi, a->w_hi, a->N, a->K, a->ldw, PBN / 2, PBK)) return 1;
  if (make_tmap_2d_f16(&tw_lo, a->w_lo, a->N, a->K, a->ldw, PBN / 2, PBK)) return 1;
  CSAM_DYN_SMEM(gemm_pair_kernel<PBN>, Cfg::SMEM_BYTES, "gemm_pair_kernel");
  const int pairs = min(tiles_m * tiles_n, num_sms() / 2);
  gemm_pair_kernel<PBN><<<2 * pairs, P_THREADS, Cfg::SMEM_BYTES, st>>>(ta_hi, ta_lo, tw_hi, tw_lo, e, a->K, tiles_m, tiles_n);
  return check_launch("gemm_pair_kernel");
}

int launch_gemm_pair(const csam_gemm_args* a, const GemmEpi& e, cudaStream_t st) {
  // CSAM_GEMM_PAIR: 0 = never, 1 = 256-wide pair tiles whenever the problem qualifies, 2 (default since the third session
  // of round 2) = 256 x 128 pair tiles whenever it qualifies: the granularity of the single-CTA kernel with half the
  // L2 -> shared-memory traffic per flop for the W operand.  Measured with the encoders on two streams under the power
  // cap (three alternating runs on one box): GEMM class 16.0 against 16.6 ms serialised, step 41.2 against 41.5 ms, end
  // to end 23.85 against 23.56 images/s, SM clock no lower.  3 = 256-wide where the wave model below favours them (the default
  // until the MMA-issue fix of round 2: with `elect.sync` role branches the single-CTA kernel issues its MMAs back to
  // back and the whole step measured 45.2 ms without pair tiles, 45.7 ms with the wave model, 46.2 ms with pair tiles
  // everywhere -- and CUDA-graph replay no longer pays the cluster-launch penalty).
  // impl == CSAM_GEMM_TC_PAIR in the arguments forces the 256-wide kernel (tests, A/B measurements).
  static const int mode = getenv("CSAM_GEMM_PAIR") ? atoi(getenv("CSAM_GEMM_PAIR")) : 2;
  const bool forced128 = a->impl == CSAM_GEMM_TC_PAIR128;
  const bool forced = a->impl == CSAM_GEMM_TC_PAIR || forced128;
  if (mode == 0 && !forced) return -1;
  // qualifies: hi/lo split operands, K-major W, standard row-per-lane epilogue, whole k-blocks and pair tiles in N
  const int sms = num_sms();
  const int tiles_m = (a->M + 2 * PM - 1) / (2 * PM);
  const int tiles_n = a->N / 256;
  const bool ok = a->a_lo && a->w_lo && !a->b_mn_major && a->epi == CSAM_EPI_STD && e.direct && (a->K % PBK) == 0 &&
                  (a->N % 256) == 0 && a->K >= 4 * PBK;
  if (!ok) {
    if (forced) return fail("%s", "csam_gemm: the CTA-pair kernel needs split operands, K-major W, N % 256 == 0, K % 64 == 0, "
                                  "K >= 256 and 32-byte aligned outputs");
    return -1;
  }
  if (forced128) return launch_pair_t<128>(a, e, st, tiles_m);
  if (!forced && mode == 2) {
    if ((long long)tiles_m * (a->N / 128) * 2 < sms) return -1;
    return launch_pair_t<128>(a, e, st, tiles_m);
  }
  if (!forced && mode == 3) {
    // Wave model (measured, round 2: the tensor pipe is ~70 % active inside a pair tile against ~55-60 % inside a
    // single-CTA 128x128 tile, but a pair tile is four of those, so the persistent grid of 74 pairs quantises much
    // coarser than 148 CTAs).  Time in units of one 128x128xK tile: pairs 2 per round, single CTAs 1 per round;
    // the pair kernel is taken only where it saves at least a round's worth (e.g. 4096 x 1024 x 4096: 64 pair tiles
    // = one round against two rounds of 256 single tiles).
    const long long t2 = (long long)tiles_m * tiles_n, t1 = (long long)((a->M + 127) / 128) * ((a->N + 127) / 128);
    const long long rounds_pair = (t2 + sms / 2 - 1) / (sms / 2);
    const long long est_pair = 2 * rounds_pair, est_single = (t1 + sms - 1) / sms;
    // ties: measured 5330 x 4096 x 1024 (5 pair rounds against 10 single rounds) 365 against 372 TFLOP/s for the
    // single-CTA kernel, 4096 x 1024 x 4096 (1 against 2) 377 against 357 for the pairs
    if (est_pair > est_single || (est_pair == est_single && rounds_pair > 1) || t2 * 4 < sms) return -1;
  }
  return launch_pair_t<256>(a, e, st, tiles_m);
}

}  // namespace csam
