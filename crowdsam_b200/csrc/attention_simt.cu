// attention_simt.cu — fp32 CUDA-core attention kernels.
//   * vit_attention_simt : flash-style ViT attention with decomposed rel-pos bias (validation /
//     bring-up implementation of K-WATTN / K-GATTN; the tcgen05 version lives in attention_tc.cu)
//   * decoder attentions (7 tokens on one side): inherently tiny, SIMT is the right tool.
// Reference call sites: image_encoder.py:224-240,325-361; dinov2/layers/attention.py:56-69;
// transformer.py:228-254.
#include "common.cuh"

namespace csam {

// rel[(g*heads+h)*tokens + q][0..S) = q . Rh[qh - kh + S-1],  [S..2S) = q . Rw[qw - kw + S-1]
// (image_encoder.py:292-361; UNSCALED q)
__global__ void relpos_kernel(const __half* __restrict__ qhi, const __half* __restrict__ qlo, int ld, int groups,
                              int tokens, int heads, int hd, const float* __restrict__ rel_h,
                              const float* __restrict__ rel_w, int S, float* __restrict__ out) {
  const size_t total = (size_t)groups * heads * tokens * 2 * S;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % (2 * S));
    const int q = (int)((i / (2 * S)) % tokens);
    const int h = (int)((i / ((size_t)2 * S * tokens)) % heads);
    const int g = (int)(i / ((size_t)2 * S * tokens * heads));
    const int qh = q / S, qw = q % S;
    const float* tab = (j < S) ? rel_h + (size_t)(qh - j + S - 1) * hd : rel_w + (size_t)(qw - (j - S) + S - 1) * hd;
    const size_t qoff = ((size_t)g * tokens + q) * ld + (size_t)h * hd;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(load_pair(qhi, qlo, qoff + d), tab[d], acc);
    out[i] = acc;
  }
}

template <int HD>
__global__ void __launch_bounds__(128) vit_attention_simt_kernel(csam_attn_args a, const float* __restrict__ rel) {
  constexpr int KT = 32;              // keys per tile
  constexpr int LDS = HD + 4;         // padded row (floats), keeps float4 alignment, conflict-free
  __shared__ __align__(16) float Ks[KT][LDS];
  __shared__ __align__(16) float Vs[KT][LDS];
  const int g = blockIdx.z, h = blockIdx.y;
  const int qi = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int q = blockIdx.x * 32 + qi;
  const bool qvalid = q < a.tokens;
  const __half* hi = static_cast<const __half*>(a.qkv_hi);
  const __half* lo = static_cast<const __half*>(a.qkv_lo);
  const int D = a.heads * HD;
  const size_t gbase = (size_t)g * a.tokens;
  float4 qv[HD / 4];
  {
    const size_t off = (gbase + (qvalid ? q : 0)) * a.ld_qkv + (size_t)h * HD;
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      qv[d].x = load_pair(hi, lo, off + d * 4 + 0) * a.scale;
      qv[d].y = load_pair(hi, lo, off + d * 4 + 1) * a.scale;
      qv[d].z = load_pair(hi, lo, off + d * 4 + 2) * a.scale;
      qv[d].w = load_pair(hi, lo, off + d * 4 + 3) * a.scale;
    }
  }
  const float* relq = rel ? rel + (((size_t)g * a.heads + h) * a.tokens + (qvalid ? q : 0)) * 2 * a.S : nullptr;
  float4 acc[HD / 4];
#pragma unroll
  for (int d = 0; d < HD / 4; ++d) acc[d] = make_float4(0, 0, 0, 0);
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < a.tokens; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (HD / 8); i += 128) {
      const int r = i / (HD / 8), c8 = (i % (HD / 8)) * 8;
      const int key = k0 + r;
      float kf[8], vf[8];
      if (key < a.tokens) {
        const size_t ko = (gbase + key) * a.ld_qkv + D + (size_t)h * HD + c8;
        const size_t vo = ko + D;
#pragma unroll
        for (int t = 0; t < 8; ++t) { kf[t] = load_pair(hi, lo, ko + t); vf[t] = load_pair(hi, lo, vo + t); }
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) { kf[t] = 0.f; vf[t] = 0.f; }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) { Ks[r][c8 + t] = kf[t]; Vs[r][c8 + t] = vf[t]; }
    }
    __syncthreads();
    float s[8];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + part;
      const int key = k0 + r;
      float d0 = 0.f;
#pragma unroll
      for (int d = 0; d < HD / 4; ++d) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[r][d * 4]);
        d0 = fmaf(qv[d].x, kk.x, d0); d0 = fmaf(qv[d].y, kk.y, d0);
        d0 = fmaf(qv[d].z, kk.z, d0); d0 = fmaf(qv[d].w, kk.w, d0);
      }
      if (key < a.tokens) {
        if (relq) d0 = (d0 + relq[key / a.S]) + relq[a.S + key % a.S];
      } else {
        d0 = -INFINITY;
      }
      s[j] = d0;
      tmax = fmaxf(tmax, d0);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float mnew = fmaxf(m, tmax);
    const float alpha = expf(m - mnew);     // first tile: exp(-inf) = 0
    l *= alpha;
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) { acc[d].x *= alpha; acc[d].y *= alpha; acc[d].z *= alpha; acc[d].w *= alpha; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + part;
      const float p = expf(s[j] - mnew);    // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int d = 0; d < HD / 4; ++d) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[r][d * 4]);
        acc[d].x = fmaf(p, vv.x, acc[d].x); acc[d].y = fmaf(p, vv.y, acc[d].y);
        acc[d].z = fmaf(p, vv.z, acc[d].z); acc[d].w = fmaf(p, vv.w, acc[d].w);
      }
    }
    m = mnew;
  }
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  const float inv = 1.0f / l;
#pragma unroll
  for (int d = 0; d < HD / 4; ++d) {
    float4 t = acc[d];
    t.x += __shfl_xor_sync(0xffffffffu, t.x, 1); t.y += __shfl_xor_sync(0xffffffffu, t.y, 1);
    t.z += __shfl_xor_sync(0xffffffffu, t.z, 1); t.w += __shfl_xor_sync(0xffffffffu, t.w, 1);
    t.x += __shfl_xor_sync(0xffffffffu, t.x, 2); t.y += __shfl_xor_sync(0xffffffffu, t.y, 2);
    t.z += __shfl_xor_sync(0xffffffffu, t.z, 2); t.w += __shfl_xor_sync(0xffffffffu, t.w, 2);
    acc[d] = t;
  }
  if (!qvalid) return;
  __half* ohi = static_cast<__half*>(a.out_hi);
  __half* olo = static_cast<__half*>(a.out_lo);
  const size_t oo = (gbase + q) * a.ld_out + (size_t)h * HD;
  // the four lanes of a quad each write a quarter of the head vector
#pragma unroll
  for (int d = 0; d < HD / 4; ++d) {
    if ((d & 3) == part) {
      store_pair(ohi, olo, oo + d * 4 + 0, acc[d].x * inv);
      store_pair(ohi, olo, oo + d * 4 + 1, acc[d].y * inv);
      store_pair(ohi, olo, oo + d * 4 + 2, acc[d].z * inv);
      store_pair(ohi, olo, oo + d * 4 + 3, acc[d].w * inv);
    }
  }
}

// ---- decoder: few keys (nk <= 8) ------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_few_keys_kernel(csam_dec_attn_args a) {
  extern __shared__ float sm[];
  const int C = a.heads * a.hd;
  float* ks = sm;
  float* vs = sm + a.nk * C;
  const int b = blockIdx.y;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * C;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * C;
  for (int i = threadIdx.x; i < a.nk * C; i += 128) { ks[i] = kb[i]; vs[i] = vb[i]; }
  __syncthreads();
  const int qi = blockIdx.x * 128 + threadIdx.x;
  if (qi >= a.nq) return;
  const float* q = a.q + ((size_t)(a.Bq == 1 ? 0 : b) * a.nq + qi) * C;
  const float scale = 1.0f / sqrtf((float)a.hd);
  const size_t oo = ((size_t)b * a.nq + qi) * C;
  for (int h = 0; h < a.heads; ++h) {
    float qv[32];
    for (int d = 0; d < a.hd; ++d) qv[d] = q[h * a.hd + d];
    float s[8];
    float m = -INFINITY;
    for (int j = 0; j < a.nk; ++j) {
      float acc = 0.f;
      for (int d = 0; d < a.hd; ++d) acc = fmaf(qv[d], ks[j * C + h * a.hd + d], acc);
      s[j] = acc * scale;
      m = fmaxf(m, s[j]);
    }
    float l = 0.f;
    for (int j = 0; j < a.nk; ++j) { s[j] = expf(s[j] - m); l += s[j]; }
    const float inv = 1.0f / l;
    for (int d = 0; d < a.hd; ++d) {
      float o = 0.f;
      for (int j = 0; j < a.nk; ++j) o = fmaf(s[j] * inv, vs[j * C + h * a.hd + d], o);
      if (a.out_f32) a.out_f32[oo + h * a.hd + d] = o;
      if (a.out_hi) store_pair(static_cast<__half*>(a.out_hi), static_cast<__half*>(a.out_lo), oo + h * a.hd + d, o);
    }
  }
}

// ---- decoder: few queries (nq <= 8), many keys ---------------------------------------------------
__global__ void __launch_bounds__(256) attn_few_queries_kernel(csam_dec_attn_args a) {
  extern __shared__ float sm[];
  const int C = a.heads * a.hd;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = a.hd, nq = a.nq, nk = a.nk;
  float* sc = sm;                       // [nq][nk]
  float* qs = sc + (size_t)nq * nk;     // [nq][hd]
  float* red = qs + nq * hd;            // [8 warps][nq] then partial outputs [16][nq][hd]
  const float* qb = a.q + (size_t)(a.Bq == 1 ? 0 : b) * nq * C + h * hd;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * nk * C + h * hd;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * nk * C + h * hd;
  const float scale = 1.0f / sqrtf((float)hd);
  for (int i = threadIdx.x; i < nq * hd; i += 256) qs[i] = qb[(size_t)(i / hd) * C + (i % hd)];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
  for (int key = threadIdx.x; key < nk; key += 256) {
    float kv[32];
    for (int d = 0; d < hd; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(kb + (size_t)key * C + d);
      kv[d] = t.x; kv[d + 1] = t.y; kv[d + 2] = t.z; kv[d + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nq) {
        float acc = 0.f;
        for (int d = 0; d < hd; ++d) acc = fmaf(qs[i * hd + d], kv[d], acc);
        acc *= scale;
        sc[(size_t)i * nk + key] = acc;
        mx[i] = fmaxf(mx[i], acc);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { mx[i] = warp_max(mx[i]); if (lane == 0) red[wid * 8 + i] = mx[i]; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = red[i];
    for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w * 8 + i]);
    mx[i] = t;
  }
  __syncthreads();
  float sum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = 0.f;
  for (int key = threadIdx.x; key < nk; key += 256) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nq) {
        const float e = expf(sc[(size_t)i * nk + key] - mx[i]);
        sc[(size_t)i * nk + key] = e;
        sum[i] += e;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sum[i] = warp_sum(sum[i]); if (lane == 0) red[wid * 8 + i] = sum[i]; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w * 8 + i];
    sum[i] = t;
  }
  __syncthreads();
  // out[i][d] = sum_key e[i][key] * v[key][d] / sum[i]; thread = (key group kg, 16-wide d slice)
  const int dgroups = (hd + 15) / 16;           // hd = 16 -> 1, hd = 32 -> 2
  const int d = threadIdx.x & 15, kg = threadIdx.x >> 4;   // 16 key groups
  for (int ds = 0; ds < dgroups; ++ds) {
    const int dd = ds * 16 + d;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (dd < hd) {
      for (int key = kg; key < nk; key += 16) {
        const float vv = vb[(size_t)key * C + dd];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < nq) acc[i] = fmaf(sc[(size_t)i * nk + key], vv, acc[i]);
      }
    }
    float* part = red;                          // [16][8][16]
#pragma unroll
    for (int i = 0; i < 8; ++i) part[(kg * 8 + i) * 16 + d] = acc[i];
    __syncthreads();
    if (threadIdx.x < 8 * 16) {
      const int i = threadIdx.x >> 4;
      if (i < nq && dd < hd) {
        float t = 0.f;
        for (int g2 = 0; g2 < 16; ++g2) t += part[(g2 * 8 + i) * 16 + d];
        t /= sum[i];
        const size_t oo = ((size_t)b * nq + i) * C + h * hd + dd;
        if (a.out_f32) a.out_f32[oo] = t;
        if (a.out_hi) store_pair(static_cast<__half*>(a.out_hi), static_cast<__half*>(a.out_lo), oo, t);
      }
    }
    __syncthreads();
  }
}

// fills a->scratch with the decomposed rel-pos terms [groups*heads*tokens, 2S] (shared by both attention paths)
int compute_relpos(const csam_attn_args* a, cudaStream_t st) {
  CSAM_REQUIRE(a->S * a->S == a->tokens, "csam_vit_attention: rel-pos needs tokens == S*S");
  const long long need = csam_vit_attention_scratch_bytes(a->groups, a->tokens, a->heads, a->hd, a->S);
  CSAM_REQUIRE(a->scratch && a->scratch_bytes >= need, "csam_vit_attention: scratch too small");
  relpos_kernel<<<148 * 8, 256, 0, st>>>(static_cast<const __half*>(a->qkv_hi), static_cast<const __half*>(a->qkv_lo),
                                         a->ld_qkv, a->groups, a->tokens, a->heads, a->hd, a->rel_h, a->rel_w, a->S,
                                         a->scratch);
  return check_launch("relpos_kernel");
}

int vit_attention_simt(const csam_attn_args* a, cudaStream_t st) {
  const float* rel = nullptr;
  if (a->rel_h) {
    if (compute_relpos(a, st)) return 1;
    rel = a->scratch;
  }
  dim3 grid((a->tokens + 31) / 32, a->heads, a->groups);
  if (a->hd == 64) vit_attention_simt_kernel<64><<<grid, 128, 0, st>>>(*a, rel);
  else if (a->hd == 80) vit_attention_simt_kernel<80><<<grid, 128, 0, st>>>(*a, rel);
  else if (a->hd == 32) vit_attention_simt_kernel<32><<<grid, 128, 0, st>>>(*a, rel);
  else return fail("%s", "csam_vit_attention: head dim must be 32, 64 or 80");
  return check_launch("vit_attention_simt_kernel");
}

}  // namespace csam

using namespace csam;

extern "C" long long csam_vit_attention_scratch_bytes(int groups, int tokens, int heads, int hd, int S) {
  (void)hd;
  return (long long)groups * heads * tokens * 2 * S * (long long)sizeof(float);
}

extern "C" int csam_attn_few_keys(const csam_dec_attn_args* a, void* stream) {
  CSAM_REQUIRE(a && a->q && a->k && a->v && (a->out_f32 || a->out_hi), "csam_attn_few_keys: bad args");
  CSAM_REQUIRE(a->nk >= 1 && a->nk <= 8 && a->hd <= 32 && a->B <= 65535, "csam_attn_few_keys: nk <= 8, hd <= 32");
  const int C = a->heads * a->hd;
  dim3 grid((a->nq + 127) / 128, a->B);
  attn_few_keys_kernel<<<grid, 128, 2 * a->nk * C * sizeof(float), (cudaStream_t)stream>>>(*a);
  return check_launch("attn_few_keys_kernel");
}

extern "C" int csam_attn_few_queries(const csam_dec_attn_args* a, void* stream) {
  CSAM_REQUIRE(a && a->q && a->k && a->v && (a->out_f32 || a->out_hi), "csam_attn_few_queries: bad args");
  CSAM_REQUIRE(a->nq >= 1 && a->nq <= 8 && a->hd <= 32 && (a->hd % 4) == 0 && a->B <= 65535,
               "csam_attn_few_queries: nq <= 8, hd <= 32");
  const size_t smem = ((size_t)a->nq * a->nk + a->nq * a->hd + 16 * 8 * 16) * sizeof(float);
  CSAM_REQUIRE(smem <= 200 * 1024, "csam_attn_few_queries: nk too large for shared memory");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_few_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  dim3 grid(a->heads, a->B);
  attn_few_queries_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("attn_few_queries_kernel");
}
