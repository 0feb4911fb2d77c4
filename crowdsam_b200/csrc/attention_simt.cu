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
  const int ldq = a.ldq ? a.ldq : C, ldk = a.ldk ? a.ldk : C, ldv = a.ldv ? a.ldv : C;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * ldk;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * ldv;
  for (int i = threadIdx.x; i < a.nk * C; i += 128) {
    ks[i] = kb[(size_t)(i / C) * ldk + (i % C)];
    vs[i] = vb[(size_t)(i / C) * ldv + (i % C)];
  }
  __syncthreads();
  const int qi = blockIdx.x * 128 + threadIdx.x;
  if (qi >= a.nq) return;
  const float* q = a.q + ((size_t)(a.Bq == 1 ? 0 : b) * a.nq + qi) * ldq;
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
  const int ldq = a.ldq ? a.ldq : C, ldk = a.ldk ? a.ldk : C, ldv = a.ldv ? a.ldv : C;
  const float* qb = a.q + (size_t)(a.Bq == 1 ? 0 : b) * nq * ldq + h * hd;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * nk * ldk + h * hd;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * nk * ldv + h * hd;
  const float scale = 1.0f / sqrtf((float)hd);
  for (int i = threadIdx.x; i < nq * hd; i += 256) qs[i] = qb[(size_t)(i / hd) * ldq + (i % hd)];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float mx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
  for (int key = threadIdx.x; key < nk; key += 256) {
    float kv[32];
    for (int d = 0; d < hd; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(kb + (size_t)key * ldk + d);
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
        const float vv = vb[(size_t)key * ldv + dd];
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

// Block = one query row qh of one (group, head): the S queries of that row share the S rel_h table rows
// Rh[qh - kh + S-1]; rel_w rows are indexed by qw - kw.  Everything is staged in shared memory once.
__global__ void __launch_bounds__(256) relpos_rows_kernel(const __half* __restrict__ qhi, const __half* __restrict__ qlo,
                                                          int ld, int tokens, int heads, int hd,
                                                          const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                                                          int S, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int ldh = hd + 1;
  float* qs = sm;                       // [S][hd+1]
  float* th = qs + S * ldh;             // [S][hd+1]      rows qh - kh + S-1, kh = 0..S-1
  float* tw = th + S * ldh;             // [2S-1][hd+1]
  const int qh = blockIdx.x, h = blockIdx.y, g = blockIdx.z;
  for (int i = threadIdx.x; i < S * hd; i += 256) {
    const int r = i / hd, d = i % hd;
    qs[r * ldh + d] = load_pair(qhi, qlo, ((size_t)g * tokens + qh * S + r) * ld + (size_t)h * hd + d);
    th[r * ldh + d] = rel_h[(size_t)(qh - r + S - 1) * hd + d];
  }
  for (int i = threadIdx.x; i < (2 * S - 1) * hd; i += 256) tw[(i / hd) * ldh + i % hd] = rel_w[i];
  __syncthreads();
  float* o = out + (((size_t)g * heads + h) * tokens + (size_t)qh * S) * 2 * S;
  for (int i = threadIdx.x; i < S * 2 * S; i += 256) {
    const int qw = i % S, j = i / S;          // lanes vary qw: distinct smem rows, conflict-free (odd stride)
    const float* q = qs + qw * ldh;
    const float* t = (j < S) ? th + j * ldh : tw + (qw - (j - S) + S - 1) * ldh;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(q[d], t[d], acc);
    o[(size_t)qw * 2 * S + j] = acc;
  }
}


// Global blocks (S = 64, head dim 64): one block per (query row qh, head), one thread per (query, quarter of the 128
// outputs) with its q row in REGISTERS and the table rows read as broadcast / conflict-free float4: the generic
// kernel above does two scalar shared loads per FMA (1.07 G loads per layer: 190 us, shared-memory bound); here it is
// one 16-byte load per 4 FMAs.
__global__ void __launch_bounds__(256) relpos_rows64_kernel(const __half* __restrict__ qhi, const __half* __restrict__ qlo,
                                                            int ld, int tokens, int heads,
                                                            const float* __restrict__ rel_h, const float* __restrict__ rel_w,
                                                            float* __restrict__ out) {
  constexpr int S = 64, HD = 64, LDT = HD + 4;
  extern __shared__ __align__(16) float sm64[];
  float* th = sm64;                     // [64][LDT]   row kh = rel_h[qh - kh + 63]
  float* tw = th + S * LDT;             // [127][LDT]
  const int qh = blockIdx.x, h = blockIdx.y, g = blockIdx.z, t = threadIdx.x;
  for (int i = t; i < S * (HD / 4); i += 256) {
    const int r = i >> 4, d4 = (i & 15) * 4;
    *reinterpret_cast<float4*>(th + r * LDT + d4) = *reinterpret_cast<const float4*>(rel_h + (size_t)(qh - r + S - 1) * HD + d4);
  }
  for (int i = t; i < (2 * S - 1) * (HD / 4); i += 256) {
    const int r = i >> 4, d4 = (i & 15) * 4;
    *reinterpret_cast<float4*>(tw + r * LDT + d4) = *reinterpret_cast<const float4*>(rel_w + (size_t)r * HD + d4);
  }
  const int qw = t & 63, part = t >> 6;
  float q[HD];
  const size_t qo = ((size_t)g * tokens + (size_t)qh * S + qw) * ld + (size_t)h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(qhi + qo + d);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(ah[k]); q[d + 2 * k] = f.x; q[d + 2 * k + 1] = f.y; }
    if (qlo) {
      const uint4 b = *reinterpret_cast<const uint4*>(qlo + qo + d);
      const __half2* bl = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(bl[k]); q[d + 2 * k] += f.x; q[d + 2 * k + 1] += f.y; }
    }
  }
  __syncthreads();
  float* o = out + (((size_t)g * heads + h) * tokens + (size_t)qh * S + qw) * 2 * S + part * 32;
#pragma unroll 1
  for (int j0 = 0; j0 < 32; j0 += 4) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (part & 1) * 32 + j0 + u;                       // kh (parts 0, 1) or kw (parts 2, 3)
      const float* row = (part < 2) ? th + j * LDT : tw + (qw - j + S - 1) * LDT;
      float a0 = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(row + d);
        a0 = fmaf(q[d], r4.x, a0); a0 = fmaf(q[d + 1], r4.y, a0);
        a0 = fmaf(q[d + 2], r4.z, a0); a0 = fmaf(q[d + 3], r4.w, a0);
      }
      acc[u] = a0;
    }
    *reinterpret_cast<float4*>(o + j0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// Window blocks (S = 14): one block per (group, head), one thread per query with its q row in registers and both
// tables in shared memory.  The row-per-block kernel above launches 5600 blocks of 392 outputs for 25 windows x 16
// heads (78 us per layer, as long as the window attention itself); this one is a 400-block launch.
__global__ void __launch_bounds__(224) relpos_win14_kernel(const __half* __restrict__ qhi, const __half* __restrict__ qlo,
                                                           int ld, int heads, const float* __restrict__ rel_h,
                                                           const float* __restrict__ rel_w, float* __restrict__ out) {
  constexpr int S = 14, HD = 64, T = 196, LDT = HD + 4;     // rows padded to 68 floats: 16-byte aligned, bank-shifted
  __shared__ __align__(16) float th[27 * LDT];
  __shared__ __align__(16) float tw[27 * LDT];
  const int h = blockIdx.x, g = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < 27 * HD; i += 224) {
    th[(i >> 6) * LDT + (i & 63)] = rel_h[i];
    tw[(i >> 6) * LDT + (i & 63)] = rel_w[i];
  }
  __syncthreads();
  if (t >= T) return;
  float q[HD];
  const size_t qo = ((size_t)g * T + t) * ld + (size_t)h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(qhi + qo + d);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(ah[k]); q[d + 2 * k] = f.x; q[d + 2 * k + 1] = f.y; }
    if (qlo) {
      const uint4 b = *reinterpret_cast<const uint4*>(qlo + qo + d);
      const __half2* bl = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(bl[k]); q[d + 2 * k] += f.x; q[d + 2 * k + 1] += f.y; }
    }
  }
  const int qh = t / S, qw = t % S;
  float* o = out + (((size_t)g * heads + h) * T + t) * 2 * S;
  // 4 outputs at a time -> one 16-byte store (rows are 112 B apart: scalar stores cost a sector each)
#pragma unroll 1
  for (int j0 = 0; j0 < 2 * S; j0 += 4) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      const float* row = (j < S) ? th + (qh - j + S - 1) * LDT : tw + (qw - (j - S) + S - 1) * LDT;
      float a0 = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(row + d);
        a0 = fmaf(q[d], r4.x, a0); a0 = fmaf(q[d + 1], r4.y, a0);
        a0 = fmaf(q[d + 2], r4.z, a0); a0 = fmaf(q[d + 3], r4.w, a0);
      }
      acc[u] = a0;
    }
    *reinterpret_cast<float4*>(o + j0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// fills a->scratch with the decomposed rel-pos terms [groups*heads*tokens, 2S] (shared by both attention paths)
int compute_relpos(const csam_attn_args* a, cudaStream_t st) {
  CSAM_REQUIRE(a->S * a->S == a->tokens, "csam_vit_attention: rel-pos needs tokens == S*S");
  const long long need = csam_vit_attention_scratch_bytes(a->groups, a->tokens, a->heads, a->hd, a->S);
  CSAM_REQUIRE(a->scratch && a->scratch_bytes >= need, "csam_vit_attention: scratch too small");
  if (a->S == 14 && a->hd == 64 && a->tokens == 196 && (a->ld_qkv & 7) == 0 && a->heads <= 65535 &&
      !(getenv("CSAM_RELPOS_WIN") && atoi(getenv("CSAM_RELPOS_WIN")) == 0)) {
    relpos_win14_kernel<<<dim3(a->heads, a->groups), 224, 0, st>>>(
        static_cast<const __half*>(a->qkv_hi), static_cast<const __half*>(a->qkv_lo), a->ld_qkv, a->heads, a->rel_h,
        a->rel_w, a->scratch);
    return check_launch("relpos_win14_kernel");
  }
  if (a->S == 64 && a->hd == 64 && a->tokens == 4096 && (a->ld_qkv & 7) == 0 && a->groups <= 65535 &&
      !(getenv("CSAM_RELPOS_ROWS64") && atoi(getenv("CSAM_RELPOS_ROWS64")) == 0)) {
    constexpr int SM64 = (64 + 127) * 68 * 4;
    CSAM_DYN_SMEM(relpos_rows64_kernel, SM64, "relpos_rows64_kernel");
    relpos_rows64_kernel<<<dim3(64, a->heads, a->groups), 256, SM64, st>>>(
        static_cast<const __half*>(a->qkv_hi), static_cast<const __half*>(a->qkv_lo), a->ld_qkv, a->tokens, a->heads,
        a->rel_h, a->rel_w, a->scratch);
    return check_launch("relpos_rows64_kernel");
  }
  const size_t smem = (size_t)(4 * a->S - 1) * (a->hd + 1) * sizeof(float);
  CSAM_REQUIRE(smem <= 100 * 1024 && a->groups <= 65535, "csam_vit_attention: rel-pos table too large");
  CSAM_DYN_SMEM(relpos_rows_kernel, 100 * 1024, "relpos_rows_kernel");
  relpos_rows_kernel<<<dim3(a->S, a->heads, a->groups), 256, smem, st>>>(
      static_cast<const __half*>(a->qkv_hi), static_cast<const __half*>(a->qkv_lo), a->ld_qkv, a->tokens, a->heads,
      a->hd, a->rel_h, a->rel_w, a->S, a->scratch);
  return check_launch("relpos_rows_kernel");
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

// ---- decoder cross attentions, C = 128 = 8 heads x 16: coalesced warp-cooperative fast paths ------
// A warp reads one 512-byte row with one float4 per lane; lane l owns dims [4l,4l+4) of head l/4, so a
// head's dot product is a 4-lane butterfly.

// image -> token: every image token (row) attends to NK <= 8 tokens of its prompt.
template <int NK>
__global__ void __launch_bounds__(256) attn_i2t_kernel(csam_dec_attn_args a, int rows_per_block) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ldq = a.ldq ? a.ldq : 128, ldk = a.ldk ? a.ldk : 128, ldv = a.ldv ? a.ldv : 128;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * NK * ldk + lane * 4;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * NK * ldv + lane * 4;
  float4 kr[NK], vr[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    kr[j] = *reinterpret_cast<const float4*>(kb + j * ldk);
    vr[j] = *reinterpret_cast<const float4*>(vb + j * ldv);
  }
  const float* qb = a.q + (size_t)(a.Bq == 1 ? 0 : b) * a.nq * ldq;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, a.nq);
  __half* ohi = static_cast<__half*>(a.out_hi);
  __half* olo = static_cast<__half*>(a.out_lo);
  for (int r = r0 + warp; r < r1; r += 8) {
    const float4 q = *reinterpret_cast<const float4*>(qb + (size_t)r * ldq + lane * 4);
    float s[NK];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      float d = q.x * kr[j].x + q.y * kr[j].y + q.z * kr[j].z + q.w * kr[j].w;
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      s[j] = d * 0.25f;                       // 1/sqrt(16)
      m = fmaxf(m, s[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NK; ++j) { s[j] = expf(s[j] - m); l += s[j]; }
    const float inv = 1.0f / l;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      const float p = s[j] * inv;
      o[0] = fmaf(p, vr[j].x, o[0]); o[1] = fmaf(p, vr[j].y, o[1]);
      o[2] = fmaf(p, vr[j].z, o[2]); o[3] = fmaf(p, vr[j].w, o[3]);
    }
    const size_t oo = ((size_t)b * a.nq + r) * 128 + lane * 4;
    if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + oo) = make_float4(o[0], o[1], o[2], o[3]);
    if (ohi) store_pair4(ohi, olo, oo, o);
  }
}

// token -> image: NQ <= 8 tokens attend to all nk image tokens of their prompt; one block per prompt,
// each warp streams a strided set of 4-key tiles with an online softmax, partials merged in smem.
template <int NQ>
__global__ void __launch_bounds__(512) attn_t2i_kernel(csam_dec_attn_args a) {
  extern __shared__ float sm[];
  constexpr int NW = 16;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ldq = a.ldq ? a.ldq : 128, ldk = a.ldk ? a.ldk : 128, ldv = a.ldv ? a.ldv : 128;
  const float* qb = a.q + (size_t)(a.Bq == 1 ? 0 : b) * NQ * ldq + lane * 4;
  const float* kb = a.k + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * ldk + lane * 4;
  const float* vb = a.v + (size_t)(a.Bk == 1 ? 0 : b) * a.nk * ldv + lane * 4;
  float4 qr[NQ];
  float m[NQ], l[NQ];
  float4 acc[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    qr[i] = *reinterpret_cast<const float4*>(qb + i * ldq);
    qr[i].x *= 0.25f; qr[i].y *= 0.25f; qr[i].z *= 0.25f; qr[i].w *= 0.25f;
    m[i] = -INFINITY; l[i] = 0.f; acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int k0 = warp * 4; k0 < a.nk; k0 += NW * 4) {
    float4 kk[4], vv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int key = min(k0 + t, a.nk - 1);
      kk[t] = *reinterpret_cast<const float4*>(kb + (size_t)key * ldk);
      vv[t] = *reinterpret_cast<const float4*>(vb + (size_t)key * ldv);
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float s[4];
      float tmax = -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float d = qr[i].x * kk[t].x + qr[i].y * kk[t].y + qr[i].z * kk[t].z + qr[i].w * kk[t].w;
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        s[t] = (k0 + t < a.nk) ? d : -INFINITY;
        tmax = fmaxf(tmax, s[t]);
      }
      const float mn = fmaxf(m[i], tmax);
      const float al = expf(m[i] - mn);
      float4 ac = acc[i];
      ac.x *= al; ac.y *= al; ac.z *= al; ac.w *= al;
      float li = l[i] * al;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float p = expf(s[t] - mn);
        li += p;
        ac.x = fmaf(p, vv[t].x, ac.x); ac.y = fmaf(p, vv[t].y, ac.y);
        ac.z = fmaf(p, vv[t].z, ac.z); ac.w = fmaf(p, vv[t].w, ac.w);
      }
      acc[i] = ac; l[i] = li; m[i] = mn;
    }
  }
  // merge the NW partial states: sm_m/sm_l [NW][NQ][8 heads], sm_acc [NW][NQ][128]
  float* sm_m = sm;
  float* sm_l = sm_m + NW * NQ * 8;
  float* sm_acc = sm_l + NW * NQ * 8;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    if ((lane & 3) == 0) { sm_m[(warp * NQ + i) * 8 + (lane >> 2)] = m[i]; sm_l[(warp * NQ + i) * 8 + (lane >> 2)] = l[i]; }
    *reinterpret_cast<float4*>(sm_acc + (size_t)(warp * NQ + i) * 128 + lane * 4) = acc[i];
  }
  __syncthreads();
  __half* ohi = static_cast<__half*>(a.out_hi);
  __half* olo = static_cast<__half*>(a.out_lo);
  for (int t = threadIdx.x; t < NQ * 128; t += 512) {
    const int i = t >> 7, d = t & 127, hgrp = d >> 4;
    float M = -INFINITY;
    for (int w = 0; w < NW; ++w) M = fmaxf(M, sm_m[(w * NQ + i) * 8 + hgrp]);
    float num = 0.f, den = 0.f;
    for (int w = 0; w < NW; ++w) {
      const float e = expf(sm_m[(w * NQ + i) * 8 + hgrp] - M);   // warps that saw no key: exp(-inf) = 0
      num = fmaf(sm_acc[(size_t)(w * NQ + i) * 128 + d], e, num);
      den = fmaf(sm_l[(w * NQ + i) * 8 + hgrp], e, den);
    }
    const float o = num / den;
    const size_t oo = ((size_t)b * NQ + i) * 128 + d;
    if (a.out_f32) a.out_f32[oo] = o;
    if (ohi) store_pair(ohi, olo, oo, o);
  }
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
  CSAM_REQUIRE(((a->ldq | a->ldk | a->ldv) & 3) == 0 && a->ldq >= 0 && a->ldk >= 0 && a->ldv >= 0,
               "csam_attn_few_keys: row strides must be multiples of 4 floats");
  const int C = a->heads * a->hd;
  if (a->heads == 8 && a->hd == 16 && a->nk == 7) {   // image -> token cross attention
    const int rows_per_block = 256;
    dim3 g((a->nq + rows_per_block - 1) / rows_per_block, a->B);
    attn_i2t_kernel<7><<<g, 256, 0, (cudaStream_t)stream>>>(*a, rows_per_block);
    return check_launch("attn_i2t_kernel");
  }
  dim3 grid((a->nq + 127) / 128, a->B);
  attn_few_keys_kernel<<<grid, 128, 2 * a->nk * C * sizeof(float), (cudaStream_t)stream>>>(*a);
  return check_launch("attn_few_keys_kernel");
}

extern "C" int csam_attn_few_queries(const csam_dec_attn_args* a, void* stream) {
  CSAM_REQUIRE(a && a->q && a->k && a->v && (a->out_f32 || a->out_hi), "csam_attn_few_queries: bad args");
  CSAM_REQUIRE(a->nq >= 1 && a->nq <= 8 && a->hd <= 32 && (a->hd % 4) == 0 && a->B <= 65535,
               "csam_attn_few_queries: nq <= 8, hd <= 32");
  CSAM_REQUIRE(((a->ldq | a->ldk | a->ldv) & 3) == 0 && a->ldq >= 0 && a->ldk >= 0 && a->ldv >= 0,
               "csam_attn_few_queries: row strides must be multiples of 4 floats");
  if (a->heads == 8 && a->hd == 16 && a->nq == 7) {   // token -> image cross attention
    const size_t sm = (size_t)16 * 7 * (8 + 8 + 128) * sizeof(float);
    CSAM_DYN_SMEM(attn_t2i_kernel<7>, (int)sm, "attn_t2i_kernel<7>");
    attn_t2i_kernel<7><<<a->B, 512, sm, (cudaStream_t)stream>>>(*a);
    return check_launch("attn_t2i_kernel");
  }
  const size_t smem = ((size_t)a->nq * a->nk + a->nq * a->hd + 16 * 8 * 16) * sizeof(float);
  CSAM_REQUIRE(smem <= 200 * 1024, "csam_attn_few_queries: nk too large for shared memory");
  CSAM_DYN_SMEM(attn_few_queries_kernel, 200 * 1024, "attn_few_queries_kernel");
  dim3 grid(a->heads, a->B);
  attn_few_queries_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(*a);
  return check_launch("attn_few_queries_kernel");
}
