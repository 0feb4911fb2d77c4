// elementwise.cu — memory-bound helper kernels of the hot path (all fp32 math).
// Call sites replaced are listed per entry point in include/csam.h.
#include "common.cuh"

namespace csam {

__constant__ float c_mean[3] = {123.675f, 116.28f, 103.53f};   // build_sam.py:148
__constant__ float c_std[3] = {58.395f, 57.12f, 57.375f};      // build_sam.py:149

// normalised + zero-padded 1024x1024 virtual image (sam.py:163-173)
// HWC: img[(y*w + x)*3 + c] (what cv2 / PIL hand over), else planar CHW
template <bool HWC>
__device__ __forceinline__ float norm_px(const uint8_t* img, int h, int w, int c, int y, int x) {
  if (y >= h || x >= w) return 0.f;
  const uint8_t v = HWC ? img[((size_t)y * w + x) * 3 + c] : img[((size_t)c * h + y) * w + x];
  return ((float)v - c_mean[c]) / c_std[c];
}

template <bool HWC>
__global__ void patchify_kernel(const uint8_t* __restrict__ img, int h, int w, int patch, int n_side, int resize_to,
                                __half* __restrict__ out_hi, __half* __restrict__ out_lo, int kpad) {
  const int pr = blockIdx.x / n_side, pc = blockIdx.x % n_side;
  const int kreal = 3 * patch * patch;
  const float scale = resize_to ? (float)1024 / (float)resize_to : 1.f;   // ATen: input/output in float
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < kreal) {
      const int c = k / (patch * patch);
      const int py = (k / patch) % patch, px = k % patch;
      const int Y = pr * patch + py, X = pc * patch + px;
      if (!resize_to) {
        v = norm_px<HWC>(img, h, w, c, Y, X);
      } else {   // upsample_bilinear2d, align_corners=False (predictor.py:104)
        float sy = fmaxf(scale * (Y + 0.5f) - 0.5f, 0.f);
        float sx = fmaxf(scale * (X + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < 1023 ? 1 : 0), x1 = x0 + (x0 < 1023 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        v = hy * (hx * norm_px<HWC>(img, h, w, c, y0, x0) + lx * norm_px<HWC>(img, h, w, c, y0, x1)) +
            ly * (hx * norm_px<HWC>(img, h, w, c, y1, x0) + lx * norm_px<HWC>(img, h, w, c, y1, x1));
      }
    }
    store_pair(out_hi, out_lo, (size_t)blockIdx.x * kpad + k, v);
  }
}

// ---- LayerNorm: one warp per output row, row cached in registers ----------------------
constexpr int LN_MAX_V4 = 16;   // cols <= 32*4*16 = 2048

// NV4 = float4 slots per lane (cols <= 128*NV4): small rows keep few registers -> high occupancy
template <int NV4>
__global__ void __launch_bounds__(256) layernorm_kernel(csam_ln_args a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.rows_out) return;
  const int src = a.row_map ? a.row_map[row] : row;
  const int nv = a.cols >> 2;
  __half* ohi = static_cast<__half*>(a.out_hi);
  __half* olo = static_cast<__half*>(a.out_lo);
  __half* o2hi = static_cast<__half*>(a.out2_hi);
  __half* o2lo = static_cast<__half*>(a.out2_lo);
  float4 v[NV4];
  if (src < 0) {   // zero row (window padding happens after the norm)
    for (int i = lane; i < nv; i += 32) {
      const int c = i * 4;
      if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + (size_t)row * a.ldo + c) = make_float4(0, 0, 0, 0);
      if (ohi) { *reinterpret_cast<uint2*>(ohi + (size_t)row * a.ldh + c) = make_uint2(0, 0);
                 if (olo) *reinterpret_cast<uint2*>(olo + (size_t)row * a.ldh + c) = make_uint2(0, 0); }
    }
    return;
  }
  const float* x = a.x + (size_t)src * a.ldx;
  const float* ad = a.add ? a.add + (size_t)(a.add_mod > 0 ? src % a.add_mod : src) * a.ldadd : nullptr;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int i = lane + j * 32;
    if (i < nv) {
      float4 t = *reinterpret_cast<const float4*>(x + i * 4);
      if (ad) { float4 u = *reinterpret_cast<const float4*>(ad + i * 4); t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
      v[j] = t;
      s += (t.x + t.y) + (t.z + t.w);
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (a.normalize) {
    mean = warp_sum(s) / a.cols;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
      const int i = lane + j * 32;
      if (i < nv) {
        const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw = v[j].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    rstd = 1.0f / sqrtf(warp_sum(q) / a.cols + a.eps);
  }
  const float* pe = a.pe ? a.pe + (size_t)(a.pe_mod > 0 ? row % a.pe_mod : row) * a.ldpe : nullptr;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int i = lane + j * 32;
    if (i < nv) {
      const int c = i * 4;
      float y[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      if (a.normalize) {
        const float4 g = a.gamma ? *reinterpret_cast<const float4*>(a.gamma + c) : make_float4(1, 1, 1, 1);
        const float4 b = a.beta ? *reinterpret_cast<const float4*>(a.beta + c) : make_float4(0, 0, 0, 0);
        y[0] = (y[0] - mean) * rstd * g.x + b.x;
        y[1] = (y[1] - mean) * rstd * g.y + b.y;
        y[2] = (y[2] - mean) * rstd * g.z + b.z;
        y[3] = (y[3] - mean) * rstd * g.w + b.w;
      }
      if (a.act) {
#pragma unroll
        for (int t = 0; t < 4; ++t) y[t] = apply_act(y[t], a.act);
      }
      if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + (size_t)row * a.ldo + c) = make_float4(y[0], y[1], y[2], y[3]);
      if (ohi) store_pair4(ohi, olo, (size_t)row * a.ldh + c, y);
      if (o2hi) {
        const float4 p = *reinterpret_cast<const float4*>(pe + c);
        const float z[4] = {y[0] + p.x, y[1] + p.y, y[2] + p.z, y[3] + p.w};
        store_pair4(o2hi, o2lo, (size_t)row * a.ldh + c, z);
      }
    }
  }
}

// ---- neck 3x3 im2col -------------------------------------------------------------------
__global__ void im2col3x3_kernel(const __half* __restrict__ xh, const __half* __restrict__ xl, int side, int C,
                                 __half* __restrict__ oh, __half* __restrict__ ol) {
  const int chunks = C / 8;
  const size_t total = (size_t)side * side * 9 * chunks;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    const int tap = (int)((i / chunks) % 9);
    const int pix = (int)(i / ((size_t)chunks * 9));
    const int y = pix / side + tap / 3 - 1, x = pix % side + tap % 3 - 1;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (y >= 0 && y < side && x >= 0 && x < side) {
      const size_t s = ((size_t)y * side + x) * C + ch * 8;
      vh = *reinterpret_cast<const uint4*>(xh + s);
      if (xl) vl = *reinterpret_cast<const uint4*>(xl + s);
    }
    const size_t d = (size_t)pix * 9 * C + (size_t)tap * C + ch * 8;
    *reinterpret_cast<uint4*>(oh + d) = vh;
    if (ol) *reinterpret_cast<uint4*>(ol + d) = vl;
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[j][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = t[threadIdx.x][j];
  }
}

// ---- bilinear, align_corners=False (ATen upsample_bilinear2d) ---------------------------
__device__ __forceinline__ void bil_coord(int d, float scale, int in, int& i0, int& i1, float& l1) {
  float s = fmaxf(scale * (d + 0.5f) - 0.5f, 0.f);
  i0 = min((int)s, in - 1);
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - i0;
}
__global__ void bilinear_kernel(const float* __restrict__ in, int n, int hin, int win, float* __restrict__ out,
                                int hout, int wout, int chlast) {
  const float sh = (float)hin / hout, sw = (float)win / wout;
  const size_t total = (size_t)n * hout * wout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c, y, x;
    if (chlast) { c = (int)(i % n); x = (int)((i / n) % wout); y = (int)(i / ((size_t)n * wout)); }
    else { x = (int)(i % wout); y = (int)((i / wout) % hout); c = (int)(i / ((size_t)wout * hout)); }
    int y0, y1, x0, x1; float ly, lx;
    bil_coord(y, sh, hin, y0, y1, ly);
    bil_coord(x, sw, win, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    float v00, v01, v10, v11;
    if (chlast) {
      v00 = in[((size_t)y0 * win + x0) * n + c]; v01 = in[((size_t)y0 * win + x1) * n + c];
      v10 = in[((size_t)y1 * win + x0) * n + c]; v11 = in[((size_t)y1 * win + x1) * n + c];
    } else {
      const float* p = in + (size_t)c * hin * win;
      v00 = p[(size_t)y0 * win + x0]; v01 = p[(size_t)y0 * win + x1];
      v10 = p[(size_t)y1 * win + x0]; v11 = p[(size_t)y1 * win + x1];
    }
    out[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// ---- prompt tokens ----------------------------------------------------------------------
__global__ void prompt_tokens_kernel(const float* __restrict__ coords01, const int* __restrict__ labels, int P,
                                     const float* __restrict__ gauss, const float* __restrict__ tok5,
                                     const float* __restrict__ point_emb, const float* __restrict__ nap,
                                     float* __restrict__ tokens) {
  const int p = blockIdx.x, c = threadIdx.x;   // 256 threads
  float* t = tokens + (size_t)p * 7 * 256;
#pragma unroll
  for (int i = 0; i < 5; ++i) t[i * 256 + c] = tok5[i * 256 + c];
  const int lab = labels[p];
  float v;
  if (lab == -1) {
    v = nap[c];
  } else {
    const float x = 2.f * coords01[p * 2 + 0] - 1.f;
    const float y = 2.f * coords01[p * 2 + 1] - 1.f;
    const int f = c & 127;
    float ang = x * gauss[f] + y * gauss[128 + f];
    ang = 6.283185307179586f * ang;
    v = (c < 128) ? sinf(ang) : cosf(ang);
    v += point_emb[(lab == 1 ? 256 : 0) + c];
  }
  t[5 * 256 + c] = v;
  t[6 * 256 + c] = nap[c];   // padding point: PE zeroed then + not_a_point_embed
}

// ---- mask upscaling tail ------------------------------------------------------------------
// one warp per ConvT1 GEMM row (256 values = 4 positions x 64 channels); lane l owns 8 values
__global__ void __launch_bounds__(256) shuffle_ln_gelu_kernel(const float* __restrict__ y1, int rows,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, __half* __restrict__ ohi, __half* __restrict__ olo) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4 a = *reinterpret_cast<const float4*>(y1 + (size_t)row * 256 + lane * 8);
  const float4 b = *reinterpret_cast<const float4*>(y1 + (size_t)row * 256 + lane * 8 + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q += d * d; }
  q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
  const float rstd = 1.0f / sqrtf(q * (1.f / 64.f) + eps);
  const int pos = lane >> 3, c0 = (lane & 7) * 8;
  const int p = row >> 12, pix = row & 4095;
  const int y = pix >> 6, x = pix & 63;
  const size_t orow = (size_t)p * 16384 + (size_t)(2 * y + (pos >> 1)) * 128 + (2 * x + (pos & 1));
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = gelu_erf((v[j] - mean) * rstd * gamma[c0 + j] + beta[c0 + j]);
  store_pair8(ohi, olo, orow * 64 + c0, v);
}

// thread per (ConvT2 GEMM row, position): 32 channels (bias + GELU already applied) . hyper_in[p, l, :]
__global__ void __launch_bounds__(256) hyper_masks_kernel(const float* __restrict__ y2, int P,
                                                          const float* __restrict__ hyper, float* __restrict__ masks) {
  __shared__ float sh[4][32];
  const int p = blockIdx.y;
  if (threadIdx.x < 128) sh[threadIdx.x >> 5][threadIdx.x & 31] = hyper[(size_t)p * 128 + threadIdx.x];
  __syncthreads();
  const int idx = blockIdx.x * 256 + threadIdx.x;     // over 16384 * 4
  const int row = idx >> 2, pos = idx & 3;
  const float* src = y2 + ((size_t)p * 16384 + row) * 128 + pos * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 32; c += 4) {
    const float4 u = *reinterpret_cast<const float4*>(src + c);
#pragma unroll
    for (int l = 0; l < 4; ++l)
      acc[l] += u.x * sh[l][c] + u.y * sh[l][c + 1] + u.z * sh[l][c + 2] + u.w * sh[l][c + 3];
  }
  const int Y = 2 * (row >> 7) + (pos >> 1), X = 2 * (row & 127) + (pos & 1);
#pragma unroll
  for (int l = 0; l < 4; ++l) masks[(((size_t)p * 4 + l) * 256 + Y) * 256 + X] = acc[l];
}

// ---- PWD softmax weights --------------------------------------------------------------------
// RV = float4 slots per thread kept in registers (n == 4096 * RV: the row is read from memory ONCE); RV = 0 is the
// generic two-pass version.  exp via ex2.approx (relative error 2^-22), 8-byte packed hi / lo stores.
template <int RV>
__global__ void __launch_bounds__(1024) softmax_weights_kernel(const float* __restrict__ x, int n, __half* __restrict__ ehi,
                                                               __half* __restrict__ elo, float* __restrict__ inv_sum) {
  __shared__ float red[32];
  const float* row = x + (size_t)blockIdx.x * n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr float LOG2E = 1.4426950408889634f;
  float4 keep[RV > 0 ? RV : 1];
  float m = -INFINITY;
  if (RV > 0) {
#pragma unroll
    for (int k = 0; k < RV; ++k) {
      keep[k] = *reinterpret_cast<const float4*>(row + threadIdx.x * 4 + k * 4096);
      m = fmaxf(m, fmaxf(fmaxf(keep[k].x, keep[k].y), fmaxf(keep[k].z, keep[k].w)));
    }
  } else {
    for (int i = threadIdx.x * 4; i < n; i += 4096) {
      const float4 v = *reinterpret_cast<const float4*>(row + i);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  }
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = warp_max(red[lane]);
  __syncthreads();
  const float ml = m * LOG2E;
  float s = 0.f;
  auto emit = [&](const float4& v, size_t at) {
    float e[4];
    e[0] = exp2f(fmaf(v.x, LOG2E, -ml)); e[1] = exp2f(fmaf(v.y, LOG2E, -ml));
    e[2] = exp2f(fmaf(v.z, LOG2E, -ml)); e[3] = exp2f(fmaf(v.w, LOG2E, -ml));
    s += (e[0] + e[1]) + (e[2] + e[3]);
#pragma unroll
    for (int t = 0; t < 4; ++t) e[t] *= 16384.f;
    store_pair4(ehi, elo, at, e);
  };
  if (RV > 0) {
#pragma unroll
    for (int k = 0; k < RV; ++k) emit(keep[k], (size_t)blockIdx.x * n + threadIdx.x * 4 + k * 4096);
  } else {
    for (int i = threadIdx.x * 4; i < n; i += 4096) emit(*reinterpret_cast<const float4*>(row + i), (size_t)blockIdx.x * n + i);
  }
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  s = warp_sum(red[lane]);
  if (threadIdx.x == 0) inv_sum[blockIdx.x] = 1.0f / (s * 16384.f);
}

__global__ void select_kernel(const float* __restrict__ iou, const float* __restrict__ cls, int P, int ncls,
                              float* __restrict__ score, int* __restrict__ sel, int* __restrict__ cat) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float best = 0.f; int bi = 0;
  for (int l = 0; l < 4; ++l) {
    // model.py:351  clamp(iou,0) * sigmoid(cls.squeeze(2))   (n_class == 1 on this path)
    const float c = cls[((size_t)p * 4 + l) * ncls];
    const float s = fmaxf(iou[p * 4 + l], 0.f) * (1.0f / (1.0f + expf(-c)));
    if (l == 0 || s > best) { best = s; bi = l; }
  }
  score[p] = best;
  sel[p] = bi;
  int bc = 0; float bv = cls[((size_t)p * 4 + bi) * ncls];
  for (int k = 1; k < ncls; ++k) { const float v = cls[((size_t)p * 4 + bi) * ncls + k]; if (v > bv) { bv = v; bc = k; } }
  cat[p] = bc;
}

}  // namespace csam

using namespace csam;

extern "C" int csam_patchify(const uint8_t* img, int h, int w, int hwc, int patch, int n_side, int resize_to,
                             void* out_hi, void* out_lo, int kpad, void* stream) {
  CSAM_REQUIRE(img && out_hi && h > 0 && w > 0 && h <= 1024 && w <= 1024, "csam_patchify: bad image");
  CSAM_REQUIRE(kpad >= 3 * patch * patch, "csam_patchify: kpad too small");
  if (hwc)
    patchify_kernel<true><<<n_side * n_side, 256, 0, (cudaStream_t)stream>>>(img, h, w, patch, n_side, resize_to,
                                                                            (__half*)out_hi, (__half*)out_lo, kpad);
  else
    patchify_kernel<false><<<n_side * n_side, 256, 0, (cudaStream_t)stream>>>(img, h, w, patch, n_side, resize_to,
                                                                             (__half*)out_hi, (__half*)out_lo, kpad);
  return check_launch("patchify_kernel");
}

extern "C" int csam_layernorm(const csam_ln_args* a, void* stream) {
  CSAM_REQUIRE(a && a->x && a->rows_out > 0, "csam_layernorm: bad args");
  CSAM_REQUIRE((a->cols % 4) == 0 && a->cols <= 32 * 4 * LN_MAX_V4, "csam_layernorm: cols must be a multiple of 4 and <= 2048");
  CSAM_REQUIRE((a->ldx % 4) == 0 && (!a->out_f32 || (a->ldo % 4) == 0) && (!a->out_hi || (a->ldh % 4) == 0) &&
                   (!a->add || (a->ldadd % 4) == 0) && (!a->pe || (a->ldpe % 4) == 0),
               "csam_layernorm: strides must be multiples of 4");
  CSAM_REQUIRE(!a->out2_hi || a->pe, "csam_layernorm: out2 needs pe");
  const dim3 grid((a->rows_out + 7) / 8);
  if (a->cols <= 256) layernorm_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  else if (a->cols <= 1024) layernorm_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  else layernorm_kernel<LN_MAX_V4><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("layernorm_kernel");
}

extern "C" int csam_im2col3x3(const void* x_hi, const void* x_lo, int side, int C, void* out_hi, void* out_lo, void* stream) {
  CSAM_REQUIRE(x_hi && out_hi && (C % 8) == 0, "csam_im2col3x3: bad args");
  im2col3x3_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>((const __half*)x_hi, (const __half*)x_lo, side, C,
                                                            (__half*)out_hi, (__half*)out_lo);
  return check_launch("im2col3x3_kernel");
}

extern "C" int csam_transpose_f32(const float* in, int rows, int cols, float* out, void* stream) {
  CSAM_REQUIRE(in && out && rows > 0 && cols > 0, "csam_transpose_f32: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, rows, cols, out);
  return check_launch("transpose_kernel");
}

extern "C" int csam_bilinear(const float* in, int n, int hin, int win, float* out, int hout, int wout, int chlast,
                             void* stream) {
  CSAM_REQUIRE(in && out && n > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, "csam_bilinear: bad args");
  const size_t total = (size_t)n * hout * wout;
  const int blocks = (int)min((total + 255) / 256, (size_t)148 * 16);
  bilinear_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, n, hin, win, out, hout, wout, chlast);
  return check_launch("bilinear_kernel");
}

extern "C" int csam_prompt_tokens(const float* coords01, const int* labels, int P, const float* gauss,
                                  const float* out_tokens5, const float* point_emb, const float* not_a_point,
                                  float* tokens, void* stream) {
  CSAM_REQUIRE(coords01 && labels && gauss && out_tokens5 && point_emb && not_a_point && tokens && P > 0,
               "csam_prompt_tokens: bad args");
  prompt_tokens_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(coords01, labels, P, gauss, out_tokens5, point_emb,
                                                            not_a_point, tokens);
  return check_launch("prompt_tokens_kernel");
}

extern "C" int csam_upscale_shuffle_ln_gelu(const float* y1, int P, const float* gamma, const float* beta, float eps,
                                            void* out_hi, void* out_lo, void* stream) {
  CSAM_REQUIRE(y1 && gamma && beta && out_hi && P > 0, "csam_upscale_shuffle_ln_gelu: bad args");
  const int rows = P * 4096;
  shuffle_ln_gelu_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(y1, rows, gamma, beta, eps,
                                                                           (__half*)out_hi, (__half*)out_lo);
  return check_launch("shuffle_ln_gelu_kernel");
}

extern "C" int csam_upscale_hyper_masks(const float* y2, int P, const float* hyper_in, float* masks, void* stream) {
  CSAM_REQUIRE(y2 && hyper_in && masks && P > 0 && P <= 65535, "csam_upscale_hyper_masks: bad args");
  dim3 grid(16384 * 4 / 256, P);
  hyper_masks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y2, P, hyper_in, masks);
  return check_launch("hyper_masks_kernel");
}

extern "C" int csam_softmax_weights(const float* x, int R, int n, void* e_hi, void* e_lo, float* inv_sum, void* stream) {
  CSAM_REQUIRE(x && e_hi && inv_sum && R > 0 && n > 0 && (n % 4) == 0, "csam_softmax_weights: bad args");
  CSAM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(e_hi) & 7) == 0 &&
                   (!e_lo || (reinterpret_cast<uintptr_t>(e_lo) & 7) == 0),
               "csam_softmax_weights: alignment");
  // (keeping a 65,536-value row in registers would need the whole register file: the second pass re-reads the row,
  //  which is 256 KB and still in L2)
  if (n == 8192)
    softmax_weights_kernel<2><<<R, 1024, 0, (cudaStream_t)stream>>>(x, n, (__half*)e_hi, (__half*)e_lo, inv_sum);
  else
    softmax_weights_kernel<0><<<R, 1024, 0, (cudaStream_t)stream>>>(x, n, (__half*)e_hi, (__half*)e_lo, inv_sum);
  return check_launch("softmax_weights_kernel");
}

extern "C" int csam_select_candidates(const float* iou, const float* cls, int P, int ncls, float* score, int* sel,
                                      int* cat, void* stream) {
  CSAM_REQUIRE(iou && cls && score && sel && cat && P > 0 && ncls > 0, "csam_select_candidates: bad args");
  select_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(iou, cls, P, ncls, score, sel, cat);
  return check_launch("select_kernel");
}
