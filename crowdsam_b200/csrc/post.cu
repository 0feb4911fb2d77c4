// post.cu — K-POST: per-prompt mask upsample + threshold + stability counts + box (HBM-bound).
// Replaces sam.py:132-161 (two bilinear resizes), model.py:372-384, amg.py:156-176 (stability),
// amg.py:303-346 (boxes).  The reference materialises [P,4,H,W] fp32 twice (~95 MB / prompt);
// here the logits are re-evaluated on the fly from the selected 256x256 plane (256 KB, L2
// resident) and only the bool mask (1 B / pixel) of surviving prompts is written.
#include "common.cuh"
#include <limits.h>

namespace csam {

struct PostGeom {
  int in_h, in_w, out_h, out_w;
  int identity;          // stage-2 resize is the identity (always true on the CrowdSAM path)
  float s2h, s2w;        // in/out scales of stage 2 (ATen: float(in)/out)
};

// stage 1: 256 -> 1024, align_corners=False (scale 0.25)
__device__ __forceinline__ void coord256(int d, int& i0, int& i1, float& l1) {
  const float s = fmaxf(0.25f * (d + 0.5f) - 0.5f, 0.f);
  i0 = (int)s;
  i1 = i0 + (i0 < 255 ? 1 : 0);
  l1 = s - i0;
}
__device__ __forceinline__ float stage1(const float* __restrict__ L, int y, int x) {
  int y0, y1, x0, x1; float ly, lx;
  coord256(y, y0, y1, ly);
  coord256(x, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * L[y0 * 256 + x0] + lx * L[y0 * 256 + x1]) + ly * (hx * L[y1 * 256 + x0] + lx * L[y1 * 256 + x1]);
}
__device__ __forceinline__ float eval_logit(const float* __restrict__ L, const PostGeom& g, int Y, int X) {
  if (g.identity) return stage1(L, Y, X);
  // stage 2: crop to (in_h,in_w) then bilinear to (out_h,out_w)
  // source index as ATen computes it on the CPU (area_pixel_compute_source_index: scale * (dst + 0.5) - 0.5 with a
  // ROUNDED product): an FMA contraction here moves the source coordinate by an ulp of ~600, i.e. the interpolation
  // weight by ~6e-5, enough to flip pixels next to a threshold
  float sy = fmaxf(__fsub_rn(__fmul_rn(g.s2h, Y + 0.5f), 0.5f), 0.f), sx = fmaxf(__fsub_rn(__fmul_rn(g.s2w, X + 0.5f), 0.5f), 0.f);
  const int y0 = min((int)sy, g.in_h - 1), x0 = min((int)sx, g.in_w - 1);
  const int y1 = y0 + (y0 < g.in_h - 1 ? 1 : 0), x1 = x0 + (x0 < g.in_w - 1 ? 1 : 0);
  const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * stage1(L, y0, x0) + lx * stage1(L, y0, x1)) + ly * (hx * stage1(L, y1, x0) + lx * stage1(L, y1, x1));
}

__device__ __forceinline__ const float* plane_of(const csam_post_args& a, int p) {
  const int l = a.sel ? a.sel[p] : 0;
  return a.low + ((size_t)p * a.planes + l) * 65536;
}

constexpr int POST_ROWS = 8;   // output rows per block

// 16 consecutive output pixels (Y, X0..X0+15), X0 a multiple of 16.
// Identity stage 2 (the CrowdSAM path): the x4 bilinear has only 4 horizontal phases, so the 16 pixels
// need 6 low-res columns of 2 rows: 12 loads instead of 64, and the weights are compile-time constants
// (.625 .875 .125 .375 -- exactly what ATen's formula yields in fp32).  Same operation order as the
// reference: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11).
__device__ __forceinline__ void eval16(const float* __restrict__ L, const PostGeom& g, int Y, int X0, float* v) {
  if (g.identity) {
    int y0, y1; float ly;
    coord256(Y, y0, y1, ly);
    const float hy = 1.f - ly;
    const int k0 = X0 >> 2;
    float r0[6], r1[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int cc = min(max(k0 - 1 + c, 0), 255);
      r0[c] = L[y0 * 256 + cc];
      r1[c] = L[y1 * 256 + cc];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i0 = (j + 2) >> 2;                                 // relative low-res column of the left tap
      float lx = ((j & 3) == 0) ? 0.625f : ((j & 3) == 1) ? 0.875f : ((j & 3) == 2) ? 0.125f : 0.375f;
      if (X0 == 0 && j < 2) lx = 0.f;                              // source index clamped to 0 at the left border
      const float hx = 1.f - lx;
      v[j] = hy * (hx * r0[i0] + lx * r0[i0 + 1]) + ly * (hx * r1[i0] + lx * r1[i0 + 1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (X0 + j < g.out_w) ? eval_logit(L, g, Y, X0 + j) : 0.f;
  }
}

__global__ void post_init_kernel(int* counts, int* boxes, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  counts[p * 3 + 0] = counts[p * 3 + 1] = counts[p * 3 + 2] = 0;
  boxes[p * 4 + 0] = INT_MAX; boxes[p * 4 + 1] = INT_MAX; boxes[p * 4 + 2] = -1; boxes[p * 4 + 3] = -1;
}
__global__ void post_finalize_kernel(int* boxes, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  // amg.py:333-338: empty mask (right < left or bottom < top) -> [0,0,0,0]
  if (boxes[p * 4 + 2] < boxes[p * 4 + 0] || boxes[p * 4 + 3] < boxes[p * 4 + 1]) {
    boxes[p * 4 + 0] = boxes[p * 4 + 1] = boxes[p * 4 + 2] = boxes[p * 4 + 3] = 0;
  }
}

// ---- identity fast path: one thread = 4 output rows x 16 pixels -------------------------------------
// Output rows 4g-2 .. 4g+1 all interpolate between low-res rows g-1 and g, so the horizontal pass
// (hx*v0 + lx*v1 for both rows) is done once for 64 pixels: 12 loads, ~5 flops / pixel, and the operation
// order is exactly the reference's hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11).
struct Quad {
  float ha[16], hb[16];   // horizontal interpolation of low-res rows g-1 (clamped) and g (clamped)
};
struct Taps {
  float ra[6], rb[6];     // the 12 low-res values the 64 output pixels of a quad depend on
};
// Loads the taps and classifies the quad.  Every output pixel is a convex combination of the 12 taps (weights
// hx + lx = 1, hy + ly = 1, all in [0, 1]) evaluated with 7 rounded fp32 operations, so it lies within
// [mn, mx] widened by a few ulps.  Returns +1 when all 64 pixels are certainly > hi, -1 when they are certainly
// <= lo, 0 when the quad has to be evaluated.  The margin (1e-6 relative = 16 ulps, plus a denormal floor) makes
// the decision safe against the rounding of the full evaluation; non-finite taps (0 * inf = NaN at the clamped
// borders, NaN compares false) always take the full evaluation.  On instance masks almost every quad is far
// from the object boundary: the stats pass is otherwise bound by its ~500 ALU operations per quad, not by HBM.
__device__ __forceinline__ int quad_load(const float* __restrict__ L, int gI, int X0, Taps& t, float hi, float lo) {
  const int ya = max(gI - 1, 0), yb = min(gI, 255);
  const int k0 = X0 >> 2;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int cc = min(max(k0 - 1 + c, 0), 255);
    t.ra[c] = L[ya * 256 + cc];
    t.rb[c] = L[yb * 256 + cc];
  }
  float mn = t.ra[0], mx = t.ra[0], sum = 0.f;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    mn = fminf(mn, fminf(t.ra[c], t.rb[c]));
    mx = fmaxf(mx, fmaxf(t.ra[c], t.rb[c]));
    sum += t.ra[c] + t.rb[c];
  }
  if (!(fabsf(sum) < 3.0e38f)) return 0;                   // NaN / inf among the taps
  const float margin = 1e-6f * fmaxf(fabsf(mn), fabsf(mx)) + 1e-30f;
  if (mn - margin > hi) return 1;
  if (mx + margin < lo) return -1;
  return 0;
}
__device__ __forceinline__ void quad_h(const Taps& t, int X0, Quad& q) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i0 = (j + 2) >> 2;
    float lx = ((j & 3) == 0) ? 0.625f : ((j & 3) == 1) ? 0.875f : ((j & 3) == 2) ? 0.125f : 0.375f;
    if (X0 == 0 && j < 2) lx = 0.f;
    const float hx = 1.f - lx;
    q.ha[j] = hx * t.ra[i0] + lx * t.ra[i0 + 1];
    q.hb[j] = hx * t.rb[i0] + lx * t.rb[i0 + 1];
  }
}
// vertical weight of output row Y = 4g - 2 + jr
__device__ __forceinline__ float quad_ly(int Y, int jr) {
  if (Y < 2) return 0.f;                                  // source index clamped to 0 at the top border
  return jr == 0 ? 0.125f : jr == 1 ? 0.375f : jr == 2 ? 0.625f : 0.875f;
}



__global__ void __launch_bounds__(256) post_stats_quad_kernel(csam_post_args a, PostGeom g, int segs_per_row) {
  const int p = blockIdx.y;
  const float* L = plane_of(a, p);
  const int groups = (g.out_h + 2 + 3) / 4;
  const int total = groups * segs_per_row;
  const float t_hi = a.thr + a.off, t_lo = a.thr - a.off;
  int c_hi = 0, c_lo = 0, c_mid = 0;
  int xmin = INT_MAX, xmax = -1, ymin = INT_MAX, ymax = -1;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < total; s += gridDim.x * 256) {
    const int gI = s / segs_per_row, X0 = (s % segs_per_row) * 16;
    const int nx = min(16, g.out_w - X0);
    Taps tp;
    const int cls = quad_load(L, gI, X0, tp, fmaxf(t_hi, fmaxf(t_lo, a.thr)), fminf(t_hi, fminf(t_lo, a.thr)));
    if (cls < 0) continue;                                  // all three comparisons false for all 64 pixels
    if (cls > 0) {                                          // all three true: counts and box from the geometry
      const int ya0 = max(4 * gI - 2, 0), yb0 = min(4 * gI + 1, g.out_h - 1);
      const int n = (yb0 - ya0 + 1) * nx;
      c_hi += n; c_lo += n; c_mid += n;
      xmin = min(xmin, X0); xmax = max(xmax, X0 + nx - 1);
      ymin = min(ymin, ya0); ymax = max(ymax, yb0);
      continue;
    }
    Quad q;
    quad_h(tp, X0, q);
#pragma unroll
    for (int jr = 0; jr < 4; ++jr) {
      const int Y = 4 * gI - 2 + jr;
      if (Y < 0 || Y >= g.out_h) continue;
      const float ly = quad_ly(Y, jr), hy = 1.f - ly;
      unsigned bits = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = hy * q.ha[j] + ly * q.hb[j];
        if (j < nx) {
          c_hi += v > t_hi;
          c_lo += v > t_lo;
          bits |= (v > a.thr ? 1u : 0u) << j;
        }
      }
      if (bits) {
        c_mid += __popc(bits);
        xmin = min(xmin, X0 + __ffs(bits) - 1);
        xmax = max(xmax, X0 + 31 - __clz(bits));
        ymin = min(ymin, Y); ymax = max(ymax, Y);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_hi += __shfl_xor_sync(0xffffffffu, c_hi, o);
    c_lo += __shfl_xor_sync(0xffffffffu, c_lo, o);
    c_mid += __shfl_xor_sync(0xffffffffu, c_mid, o);
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  __shared__ int sred[8][7];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sred[wid][0] = c_hi; sred[wid][1] = c_lo; sred[wid][2] = c_mid;
    sred[wid][3] = xmin; sred[wid][4] = ymin; sred[wid][5] = xmax; sred[wid][6] = ymax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      c_hi += sred[w][0]; c_lo += sred[w][1]; c_mid += sred[w][2];
      xmin = min(xmin, sred[w][3]); ymin = min(ymin, sred[w][4]);
      xmax = max(xmax, sred[w][5]); ymax = max(ymax, sred[w][6]);
    }
    if (c_hi) atomicAdd(&a.counts[p * 3 + 0], c_hi);
    if (c_lo) atomicAdd(&a.counts[p * 3 + 1], c_lo);
    if (c_mid) {
      atomicAdd(&a.counts[p * 3 + 2], c_mid);
      atomicMin(&a.boxes[p * 4 + 0], xmin); atomicMin(&a.boxes[p * 4 + 1], ymin);
      atomicMax(&a.boxes[p * 4 + 2], xmax); atomicMax(&a.boxes[p * 4 + 3], ymax);
    }
  }
}

__global__ void __launch_bounds__(256) post_write_quad_kernel(csam_post_args a, PostGeom g, int segs_per_row) {
  const int i = blockIdx.y;
  const int p = a.keep ? a.keep[i] : i;
  if (p < 0) return;                       // slot without a survivor (keep lists compacted on the device carry -1 tails)
  const float* L = plane_of(a, p);
  const int groups = (g.out_h + 2 + 3) / 4;
  const int total = groups * segs_per_row;
  uint8_t* mo = a.masks ? a.masks + (size_t)i * g.out_h * g.out_w : nullptr;
  float* lo = a.logits ? a.logits + (size_t)i * g.out_h * g.out_w : nullptr;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < total; s += gridDim.x * 256) {
    const int gI = s / segs_per_row, X0 = (s % segs_per_row) * 16;
    Taps tp;
    const int cls = quad_load(L, gI, X0, tp, a.thr, a.thr);
    if (cls != 0 && !lo) {                                  // uniform quad: four constant 16-byte stores
      const uint32_t w = cls > 0 ? 0x01010101u : 0u;
#pragma unroll
      for (int jr = 0; jr < 4; ++jr) {
        const int Y = 4 * gI - 2 + jr;
        if (Y < 0 || Y >= g.out_h) continue;
        *reinterpret_cast<uint4*>(mo + (size_t)Y * g.out_w + X0) = make_uint4(w, w, w, w);
      }
      continue;
    }
    Quad q;
    quad_h(tp, X0, q);
#pragma unroll
    for (int jr = 0; jr < 4; ++jr) {
      const int Y = 4 * gI - 2 + jr;
      if (Y < 0 || Y >= g.out_h) continue;
      const float ly = quad_ly(Y, jr), hy = 1.f - ly;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = hy * q.ha[j] + ly * q.hb[j];
      const size_t o = (size_t)Y * g.out_w + X0;
      if (mo) {
        uint32_t w[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          w[qq] = (v[qq * 4] > a.thr ? 1u : 0u) | (v[qq * 4 + 1] > a.thr ? 0x100u : 0u) |
                  (v[qq * 4 + 2] > a.thr ? 0x10000u : 0u) | (v[qq * 4 + 3] > a.thr ? 0x1000000u : 0u);
        *reinterpret_cast<uint4*>(mo + o) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (lo) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          *reinterpret_cast<float4*>(lo + o + qq * 4) = make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) post_stats_kernel(csam_post_args a, PostGeom g, int segs_per_row) {
  const int p = blockIdx.y;
  const float* L = plane_of(a, p);
  const int y_begin = blockIdx.x * POST_ROWS;
  const int y_end = min(y_begin + POST_ROWS, g.out_h);
  const float t_hi = a.thr + a.off, t_lo = a.thr - a.off;
  int c_hi = 0, c_lo = 0, c_mid = 0;
  int xmin = INT_MAX, xmax = -1, ymin = INT_MAX, ymax = -1;
  const int n = (y_end - y_begin) * segs_per_row;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int Y = y_begin + i / segs_per_row, X0 = (i % segs_per_row) * 16;
    const int nx = min(16, g.out_w - X0);
    float v[16];
    eval16(L, g, Y, X0, v);
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nx) {
        c_hi += v[j] > t_hi;
        c_lo += v[j] > t_lo;
        bits |= (v[j] > a.thr ? 1u : 0u) << j;
      }
    }
    if (bits) {
      c_mid += __popc(bits);
      xmin = min(xmin, X0 + __ffs(bits) - 1);
      xmax = max(xmax, X0 + 31 - __clz(bits));
      ymin = min(ymin, Y); ymax = max(ymax, Y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_hi += __shfl_xor_sync(0xffffffffu, c_hi, o);
    c_lo += __shfl_xor_sync(0xffffffffu, c_lo, o);
    c_mid += __shfl_xor_sync(0xffffffffu, c_mid, o);
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (c_hi) atomicAdd(&a.counts[p * 3 + 0], c_hi);
    if (c_lo) atomicAdd(&a.counts[p * 3 + 1], c_lo);
    if (c_mid) {
      atomicAdd(&a.counts[p * 3 + 2], c_mid);
      atomicMin(&a.boxes[p * 4 + 0], xmin); atomicMin(&a.boxes[p * 4 + 1], ymin);
      atomicMax(&a.boxes[p * 4 + 2], xmax); atomicMax(&a.boxes[p * 4 + 3], ymax);
    }
  }
}

// one thread = 16 consecutive output pixels of one row -> one 16-byte store
__global__ void __launch_bounds__(256) post_write_kernel(csam_post_args a, PostGeom g, int segs_per_row) {
  const int i = blockIdx.y;
  const int p = a.keep ? a.keep[i] : i;
  if (p < 0) return;                       // slot without a survivor (keep lists compacted on the device carry -1 tails)
  const float* L = plane_of(a, p);
  const int total = g.out_h * segs_per_row;
  uint8_t* mo = a.masks ? a.masks + (size_t)i * g.out_h * g.out_w : nullptr;
  float* lo = a.logits ? a.logits + (size_t)i * g.out_h * g.out_w : nullptr;
  const bool vec = (g.out_w & 15) == 0;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < total; s += gridDim.x * 256) {
    const int Y = s / segs_per_row, X0 = (s % segs_per_row) * 16;
    const int nx = min(16, g.out_w - X0);
    float v[16];
    eval16(L, g, Y, X0, v);
    const size_t o = (size_t)Y * g.out_w + X0;
    if (mo) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[q] = (v[q * 4] > a.thr ? 1u : 0u) | (v[q * 4 + 1] > a.thr ? 0x100u : 0u) |
               (v[q * 4 + 2] > a.thr ? 0x10000u : 0u) | (v[q * 4 + 3] > a.thr ? 0x1000000u : 0u);
      if (vec) {
        *reinterpret_cast<uint4*>(mo + o) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        for (int j = 0; j < nx; ++j) mo[o + j] = (uint8_t)((w[j >> 2] >> ((j & 3) * 8)) & 1u);
      }
    }
    if (lo) {
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(lo + o + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
        for (int j = 0; j < nx; ++j) lo[o + j] = v[j];
      }
    }
  }
}

static int make_geom(const csam_post_args* a, PostGeom& g) {
  CSAM_REQUIRE(a && a->low && a->P > 0, "csam_mask_post: bad args");
  CSAM_REQUIRE(a->in_h > 0 && a->in_h <= 1024 && a->in_w > 0 && a->in_w <= 1024 && a->out_h > 0 && a->out_w > 0,
               "csam_mask_post: bad sizes");
  CSAM_REQUIRE(a->planes == 4 || (a->planes == 1 && !a->sel), "csam_mask_post: planes must be 4 (with sel) or 1");
  g.in_h = a->in_h; g.in_w = a->in_w; g.out_h = a->out_h; g.out_w = a->out_w;
  g.identity = (a->in_h == a->out_h && a->in_w == a->out_w);
  g.s2h = (float)a->in_h / (float)a->out_h;
  g.s2w = (float)a->in_w / (float)a->out_w;
  return 0;
}

}  // namespace csam

using namespace csam;

extern "C" int csam_mask_post_stats(const csam_post_args* a, void* stream) {
  PostGeom g;
  if (make_geom(a, g)) return 1;
  CSAM_REQUIRE(a->counts && a->boxes && a->P <= 65535, "csam_mask_post_stats: need counts/boxes, P <= 65535");
  cudaStream_t st = (cudaStream_t)stream;
  post_init_kernel<<<(a->P + 255) / 256, 256, 0, st>>>(a->counts, a->boxes, a->P);
  if (check_launch("post_init_kernel")) return 1;
  const int segs = (g.out_w + 15) / 16;
  if (g.identity) {
    const int total = ((g.out_h + 5) / 4) * segs;
    // blocks per mask: a block's fixed cost (dependent sel -> plane loads, block reduction, 7 global atomics) is
    // amortised over several quads per thread; 65 blocks (one quad per thread) left the pass latency-bound (0.32 ms at P = 1024 on instance masks; 9 blocks: 0.21 ms)
    static const int gx_env = getenv("CSAM_POST_GX") ? atoi(getenv("CSAM_POST_GX")) : 0;
    const int gx = gx_env > 0 ? gx_env : 9;
    post_stats_quad_kernel<<<dim3(min((total + 255) / 256, gx), a->P), 256, 0, st>>>(*a, g, segs);
    if (check_launch("post_stats_quad_kernel")) return 1;
  } else {
    dim3 grid((g.out_h + POST_ROWS - 1) / POST_ROWS, a->P);
    post_stats_kernel<<<grid, 256, 0, st>>>(*a, g, segs);
    if (check_launch("post_stats_kernel")) return 1;
  }
  post_finalize_kernel<<<(a->P + 255) / 256, 256, 0, st>>>(a->boxes, a->P);
  return check_launch("post_finalize_kernel");
}

extern "C" int csam_mask_post_write(const csam_post_args* a, void* stream) {
  PostGeom g;
  if (make_geom(a, g)) return 1;
  const int n = a->keep ? a->n_keep : a->P;
  if (n == 0) return 0;
  CSAM_REQUIRE((a->masks || a->logits) && n > 0 && n <= 65535, "csam_mask_post_write: need an output, n <= 65535");
  const int segs = (g.out_w + 15) / 16;
  const int total = g.out_h * segs;
  dim3 grid(min((total + 255) / 256, 256), n);
  if (g.identity && (g.out_w & 15) == 0) {
    const int tq = ((g.out_h + 5) / 4) * segs;
    post_write_quad_kernel<<<dim3(min((tq + 255) / 256, 65), n), 256, 0, (cudaStream_t)stream>>>(*a, g, segs);
    return check_launch("post_write_quad_kernel");
  }
  post_write_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a, g, segs);
  return check_launch("post_write_kernel");
}
