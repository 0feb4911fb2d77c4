// nms.cu — K-NMS (bit-exact torchvision.ops.nms), K-MIOU, EPS occupancy test, RLE encoding.
// Call sites replaced: model.py:171-176,257-263,429-434 (batched_nms with all-zero category ids,
// so the coordinate offset trick is a no-op), crowdsam/utils.py:422-479 (mask overlap, dead code),
// model.py:229-246 (occupancy), amg.py:107-135 (RLE).
#include "common.cuh"

namespace csam {

// Total order on scores that matches torch.sort(descending=True, stable=True), which torchvision's nms uses:
// NaN sorts as the greatest value (first), -0.0 == +0.0, everything else by value.  Mapping the float bits to
// an unsigned key makes the rank below a permutation for ANY input (with NaN scores the plain `>` / `==`
// comparisons are all false, several boxes would share a rank and slots of order[] would stay unwritten).
__device__ __forceinline__ uint32_t score_key(float s) {
  if (s != s) return 0xFFFFFFFFu;
  if (s == 0.f) s = 0.f;                      // -0.0 -> +0.0
  const uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// stable descending order by rank counting: rank_i = #{j : key_j > key_i or (key_j == key_i and j < i)}
__global__ void nms_rank_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, int n,
                                int* __restrict__ order, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ uint32_t sh[256];
  const uint32_t si = i < n ? score_key(scores[i]) : 0u;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    if (j0 + threadIdx.x < n) sh[threadIdx.x] = score_key(scores[j0 + threadIdx.x]);
    __syncthreads();
    const int lim = min(256, n - j0);
    for (int t = 0; t < lim; ++t) {
      const uint32_t sj = sh[t];
      rank += (sj > si) || (sj == si && (j0 + t) < i);
    }
  }
  if (i < n) {
    order[rank] = i;
    sorted[rank] = make_float4(boxes[i * 4 + 0], boxes[i * 4 + 1], boxes[i * 4 + 2], boxes[i * 4 + 3]);
  }
}

// torchvision arithmetic, no FMA contraction anywhere (explicit round-to-nearest ops)
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;   // NaN compares false
}

// mask[i][w] bit b set <=> sorted box j = 64 w + b (j > i) overlaps sorted box i above thr
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ sorted, int n, float thr,
                                                      unsigned long long* __restrict__ mask, int nw) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;   // only the upper triangle is ever read
  __shared__ float4 cbx[64];
  const int j = cb * 64 + threadIdx.x;
  if (j < n) cbx[threadIdx.x] = sorted[j];
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  const float4 bi = sorted[i];
  unsigned long long bits = 0;
  const int lim = min(64, n - cb * 64);
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  for (int t = start; t < lim; ++t)
    if (iou_gt(bi, cbx[t], thr)) bits |= 1ULL << t;
  mask[(size_t)i * nw + cb] = bits;
}

// greedy scan in sorted order (one block; thread w owns word w of the removed bitmap)
__global__ void nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n, int nw,
                                int* __restrict__ keep, int* __restrict__ n_keep) {
  extern __shared__ unsigned long long remv[];
  __shared__ int s_cnt;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) remv[w] = 0;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    const int wi = i >> 6, bi = i & 63;
    const bool alive = !((remv[wi] >> bi) & 1ULL);
    __syncthreads();
    if (alive) {
      if (threadIdx.x == 0) keep[s_cnt++] = order[i];
      for (int w = wi + threadIdx.x; w < nw; w += blockDim.x) remv[w] |= mask[(size_t)i * nw + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_keep = s_cnt;
}

// ---- K-MIOU ------------------------------------------------------------------------------------
constexpr int MO_SIDE = 150, MO_BITS = MO_SIDE * MO_SIDE, MO_WORDS = (MO_BITS + 31) / 32;   // 704

__global__ void mo_pack_kernel(const uint8_t* __restrict__ masks, int h, int w, uint32_t* __restrict__ packed) {
  const int i = blockIdx.y;
  const int word = blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= MO_WORDS) return;
  const float sh = (float)h / MO_SIDE, sw = (float)w / MO_SIDE;   // F.interpolate nearest (utils.py:431)
  uint32_t bits = 0;
  for (int b = 0; b < 32; ++b) {
    const int t = word * 32 + b;
    if (t < MO_BITS) {
      const int y = min((int)floorf((t / MO_SIDE) * sh), h - 1), x = min((int)floorf((t % MO_SIDE) * sw), w - 1);
      if (masks[((size_t)i * h + y) * w + x]) bits |= 1u << b;
    }
  }
  packed[(size_t)i * MO_WORDS + word] = bits;
}
__global__ void mo_inter_kernel(const uint32_t* __restrict__ packed, int n, int* __restrict__ inter, int* __restrict__ area) {
  const int i = blockIdx.y, j = blockIdx.x;   // one warp per pair
  const uint32_t* a = packed + (size_t)i * MO_WORDS;
  const uint32_t* b = packed + (size_t)j * MO_WORDS;
  int c = 0;
  for (int w = threadIdx.x; w < MO_WORDS; w += 32) c += __popc(a[w] & b[w]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (threadIdx.x == 0) {
    inter[(size_t)i * n + j] = c;
    if (i == j) area[i] = c;
  }
}

// ---- EPS occupancy ------------------------------------------------------------------------------
__global__ void points_occupied_kernel(const uint8_t* __restrict__ masks, int n_masks, int h, int w,
                                       const uint8_t* __restrict__ flag, const int* __restrict__ pts, int n_pts,
                                       uint8_t* __restrict__ occ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const int x = pts[i * 2], y = pts[i * 2 + 1];
  uint8_t o = 0;
  if (x >= 0 && x < w && y >= 0 && y < h)
    for (int m = 0; m < n_masks && !o; ++m)
      if (flag[m] && masks[((size_t)m * h + y) * w + x]) o = 1;
  occ[i] = o;
}

// ---- RLE (column-major) -----------------------------------------------------------------------
// Element t of the Fortran-order flattening is mask[t % h][t / h]; a "change" is a position t >= 1 whose value differs
// from position t - 1.  Work item = (column c, row segment s) with RLE_NSEG segments per column, numbered c * NSEG + s
// (= the order of the flattening), so one 1024 x 1024 mask is 8192 items spread over the whole GPU instead of one block
// walking it (the kept masks of an image are often few: one block per mask left 147 SMs idle for ~0.3 ms per pass).
// Threads of a block take adjacent columns of the same segment: every row read is a coalesced byte row.
//   count: changes per item            -> item_cnt[n][w * NSEG]
//   scan : exclusive scan per mask (in place) + n_runs[i] = changes + 1 + (mask starts with 1)
//   fill : change positions            -> pos[offsets[i] + first_one + 1 + k]
//   diff : run lengths = successive differences of {0, positions..., h * w}
constexpr int RLE_NSEG = 8;

__device__ __forceinline__ uint8_t rle_pred(const uint8_t* m, int h, int w, int c, int r0) {
  // the value in front of element (r0, c) in column-major order; the very first element is its own predecessor
  if (r0 > 0) return m[(size_t)(r0 - 1) * w + c];
  if (c > 0) return m[(size_t)(h - 1) * w + c - 1];
  return m[0];
}

template <bool FILL>
__global__ void __launch_bounds__(256) rle_items_kernel(const uint8_t* __restrict__ masks, int h, int w, int seg_rows,
                                                        int* __restrict__ item_cnt, const long long* __restrict__ offsets,
                                                        int* __restrict__ pos) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= w * RLE_NSEG) return;
  const int c = tid % w, sgm = tid / w;
  const uint8_t* m = masks + (size_t)blockIdx.y * h * w;
  int* ic = item_cnt + (size_t)blockIdx.y * w * RLE_NSEG;
  const int r0 = sgm * seg_rows, r1 = min(h, r0 + seg_rows);
  int cnt = 0;
  int* out = nullptr;
  if (FILL) out = pos + offsets[blockIdx.y] + (m[0] ? 1 : 0) + 1 + ic[c * RLE_NSEG + sgm];
  if (r0 < r1) {
    uint8_t prev = rle_pred(m, h, w, c, r0);
    const uint8_t* col = m + c;
    int r = r0;
    for (; r + 8 <= r1; r += 8) {               // eight independent loads in flight
      uint8_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = col[(size_t)(r + u) * w];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (v[u] != prev) {
          if (FILL) out[cnt] = c * h + r + u;
          ++cnt;
        }
        prev = v[u];
      }
    }
    for (; r < r1; ++r) {
      const uint8_t v = col[(size_t)r * w];
      if (v != prev) {
        if (FILL) out[cnt] = c * h + r;
        ++cnt;
      }
      prev = v;
    }
  }
  if (!FILL) ic[c * RLE_NSEG + sgm] = cnt;
}

// one block per mask: exclusive scan of its item counts in place, n_runs
__global__ void __launch_bounds__(1024) rle_scan_kernel(const uint8_t* __restrict__ masks, int h, int w,
                                                        int* __restrict__ item_cnt, int* __restrict__ n_runs) {
  __shared__ int s_warp[32];
  const int items = w * RLE_NSEG;
  int* ic = item_cnt + (size_t)blockIdx.x * items;
  const int chunk = (items + 1023) / 1024;
  const int i0 = min(threadIdx.x * chunk, items), i1 = min(i0 + chunk, items);
  int sum = 0;
  for (int i = i0; i < i1; ++i) sum += ic[i];
  // block-wide exclusive scan of the 1024 partial sums
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = s_warp[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += v;
    }
    s_warp[lane] = ws;
  }
  __syncthreads();
  int run = inc - sum + (wid > 0 ? s_warp[wid - 1] : 0);
  for (int i = i0; i < i1; ++i) {
    const int v = ic[i];
    ic[i] = run;
    run += v;
  }
  if (threadIdx.x == 1023) {
    const uint8_t first = masks[(size_t)blockIdx.x * h * w];
    n_runs[blockIdx.x] = s_warp[31] + 1 + (first ? 1 : 0);      // counts start with the number of zeros
  }
}

__global__ void __launch_bounds__(256) rle_diff_kernel(const uint8_t* __restrict__ masks, int h, int w,
                                                       const long long* __restrict__ offsets, const int* __restrict__ n_runs,
                                                       const int* __restrict__ pos, int* __restrict__ runs) {
  const int b = blockIdx.y;
  const int first_one = masks[(size_t)b * h * w] ? 1 : 0;
  const int changes = n_runs[b] - 1 - first_one;
  const long long off = offsets[b];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0 && first_one) runs[off] = 0;
  if (r > changes) return;
  const int start = r == 0 ? 0 : pos[off + first_one + r];           // slot first_one + 1 + k = position of change k
  const int end = r == changes ? h * w : pos[off + first_one + r + 1];
  runs[off + first_one + r] = end - start;
}

}  // namespace csam

using namespace csam;

extern "C" long long csam_box_nms_scratch_bytes(int n) {
  const long long nw = (n + 63) / 64;
  return (long long)n * 4 + (long long)n * 16 + (long long)n * nw * 8 + 256;
}

extern "C" int csam_box_nms(const float* boxes, const float* scores, int n, float thr, int* keep_out, int* n_keep,
                            void* scratch, long long scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CSAM_REQUIRE(keep_out && n_keep && n >= 0, "csam_box_nms: bad args");
  if (n == 0) {
    cudaMemsetAsync(n_keep, 0, sizeof(int), st);
    return 0;
  }
  CSAM_REQUIRE(boxes && scores && scratch && scratch_bytes >= csam_box_nms_scratch_bytes(n), "csam_box_nms: scratch too small");
  CSAM_REQUIRE(n <= 65536, "csam_box_nms: n <= 65536");
  const int nw = (n + 63) / 64;
  uint8_t* s = static_cast<uint8_t*>(scratch);
  float4* sorted = reinterpret_cast<float4*>(s);                                  // 16-byte aligned first
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(s + (size_t)n * 16);
  int* order = reinterpret_cast<int*>(s + (size_t)n * 16 + (size_t)n * nw * 8);
  nms_rank_kernel<<<(n + 255) / 256, 256, 0, st>>>(scores, boxes, n, order, sorted);
  if (check_launch("nms_rank_kernel")) return 1;
  nms_mask_kernel<<<dim3(nw, nw), 64, 0, st>>>(sorted, n, thr, mask, nw);
  if (check_launch("nms_mask_kernel")) return 1;
  const int threads = min(1024, ((nw + 31) / 32) * 32);
  nms_scan_kernel<<<1, threads, nw * sizeof(unsigned long long), st>>>(mask, order, n, nw, keep_out, n_keep);
  return check_launch("nms_scan_kernel");
}

extern "C" long long csam_mask_overlap_scratch_bytes(int n) { return (long long)n * MO_WORDS * 4; }

extern "C" int csam_mask_overlap(const uint8_t* masks, int n, int h, int w, int* inter, int* area, void* scratch,
                                 long long scratch_bytes, void* stream) {
  CSAM_REQUIRE(masks && inter && area && n > 0 && n <= 65535, "csam_mask_overlap: bad args");
  CSAM_REQUIRE(scratch && scratch_bytes >= csam_mask_overlap_scratch_bytes(n), "csam_mask_overlap: scratch too small");
  uint32_t* packed = static_cast<uint32_t*>(scratch);
  mo_pack_kernel<<<dim3((MO_WORDS + 127) / 128, n), 128, 0, (cudaStream_t)stream>>>(masks, h, w, packed);
  if (check_launch("mo_pack_kernel")) return 1;
  mo_inter_kernel<<<dim3(n, n), 32, 0, (cudaStream_t)stream>>>(packed, n, inter, area);
  return check_launch("mo_inter_kernel");
}

extern "C" int csam_points_occupied(const uint8_t* masks, int n_masks, int h, int w, const uint8_t* flag,
                                    const int* pts_xy, int n_pts, uint8_t* occ, void* stream) {
  CSAM_REQUIRE(occ && n_pts >= 0, "csam_points_occupied: bad args");
  if (n_pts == 0) return 0;
  CSAM_REQUIRE(pts_xy && (n_masks == 0 || (masks && flag)), "csam_points_occupied: bad args");
  points_occupied_kernel<<<(n_pts + 127) / 128, 128, 0, (cudaStream_t)stream>>>(masks, n_masks, h, w, flag, pts_xy, n_pts, occ);
  return check_launch("points_occupied_kernel");
}

extern "C" long long csam_rle_scratch_bytes(int n, int h, int w) {
  (void)h;
  return (long long)sizeof(int) * (long long)(n > 0 ? n : 0) * (w > 0 ? w : 0) * RLE_NSEG;
}

extern "C" int csam_rle_count(const uint8_t* masks, int n, int h, int w, int* n_runs, void* scratch,
                              long long scratch_bytes, void* stream) {
  CSAM_REQUIRE(masks && n_runs && scratch && n > 0 && h > 0 && w > 0, "csam_rle_count: bad args");
  CSAM_REQUIRE(n <= 65535 && (long long)h * w < (1ll << 31), "csam_rle_count: at most 65535 masks of less than 2^31 pixels");
  CSAM_REQUIRE(scratch_bytes >= csam_rle_scratch_bytes(n, h, w), "csam_rle_count: scratch too small");
  const int seg_rows = (h + RLE_NSEG - 1) / RLE_NSEG;
  dim3 grid((w * RLE_NSEG + 255) / 256, n);
  rle_items_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(masks, h, w, seg_rows, (int*)scratch, nullptr, nullptr);
  if (check_launch("rle_items_kernel(count)")) return 1;
  rle_scan_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(masks, h, w, (int*)scratch, n_runs);
  return check_launch("rle_scan_kernel");
}

extern "C" int csam_rle_fill(const uint8_t* masks, int n, int h, int w, const long long* offsets, const int* n_runs,
                             int max_runs, int* pos, int* runs, const void* scratch, void* stream) {
  CSAM_REQUIRE(masks && offsets && n_runs && pos && runs && scratch && n > 0 && h > 0 && w > 0 && max_runs > 0,
               "csam_rle_fill: bad args");
  CSAM_REQUIRE(n <= 65535, "csam_rle_fill: at most 65535 masks");
  const int seg_rows = (h + RLE_NSEG - 1) / RLE_NSEG;
  dim3 grid((w * RLE_NSEG + 255) / 256, n);
  rle_items_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(masks, h, w, seg_rows, (int*)scratch, offsets, pos);
  if (check_launch("rle_items_kernel(fill)")) return 1;
  dim3 grid2((max_runs + 255) / 256, n);
  rle_diff_kernel<<<grid2, 256, 0, (cudaStream_t)stream>>>(masks, h, w, offsets, n_runs, pos, runs);
  return check_launch("rle_diff_kernel");
}
