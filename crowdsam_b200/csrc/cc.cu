// cc.cu — K-CC: remove_small_regions on the device.
//
// Reference: segment_anything_cs/utils/amg.py:267-291, called per mask from crowdsam/model.py:395-443 after a
// D2H copy of every mask: cv2.connectedComponentsWithStats(working, 8) with working = mask ("islands") or ~mask
// ("holes"), then
//   holes  : fill every background component smaller than area_thresh;
//   islands: drop every foreground component smaller than area_thresh; if ALL are smaller keep the largest
//            (np.argmax over OpenCV's labels: ties go to the component OpenCV numbers first);
//   the second return value says whether any component was smaller than the threshold.
// Here: 8-connected components by union-find on the pixel grid (atomicMin hooking, label = smallest pixel index of
// the component), warp-aggregated area counts, one decision pass.  Integer-exact against OpenCV, including the
// tie-break: OpenCV's 8-way labelling (Spaghetti / BBDT) scans 2x2 blocks in raster order and numbers components by
// first encounter, so the order key of a component is the smallest (y/2, x/2) block index over its pixels.
#include "common.cuh"

namespace csam {

struct CcScratch {
  int* label;              // [n, h*w]  root pixel index, -1 = not in the working set
  int* area;               // [n, h*w]  valid at root pixels
  int* keymin;             // [n, h*w]  valid at root pixels: first 2x2 block (raster order) touching the component
  unsigned long long* best;   // [n]  (area << 32) | (0xFFFFFFFF - key) of the largest component
  int* n_large;            // [n]
  int* any_small;          // [n]
};

// find with path halving: labels only ever decrease towards the root, so shortening x -> grandparent with an
// atomicMin is safe under concurrent unions and keeps the trees shallow (solid regions would otherwise build
// row-long chains)
__device__ __forceinline__ int cc_find(int* label, int x) {
  while (true) {
    const int p = label[x];
    if (p == x) return x;
    const int gp = label[p];
    if (gp != p) atomicMin(&label[x], gp);
    x = p;
  }
}
__device__ __forceinline__ void cc_union(int* label, int a, int b) {
  while (true) {
    a = cc_find(label, a);
    b = cc_find(label, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }      // hook the larger root under the smaller one
    const int old = atomicMin(&label[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ masks, int n, int hw, int holes, CcScratch s) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int p = (int)(i % hw);
    const bool fg = (masks[i] != 0) != (holes != 0);
    s.label[i] = fg ? p : -1;
    s.area[i] = 0;
    s.keymin[i] = 0x7FFFFFFF;
  }
  if (blockIdx.x == 0 && threadIdx.x < n) { s.best[threadIdx.x] = 0ull; s.n_large[threadIdx.x] = 0; s.any_small[threadIdx.x] = 0; }
}

__global__ void __launch_bounds__(256) cc_merge_kernel(int n, int h, int w, CcScratch s) {
  const int hw = h * w;
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int m = (int)(i / hw), p = (int)(i % hw);
    int* L = s.label + (size_t)m * hw;
    if (L[p] < 0) continue;
    const int y = p / w, x = p % w;
    // the four already-visited neighbours of an 8-neighbourhood
    if (x > 0 && L[p - 1] >= 0) cc_union(L, p, p - 1);
    if (y > 0) {
      if (L[p - w] >= 0) cc_union(L, p, p - w);
      if (x > 0 && L[p - w - 1] >= 0) cc_union(L, p, p - w - 1);
      if (x + 1 < w && L[p - w + 1] >= 0) cc_union(L, p, p - w + 1);
    }
  }
}

__global__ void __launch_bounds__(256) cc_count_kernel(int n, int h, int w, CcScratch s) {
  const int hw = h * w;
  const size_t total = (size_t)n * hw;
  const size_t total_r = (total + 31) / 32 * 32;      // whole warps stay in the loop (match_any needs the mask)
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total_r; i += (size_t)gridDim.x * 256) {
    long long key = -1;                               // (mask, root) or -1
    int root = -1, m = 0, p = 0;
    if (i < total) {
      m = (int)(i / hw); p = (int)(i % hw);
      int* L = s.label + (size_t)m * hw;
      if (L[p] >= 0) {
        root = cc_find(L, p);
        L[p] = root;                                   // flatten (roots keep pointing at themselves)
        key = ((long long)m << 32) | (unsigned)root;
      }
    }
    // warp-aggregated statistics: one atomic per distinct component per warp instead of one per pixel
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, key);
    if (root >= 0) {
      const int y = p / w, x = p % w;
      int bk = (y >> 1) * ((w + 1) >> 1) + (x >> 1);
      // min of the block keys and the count over the peer group, computed by its leader
      const int leader = __ffs(peers) - 1;
      const int lane = threadIdx.x & 31;
      int cnt = __popc(peers);
      unsigned rem = peers;
      int kmin = bk;
      while (rem) {                                    // peers of a group all run this loop with the same `rem`
        const int src = __ffs(rem) - 1;
        kmin = min(kmin, __shfl_sync(peers, bk, src));
        rem &= rem - 1;
      }
      if (lane == leader) {
        atomicAdd(&s.area[(size_t)m * hw + root], cnt);
        atomicMin(&s.keymin[(size_t)m * hw + root], kmin);
      }
    }
  }
}

__global__ void __launch_bounds__(256) cc_stats_kernel(int n, int hw, int thresh, CcScratch s) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int m = (int)(i / hw), p = (int)(i % hw);
    if (s.label[i] != p) continue;                     // roots only
    const int a = s.area[i];
    if (a < thresh) atomicOr(&s.any_small[m], 1); else atomicAdd(&s.n_large[m], 1);
    const unsigned long long v = ((unsigned long long)(unsigned)a << 32) | (0xFFFFFFFFu - (unsigned)s.keymin[i]);
    atomicMax(&s.best[m], v);
  }
}

__global__ void __launch_bounds__(256) cc_apply_kernel(uint8_t* masks, int n, int hw, int thresh, int holes, CcScratch s,
                                                       uint8_t* changed) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int m = (int)(i / hw);
    if (!s.any_small[m]) continue;                     // reference: no small region -> mask returned untouched
    const int root = s.label[i];
    const size_t ri = (size_t)m * hw + (root >= 0 ? root : 0);
    if (holes) {
      // isin(regions, [0] + small): the mask itself plus the small background components
      if (root >= 0 && s.area[ri] < thresh) masks[i] = 1;
    } else {
      bool keep = false;
      if (root >= 0) {
        if (s.n_large[m] > 0) keep = s.area[ri] >= thresh;
        else {
          const unsigned long long b = s.best[m];
          keep = (unsigned)s.area[ri] == (unsigned)(b >> 32) && (0xFFFFFFFFu - (unsigned)s.keymin[ri]) == (unsigned)(b & 0xFFFFFFFFu);
        }
      }
      masks[i] = keep ? 1 : 0;
    }
  }
  if (changed && blockIdx.x == 0 && threadIdx.x < n) changed[threadIdx.x] = s.any_small[threadIdx.x] ? 1 : 0;
}

constexpr int CC_CHUNK = 32;      // masks per launch group: bounds the scratch (12 B per pixel) to ~400 MB at 1024 x 1024

}  // namespace csam

using namespace csam;

extern "C" long long csam_small_regions_scratch_bytes(int n, int h, int w) {
  const long long c = n < CC_CHUNK ? n : CC_CHUNK;
  return c * (long long)h * w * 12 + c * 16 + 256;
}

extern "C" int csam_remove_small_regions(uint8_t* masks, int n, int h, int w, int area_thresh, int mode, uint8_t* changed,
                                         void* scratch, long long scratch_bytes, void* stream) {
  CSAM_REQUIRE(masks && scratch && n >= 0 && h > 0 && w > 0 && (mode == 0 || mode == 1), "csam_remove_small_regions: bad args");
  CSAM_REQUIRE((long long)h * w < (1ll << 31), "csam_remove_small_regions: mask too large");
  CSAM_REQUIRE(scratch_bytes >= csam_small_regions_scratch_bytes(n, h, w), "csam_remove_small_regions: scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int hw = h * w;
  for (int m0 = 0; m0 < n; m0 += CC_CHUNK) {
    const int c = (n - m0) < CC_CHUNK ? (n - m0) : CC_CHUNK;
    CcScratch s;
    char* base = static_cast<char*>(scratch);
    s.label = reinterpret_cast<int*>(base);
    s.area = s.label + (size_t)c * hw;
    s.keymin = s.area + (size_t)c * hw;
    char* tail = reinterpret_cast<char*>(s.keymin + (size_t)c * hw);
    tail += (8 - (reinterpret_cast<uintptr_t>(tail) & 7)) & 7;
    s.best = reinterpret_cast<unsigned long long*>(tail);
    s.n_large = reinterpret_cast<int*>(s.best + c);
    s.any_small = s.n_large + c;
    uint8_t* mk = masks + (size_t)m0 * hw;
    const size_t total = (size_t)c * hw;
    const int blocks = (int)((total + 255) / 256 < (size_t)148 * 32 ? (total + 255) / 256 : (size_t)148 * 32);
    const int holes = mode == 0 ? 1 : 0;
    cc_init_kernel<<<blocks, 256, 0, st>>>(mk, c, hw, holes, s);
    if (check_launch("cc_init_kernel")) return 1;
    cc_merge_kernel<<<blocks, 256, 0, st>>>(c, h, w, s);
    if (check_launch("cc_merge_kernel")) return 1;
    cc_count_kernel<<<blocks, 256, 0, st>>>(c, h, w, s);
    if (check_launch("cc_count_kernel")) return 1;
    cc_stats_kernel<<<blocks, 256, 0, st>>>(c, hw, area_thresh, s);
    if (check_launch("cc_stats_kernel")) return 1;
    cc_apply_kernel<<<blocks, 256, 0, st>>>(mk, c, hw, area_thresh, holes, s, changed ? changed + m0 : nullptr);
    if (check_launch("cc_apply_kernel")) return 1;
  }
  return 0;
}
