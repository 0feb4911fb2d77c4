// abi.cu — library-wide state and dispatch for libcsam_sm100.so
#include "common.cuh"

namespace csam {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};
int vit_attention_simt(const csam_attn_args* a, cudaStream_t st);
int vit_attention_tc(const csam_attn_args* a, cudaStream_t st);
}  // namespace csam

using namespace csam;

extern "C" const char* csam_last_error(void) { return g_err; }
extern "C" int csam_abi_version(void) { return CSAM_ABI_VERSION; }
extern "C" long long csam_launch_count(void) { return g_launches.load(); }

extern "C" int csam_vit_attention(const csam_attn_args* a, void* stream) {
  CSAM_REQUIRE(a && a->qkv_hi && a->out_hi, "csam_vit_attention: null operand");
  CSAM_REQUIRE(a->groups > 0 && a->tokens > 0 && a->heads > 0, "csam_vit_attention: empty problem");
  CSAM_REQUIRE((a->rel_h == nullptr) == (a->rel_w == nullptr), "csam_vit_attention: rel_h and rel_w go together");
  if (a->impl == 1) return vit_attention_simt(a, (cudaStream_t)stream);
  return vit_attention_tc(a, (cudaStream_t)stream);
}

// Host-side COCO run-length string encoder (pycocotools maskApi.c rleToString), batched over masks.  A crowd image
// yields hundreds of instances; encoding them one by one through numpy cost ~50 us of interpreter time per mask.
extern "C" int csam_coco_rle_strings(const int* counts, const long long* offsets, int n_masks, char* out, long long cap,
                                     long long* out_offsets) {
  CSAM_REQUIRE(counts && offsets && out && out_offsets && n_masks >= 0, "csam_coco_rle_strings: null argument");
  long long p = 0;
  for (int m = 0; m < n_masks; ++m) {
    out_offsets[m] = p;
    const int* c = counts + offsets[m];
    const long long n = offsets[m + 1] - offsets[m];
    if (p + 7 * n > cap) return fail("%s", "csam_coco_rle_strings: output buffer too small (7 bytes per run suffice)");
    for (long long i = 0; i < n; ++i) {
      long long x = c[i];
      if (i > 2) x -= c[i - 2];
      if ((unsigned long long)(x + 16) < 32ull) {       // one character: the common case (|delta| < 16)
        out[p++] = (char)((x & 0x1f) + 48);
        continue;
      }
      bool more = true;
      while (more) {
        char ch = (char)(x & 0x1f);
        x >>= 5;
        more = (ch & 0x10) ? x != -1 : x != 0;
        if (more) ch |= 0x20;
        out[p++] = (char)(ch + 48);
      }
    }
  }
  out_offsets[n_masks] = p;
  return 0;
}
