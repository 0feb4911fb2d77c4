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
