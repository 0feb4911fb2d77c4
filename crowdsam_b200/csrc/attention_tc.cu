// attention_tc.cu — tcgen05 ViT attention (K-WATTN / K-GATTN).  Placeholder dispatch until the
// tensor-core kernel lands: fail loudly rather than silently computing on another path.
#include "common.cuh"
namespace csam {
int vit_attention_simt(const csam_attn_args* a, cudaStream_t st);
int vit_attention_tc(const csam_attn_args* a, cudaStream_t st) {
  (void)a; (void)st;
  return fail("%s", "csam_vit_attention: tcgen05 implementation not built yet (use impl=1)");
}
}  // namespace csam
