// gemm_shared.cuh -- epilogue parameter block shared by the K-GEMM kernels (gemm.cu, gemm_pair.cu).
#pragma once
#include "common.cuh"

namespace csam {

struct GemmEpi {
  int M, N;
  const float* bias; const float* row_scale; const float* col_scale; int act;
  const float* residual; int ldr; int res_mod;
  const int* row_map;
  float* out_f32; int ldo;
  __half* out_hi; __half* out_lo; int ldh;
  int vec_ok;   // all strides / bases allow 16-byte vector access
  int direct;   // N % 16 == 0 and all strides / bases allow 32-byte row-per-lane access
  int l2_prefetch;   // resident-weight mode: tiles of look-ahead for the A operand's L2 prefetch (0 = off)
  // fused epilogues
  const float* gamma; const float* beta; float eps;
  const float* pe; int ldpe; int pe_mod; __half* out2_hi; __half* out2_lo;
  const float* hyper; float* masks;
  const __half* res_hi; const __half* res_lo; int ldrh;   // EPI_LN: residual given as an h16 pair
};


// gemm_pair.cu: 2-CTA (cta_group::2) 256x256 pair-tile GEMM for the big encoder shapes; returns -1 when the
// problem does not qualify (the caller then takes the single-CTA kernel)
int launch_gemm_pair(const csam_gemm_args* a, const GemmEpi& e, cudaStream_t st);

}  // namespace csam
