#!/bin/bash
# compute-sanitizer passes over the tcgen05 / TMA / mbarrier kernels: memcheck over the WHOLE -m gpu
# suite, racecheck and synccheck over the GEMM (single-CTA and CTA-pair), attention and fused-decoder kernel tests.
# Summaries go to gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export CSAM_TEST_IMPLS=0 CSAM_TEST_ATTN_IMPLS=0 CSAM_GRAPHS=0
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='gemm_plain and 256-384-1024 or gemm_epilogue or pair_tiles and 4096-1024-1024 or vit_attention_relpos and 25-14 or vit_attention_plain_ragged and 333 or decoder_fused_i2t_layer and 3-False or decoder_fused_t2i and 3 or layernorm256 or epilogue_upscaling'
for tool in racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -x -k "$SEL" > gpurun_out/${tool}_r02.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${tool}_r02.log | tail -8
done
timeout 2400 $CS --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_model.py::test_config2_vit_l_grid64_vs_reference --deselect tests/test_gpu_model.py::test_full_scale_vit_l_grid32_against_oracle_run > gpurun_out/memcheck_r02.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02.log | tail -5
# keep the transfer small
for f in gpurun_out/racecheck_r02.log gpurun_out/synccheck_r02.log gpurun_out/memcheck_r02.log; do
  (head -c 20000 $f; echo; echo "[...]"; tail -c 6000 $f) > $f.short; mv $f.short $f
done
