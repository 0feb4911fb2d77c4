#!/bin/bash
# compute-sanitizer passes over the kernels whose synchronisation changed in the second session of round 2
# (elect.sync role branches everywhere; K-I2T / K-T2I protocols; the two-MMA k-step of the GEMM).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export CSAM_TEST_IMPLS=0 CSAM_TEST_ATTN_IMPLS=0 CSAM_GRAPHS=0
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='gemm_plain and 256-384-1024 or gemm_epilogue or vit_attention_relpos and 25-14 or vit_attention_plain_ragged and 333 or decoder_fused_i2t_layer and 3-False or decoder_fused_t2i and 3 or layernorm256 or epilogue_upscaling'
for tool in racecheck synccheck; do
  timeout 1200 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -q -x -k "$SEL" > gpurun_out/${tool}_r03.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${tool}_r03.log | tail -8
done
timeout 1500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/memcheck_r03.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r03.log | tail -5
for f in gpurun_out/racecheck_r03.log gpurun_out/synccheck_r03.log gpurun_out/memcheck_r03.log; do
  (head -c 20000 $f; echo; echo "[...]"; tail -c 6000 $f) > $f.short; mv $f.short $f
done
