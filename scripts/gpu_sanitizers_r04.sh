#!/bin/bash
# compute-sanitizer passes over what changed in the third session of round 2: the RLE kernels (count / scan / fill /
# diff), the attention kernel with V as one fp16 (no V_lo load: expect_tx bytes changed), the in-kernel window rel-pos
# MMA, and the two-stream encoder launch order (memcheck over one whole set_image + decode).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
export CSAM_TEST_IMPLS=0 CSAM_TEST_ATTN_IMPLS=0 CSAM_GRAPHS=0
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='rle or vit_attention_relpos and 25-14 or vit_attention_plain_ragged and 333 or vit_attention_plain_groups'
for tool in racecheck synccheck memcheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x -k "$SEL" > gpurun_out/${tool}_r04.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${tool}_r04.log | tail -6
done
timeout 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_graphs.py -q -x -k "two_stream" > gpurun_out/memcheck_two_streams_r04.log 2>&1
echo "memcheck two streams exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_two_streams_r04.log | tail -4
for f in gpurun_out/racecheck_r04.log gpurun_out/synccheck_r04.log gpurun_out/memcheck_r04.log gpurun_out/memcheck_two_streams_r04.log; do
  (head -c 8000 $f; echo; echo "[...]"; tail -c 4000 $f) > $f.short; mv $f.short $f
done
