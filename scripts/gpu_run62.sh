#!/bin/bash
# compute-sanitizer memcheck over the kernels added this round (small test shapes only)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 1000 \
  -k "fused_i2t_layer and True-3 or fused_t2i and True-3 or small_regions and 64-64 or relpos and 14-2-80 or head_dim_80 and 200 or two_query_tiles and 129" > gpurun_out/memcheck.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -12; tail -3 gpurun_out/memcheck.log | cut -c1-200
