#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 12 gpurun_out/$name.log; }
export PYTHONPATH=$PWD
run attn_tc env CSAM_TEST_ATTN_IMPLS=0 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "vit_attention" --timeout 300
run model_attn_tc env CSAM_ATTN_IMPL=0 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 800
run bench_simt python bench.py --steps 3 --warmup 3 --no-cpu-baseline
run bench_tc env CSAM_ATTN_IMPL=0 python bench.py --steps 3 --warmup 3
