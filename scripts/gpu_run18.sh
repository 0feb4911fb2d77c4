#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 14 gpurun_out/$name.log | cut -c1-400; }
export PYTHONPATH=$PWD
run tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run attn_micro python scripts/bench_attn.py
CSAM_ATTN_NQ=1 run attn_micro_nq1 python scripts/bench_attn.py
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
