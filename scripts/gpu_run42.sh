#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 4 gpurun_out/$name.log | cut -c1-330; }
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run bench_dual python bench.py --steps 5 --warmup 3 --no-cpu-baseline
CSAM_DUAL_STREAM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "resident leg"
