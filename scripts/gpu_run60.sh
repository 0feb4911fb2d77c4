#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 5 gpurun_out/$name.log | cut -c1-900; }
run tests_attn python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --timeout 600
timeout 300 python scripts/bench_attn.py 2>&1 | head -3
run bench python bench.py --steps 5 --warmup 3
