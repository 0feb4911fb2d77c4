#!/bin/bash
# Fused image->token decoder layer: kernel parity, model parity, bench, host-overhead probe.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 12 gpurun_out/$name.log | cut -c1-400; }
run tests_i2t python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused_i2t" --timeout 300 -x
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gemm-shapes
grep '^\[gemm\]' gpurun_out/bench.log | head -8
run overhead python scripts/cpu_overhead.py
