#!/bin/bash
# Two MMA-issuing warps in attention + one-add descriptor stepping (attention, GEMM): parity, micro-benchmarks, bench.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 8 gpurun_out/$name.log | cut -c1-400; }
run tests_k python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -x
echo "=== NQ=1"; CSAM_ATTN_NQ=1 CSAM_ATTN_WIN_NQ=1 timeout 300 python scripts/bench_attn.py 2>&1 | tail -6
echo "=== NQ=2 (default)"; timeout 300 python scripts/bench_attn.py 2>&1 | tail -6
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gemm-shapes
grep '^\[gemm\]' gpurun_out/bench.log | head -14
