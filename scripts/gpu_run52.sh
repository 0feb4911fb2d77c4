#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 6 gpurun_out/$name.log | cut -c1-330; }
run tests_attn python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --timeout 600 -x
echo "tail split"; timeout 300 python scripts/bench_attn.py dino 2>&1 | tail -2
echo "no tail split"; CSAM_ATTN_TAIL=0 timeout 300 python scripts/bench_attn.py dino 2>&1 | tail -2
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
