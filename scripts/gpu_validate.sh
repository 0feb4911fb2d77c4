#!/bin/bash
# Validation of HEAD on a B200 box: GPU tests, smoke, bench (both arms).  usage: gpurun -- bash scripts/gpu_validate.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 14 gpurun_out/$name.log | cut -c1-600; }
run tests python -m pytest tests -q -m gpu --timeout 900 -x
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 5 --warmup 3 --gemm-shapes
run benchref python bench.py --impl reference --steps 1 --warmup 0
