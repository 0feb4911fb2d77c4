#!/bin/bash
# Validation after making two-stream encoders and the single-fp16 V operand the defaults: GPU tests, smoke, bench,
# host profile of generate().  usage: gpurun -- bash scripts/gpu_validate_r04.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests -q -m gpu --timeout 900 -x
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 20 --warmup 3
cp gpurun_out/bench.log gpurun_out/bench_r04.json
echo "=== host profile"; timeout 600 python scripts/prof_e2e_host.py 2>&1 | head -60 | cut -c1-200
