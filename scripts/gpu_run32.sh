#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run tests_i2t python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused_i2t" --timeout 300 -x
timeout 300 python scripts/prof_i2t.py 256
timeout 300 python scripts/prof_i2t.py 1024
NB="--kernel-name-base demangled"
timeout 600 ncu --set full --clock-control none --import-source on $NB -k 'regex:dec_i2t_layer' -s 2 -c 1 -f -o gpurun_out/prof_i2t_r01 python scripts/prof_i2t.py 256 > gpurun_out/ncu_i2t.log 2>&1; tail -2 gpurun_out/ncu_i2t.log
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
