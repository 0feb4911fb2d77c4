#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 6 gpurun_out/$name.log | cut -c1-3000; }
export PYTHONPATH=$PWD
run tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --timeout 800 -k "post or generate or model"
run bench python bench.py --steps 5 --warmup 3
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 2
run ncu_post ncu --set full --clock-control none --import-source on -k regex:post_ -c 6 -o gpurun_out/prof_post_r01 python scripts/profile_step.py 1
