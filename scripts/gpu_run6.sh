#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 8 gpurun_out/$name.log; }
export PYTHONPATH=$PWD CSAM_ATTN_IMPL=0
run tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --timeout 800 -k "decoder or post or model or generate or set_image or non_square or layernorm"
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 2
