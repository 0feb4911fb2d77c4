#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 4 gpurun_out/$name.log | cut -c1-700; }
export PYTHONPATH=$PWD
run bench512 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --points-per-batch 512
run bench1024 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --points-per-batch 1024
run ncu_ln ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<256, 3, 0, 1>" -c 2 -o gpurun_out/prof_gemm_ln_r01 python scripts/profile_step.py 1
run ncu_dec ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<128, 3, 0, 0>" -s 200 -c 1 -o gpurun_out/prof_gemm_std_r01 python scripts/profile_step.py 1
