#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 12 gpurun_out/$name.log; }
export PYTHONPATH=$PWD CSAM_ATTN_IMPL=0
run stats python scripts/profile_step.py 1 --stats
run launches ncu --metrics gpu__time_duration.sum --clock-control none -s 740 -c 800 --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 2
run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 4 -o gpurun_out/prof_gemm_r01 python scripts/profile_step.py 1
run ncu_attn ncu --set full --clock-control none --import-source on -k regex:vit_attention_tc_kernel -s 30 -c 3 -o gpurun_out/prof_attn_r01 python scripts/profile_step.py 1
run ncu_post ncu --set full --clock-control none --import-source on -k regex:post_ -c 6 -o gpurun_out/prof_post_r01 python scripts/profile_step.py 1
ls -la gpurun_out
