#!/bin/bash
# Round-end evidence: launch list of one step + ncu --set full captures of the named kernels.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 3 gpurun_out/$name.log | cut -c1-200; }
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 2
run ncu_gemm ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_tc_kernel<\(int\)128, \(int\)3, \(bool\)0, \(int\)0, \(bool\)0>' -s 30 -c 4 -o gpurun_out/prof_gemm_enc_r01 python scripts/profile_step.py 1
run ncu_attn ncu --set full --clock-control none --import-source on $NB -k 'regex:vit_attention_tc_kernel' -s 22 -c 4 -o gpurun_out/prof_attn_r01 python scripts/profile_step.py 1
run ncu_post ncu --set full --clock-control none --import-source on $NB -k 'regex:post_' -c 6 -o gpurun_out/prof_post_r01 python scripts/profile_step.py 1
run ncu_ln ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_tc_kernel<\(int\)256' -c 3 -o gpurun_out/prof_gemm_dec_r01 python scripts/profile_step.py 1
run ncu_postfull ncu --set full --clock-control none $NB -k 'regex:post_write_quad' -c 2 -o gpurun_out/prof_post_p1024_r01 python scripts/bench_post.py
