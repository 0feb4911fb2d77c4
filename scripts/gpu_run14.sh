#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 4 gpurun_out/$name.log | cut -c1-300; }
export PYTHONPATH=$PWD
run ncu_ln ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)256, \(int\)3, \(bool\)0, \(int\)1>' -c 2 -o gpurun_out/prof_gemm_ln_r01 python scripts/profile_step.py 1
run ncu_dec ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)128, \(int\)3, \(bool\)0, \(int\)0>' -s 200 -c 2 -o gpurun_out/prof_gemm_std_r01 python scripts/profile_step.py 1
