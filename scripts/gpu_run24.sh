#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
timeout 600 ncu --set full --clock-control none --import-source on $NB -k 'regex:vit_attention_tc_kernel' -s 2 -c 1 -f -o gpurun_out/prof_attn2_r01 python scripts/prof_micro.py attn > gpurun_out/ncu_attn2.log 2>&1; tail -2 gpurun_out/ncu_attn2.log
timeout 600 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_tc_kernel' -s 4 -c 2 -f -o gpurun_out/prof_gemm2_r01 python scripts/prof_micro.py gemm > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
