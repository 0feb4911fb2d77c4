#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
timeout 300 python scripts/prof_t2i.py 296 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on $NB -k 'regex:dec_t2i_kernel' -s 2 -c 1 -f -o gpurun_out/prof_t2i_r01 python scripts/prof_t2i.py 296 > gpurun_out/ncu_t2i.log 2>&1; tail -1 gpurun_out/ncu_t2i.log
timeout 600 ncu --set full --clock-control none --import-source on $NB -k 'regex:dec_i2t_layer' -s 2 -c 1 -f -o gpurun_out/prof_i2t_r01 python scripts/prof_i2t.py 256 > gpurun_out/ncu_i2t.log 2>&1; tail -1 gpurun_out/ncu_i2t.log
