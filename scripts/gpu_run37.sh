#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
for g in 148 96 64 32; do echo "grid $g"; CSAM_I2T_GRID=$g timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none $NB -k 'regex:dec_i2t_layer' -s 2 -c 1 python scripts/prof_i2t.py 256 2>&1 | grep -E "dram__|gpu__time"; done
