#!/bin/bash
export PYTHONPATH=$PWD
echo "A: setmaxnreg 56/104, 16-wide streaming stores"; timeout 300 python scripts/prof_i2t.py 256 2>&1 | tail -1; timeout 300 python scripts/prof_i2t.py 1024 2>&1 | tail -1
echo "B: no setmaxnreg"; CSAM_LIB_PATH=$PWD/crowdsam_b200/_C/alt/libcsam_sm100.so timeout 300 python scripts/prof_i2t.py 256 2>&1 | tail -1; CSAM_LIB_PATH=$PWD/crowdsam_b200/_C/alt/libcsam_sm100.so timeout 300 python scripts/prof_i2t.py 1024 2>&1 | tail -1
