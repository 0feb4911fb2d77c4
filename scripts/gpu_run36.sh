#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
for mb in 0 32 64 200; do echo "persist $mb MB"; CSAM_L2_PERSIST_MB=$mb timeout 300 python scripts/prof_i2t.py 256 2>&1 | tail -1; done
CSAM_L2_PERSIST_MB=64 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none $NB -k 'regex:dec_i2t_layer' -s 2 -c 1 python scripts/prof_i2t.py 256 2>&1 | grep -E "dram__|gpu__time"
python -c "
import torch
p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size, 'persist max', getattr(p,'persisting_l2_cache_max_size', None))"
