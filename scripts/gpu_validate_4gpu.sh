#!/bin/bash
# 4-GPU check of the bench under torchrun (NCCL all-gather of detections).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench4.log 2>&1
echo "exit $?"; grep -E "resident leg" gpurun_out/bench4.log | head -4; tail -1 gpurun_out/bench4.log | cut -c1-300
