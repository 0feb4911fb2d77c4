#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for nq in 1 2; do
for dbg in 0 15 96 111; do
  echo "=== NQ=$nq DBG=$dbg"
  CSAM_ATTN_NQ=$nq CSAM_ATTN_DBG=$dbg timeout 300 python scripts/bench_attn.py dino 2>&1 | grep dino
done; done
echo "=== all shapes"
timeout 300 python scripts/bench_attn.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or gemm" 2>&1 | tail -3
