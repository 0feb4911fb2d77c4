#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for nq in 1 2; do
for dbg in 0 1 2 4 8 15 32 64 96 111; do
  echo "=== NQ=$nq DBG=$dbg"
  CSAM_ATTN_NQ=$nq CSAM_ATTN_DBG=$dbg timeout 300 python scripts/bench_attn.py dino 2>&1 | grep dino
done; done
