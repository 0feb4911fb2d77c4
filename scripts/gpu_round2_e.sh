#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python scripts/debug_determinism_pipeline.py vit_l 2>&1 | grep "image" | tail -12
timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool memcheck python scripts/debug_determinism_pipeline.py vit_l 2>&1 | grep -E "image |ERROR SUMMARY|Error" | tail -14
