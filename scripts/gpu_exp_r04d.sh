#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py tests/test_gpu_pipeline_injected.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/prof_rle_path.py 2>&1 | tail -12
timeout 300 python scripts/prof_e2e_segments.py 2>&1 | tail -7
