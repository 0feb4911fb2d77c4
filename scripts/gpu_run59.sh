#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 15 gpurun_out/$name.log | cut -c1-330; }
run tests_attn python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --timeout 600
timeout 600 python scripts/run_config.py vit_h 32 2>&1 | tail -2
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
