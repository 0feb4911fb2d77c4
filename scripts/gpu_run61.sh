#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 25 gpurun_out/$name.log | cut -c1-300; }
run tests_amg python -m pytest tests/test_gpu_model.py -q -m gpu -k "automatic" --timeout 600
