#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 15 gpurun_out/$name.log | cut -c1-330; }
run tests_cfg0 python -m pytest tests/test_gpu_model.py -q -m gpu -k "config0" --timeout 900
