#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 25 gpurun_out/$name.log; }
export PYTHONPATH=$PWD
run post env CSAM_TEST_IMPLS=1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "mask_post" --timeout 300
run model_tc python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 800 -x
run model_simt env CSAM_GEMM_IMPL=1 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 800
