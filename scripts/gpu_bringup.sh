#!/bin/bash
# One gpurun call = the whole bring-up ladder; every stage runs in its own process under `timeout` so a
# trap or hang in one kernel cannot take the others (or the box) down.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 15 gpurun_out/$name.log; }
export PYTHONPATH=$PWD
run simt_kernels env CSAM_TEST_IMPLS=1 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300
run debug_gemm python scripts/debug_gemm.py
run tc_gemm env CSAM_TEST_IMPLS=0 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k gemm --timeout 300
run model_simt env CSAM_GEMM_IMPL=1 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 500
run model_tc python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 500
