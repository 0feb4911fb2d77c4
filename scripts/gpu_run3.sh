#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 12 gpurun_out/$name.log; }
export PYTHONPATH=$PWD
run model_tc python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu --timeout 800
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 3 --warmup 3
run bench_ref python bench.py --impl reference --steps 1 --warmup 1
