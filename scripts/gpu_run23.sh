#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for nq in 1 2; do
for dbg in 0 111; do
  echo "=== NQ=$nq DBG=$dbg"
  CSAM_ATTN_NQ=$nq CSAM_ATTN_DBG=$dbg timeout 300 python scripts/bench_attn.py dino 2>&1 | grep dino
done; done
echo "=== all shapes"
timeout 300 python scripts/bench_attn.py 2>&1 | tail -6
for pf in 0 2 4; do echo "=== gemm micro L2PF=$pf"; CSAM_GEMM_L2PF=$pf timeout 300 python scripts/bench_gemm.py 2>&1 | tail -10; done
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 14 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --timeout 900 -x
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gemm-shapes
