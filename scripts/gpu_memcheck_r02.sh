#!/bin/bash
# compute-sanitizer memcheck over the -m gpu suite, one process per test file (the sanitizer's own memory overhead plus
# every cached full-size model of a single-process run exhausts the device: cudaMalloc failures, not access errors).
mkdir -p gpurun_out
export PYTHONPATH=$PWD CSAM_TEST_IMPLS=0 CSAM_TEST_ATTN_IMPLS=0
CS="/usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 10"
: > gpurun_out/memcheck_r02.log
for f in test_gpu_kernels.py test_gpu_pipeline_injected.py test_gpu_properties.py test_gpu_graphs.py test_gpu_dropin.py; do
  timeout 1500 $CS python -m pytest tests/$f -q > gpurun_out/mc_$f.log 2>&1
  echo "== $f rc=$? $(grep -E 'passed|failed' gpurun_out/mc_$f.log | tail -1) $(grep -E 'ERROR SUMMARY' gpurun_out/mc_$f.log | tail -1)" | tee -a gpurun_out/memcheck_r02.log
done
timeout 1500 $CS python -m pytest tests/test_gpu_model.py -q -k "not config2 and not full_scale and not config3 and not config1" > gpurun_out/mc_model.log 2>&1
echo "== test_gpu_model.py (configs[0], tiny, AMG, errors) rc=$? $(grep -E 'passed|failed' gpurun_out/mc_model.log | tail -1) $(grep -E 'ERROR SUMMARY' gpurun_out/mc_model.log | tail -1)" | tee -a gpurun_out/memcheck_r02.log
timeout 1500 $CS python -m pytest tests/test_gpu_model.py -q -k "config1" > gpurun_out/mc_model1.log 2>&1
echo "== test_gpu_model.py (configs[1] ViT-L headline) rc=$? $(grep -E 'passed|failed' gpurun_out/mc_model1.log | tail -1) $(grep -E 'ERROR SUMMARY' gpurun_out/mc_model1.log | tail -1)" | tee -a gpurun_out/memcheck_r02.log
grep -h "Invalid\|out of bounds\|misaligned" gpurun_out/mc_*.log | head -5
rm -f gpurun_out/mc_*.log
