#!/bin/bash
# Repeated A/B on one box: CTA-pair GEMM (less L2 -> shared-memory traffic per flop) against the single-CTA kernel under
# the power cap, with the encoders on two streams.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for rep in 1 2 3; do for v in "CSAM_GEMM_PAIR=0" "CSAM_GEMM_PAIR=1" "CSAM_GEMM_PAIR=2"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp.json 2> gpurun_out/exp.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/exp.json').read().strip().splitlines()[-1])
    print('$v', round(d['ms_per_step'], 2), 'ms', round(d['value'], 2), 'img/s  e2e', round(d['e2e']['value'], 2), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['kernel_ms_per_step']['gemm_tensor'], 2))
except Exception as e:
    print('$v unparsed', e); print(open('gpurun_out/exp.err').read()[-800:])
PY
done; done
