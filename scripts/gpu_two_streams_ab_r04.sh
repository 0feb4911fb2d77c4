#!/bin/bash
# A/B on one box: encoders on two streams (CSAM_TWO_STREAMS) x P V with V as one fp16 (CSAM_ATTN_PSPLIT=-1).
# usage: gpurun -- bash scripts/gpu_two_streams_ab_r04.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python scripts/bench_overlap.py 2>&1 | tail -2
for ps in 0 -1; do
  echo "=== smoke PSPLIT=$ps"; CSAM_ATTN_PSPLIT=$ps timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
done
echo "=== model tests, two streams + PSPLIT=-1"
CSAM_ATTN_PSPLIT=-1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
for ts in 0 1; do for ps in 0 -1; do
  CSAM_TWO_STREAMS=$ts CSAM_ATTN_PSPLIT=$ps timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${ts}_${ps}.json 2> gpurun_out/ab_${ts}_${ps}.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_${ts}_${ps}.json').read().strip().splitlines()[-1])
    print('TWO_STREAMS=$ts PSPLIT=$ps', round(d['ms_per_step'], 2), 'ms', round(d['value'], 2), 'img/s  e2e', round(d['e2e']['value'], 2), 'clk', d['clocks']['sm_mhz'], {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items() if 'attention' in k})
except Exception as e:
    print('TWO_STREAMS=$ts PSPLIT=$ps unparsed', e); print(open('gpurun_out/ab_${ts}_${ps}.err').read()[-1500:])
PY
done; done
