#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python scripts/prof_e2e_segments.py 2>&1 | tail -16
for v in "CSAM_SIDE_PRIO=0 CSAM_SIDE_FIRST=1" "CSAM_SIDE_PRIO=-1 CSAM_SIDE_FIRST=1" "CSAM_SIDE_PRIO=0 CSAM_SIDE_FIRST=0"; do
  env $v timeout 300 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/exp.json 2> gpurun_out/exp.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/exp.json').read().strip().splitlines()[-1])
    print('$v', round(d['ms_per_step'], 2), 'ms', round(d['value'], 2), 'img/s  e2e', round(d['e2e']['value'], 2), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$v unparsed', e); print(open('gpurun_out/exp.err').read()[-1500:])
PY
done
