#!/bin/bash
# GEMM variants again, now that the encoders share the machine on two streams and the chip is power-capped: less L2 traffic
# per flop (CTA pairs, 128x256 tiles) could matter more than it did on one stream.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for v in "CSAM_X=0" "CSAM_GEMM_PAIR=1" "CSAM_GEMM_PAIR=3" "CSAM_GEMM_BN256=1" "CSAM_GEMM_L2PF=0" "CSAM_X=1"; do
  env $v timeout 300 python bench.py --steps 15 --warmup 3 --no-cpu-baseline > gpurun_out/exp.json 2> gpurun_out/exp.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/exp.json').read().strip().splitlines()[-1])
    print('$v', round(d['ms_per_step'], 2), 'ms', round(d['value'], 2), 'img/s  e2e', round(d['e2e']['value'], 2), 'clk', d['clocks']['sm_mhz'], 'gemm', round(d['kernel_ms_per_step']['gemm_tensor'], 2))
except Exception as e:
    print('$v unparsed', e); print(open('gpurun_out/exp.err').read()[-800:])
PY
done
