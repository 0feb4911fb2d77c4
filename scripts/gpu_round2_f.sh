#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_kernels.py -q -k "vit_attention" 2>&1 | tail -4
python -m pytest tests/test_gpu_model.py -q -k "set_image or config1 or config3 or config0" 2>&1 | tail -4
for pv in 0 1; do
  CSAM_ATTN_POLY=$pv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_poly$pv.json 2> gpurun_out/r2f_poly$pv.err; grep resident gpurun_out/r2f_poly$pv.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2f_poly$pv.json').read().strip().splitlines()[-1]); print('poly=$pv value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'attn ms', round(d['kernel_ms_per_step']['vit_attention'],2), 'clocks', d['clocks']['sm_mhz'])"
done
