export PYTHONPATH=$PWD
CSAM_GEMM_PAIR=2 python -m pytest tests/test_gpu_kernels.py -q -k "gemm_plain or gemm_epilogue or gemm_x3" 2>&1 | tail -3
for m in 0 2; do
  CSAM_GEMM_PAIR=$m python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-shapes > gpurun_out/r2j_$m.json 2> gpurun_out/r2j_$m.err
  grep resident gpurun_out/r2j_$m.err; grep "gemm\]" gpurun_out/r2j_$m.err | head -10
  python -c "
import json; d=json.loads(open('gpurun_out/r2j_$m.json').read().strip().splitlines()[-1]); print('PAIR mode $m: value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'gemm ms', round(d['kernel_ms_per_step']['gemm'],2), 'clocks', d['clocks']['sm_mhz'])"
done
