export PYTHONPATH=$PWD
CSAM_GEMM_BN256=1 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" 2>&1 | tail -3
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -k "gemm or config1 or set_image" 2>&1 | tail -3
for m in 0 -1 1; do
  if [ $m = -1 ]; then unset CSAM_GEMM_BN256; else export CSAM_GEMM_BN256=$m; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-shapes > gpurun_out/r2i_$m.json 2> gpurun_out/r2i_$m.err
  grep resident gpurun_out/r2i_$m.err; grep "gemm\]" gpurun_out/r2i_$m.err | head -10
  python -c "
import json; d=json.loads(open('gpurun_out/r2i_$m.json').read().strip().splitlines()[-1]); print('BN256 mode $m: value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'gemm ms', round(d['kernel_ms_per_step']['gemm'],2), 'gemm_tensor frac', round([r for r in d['rooflines'] if r['kernel']=='gemm_tensor'][0]['frac'],3), 'clocks', d['clocks']['sm_mhz'])"
done
