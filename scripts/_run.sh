timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -4
for w in 0 1; do
CSAM_GEMM_WIDE=$w timeout 600 python bench.py --steps 8 --warmup 3 --gemm-shapes > gpurun_out/r3h_w$w.json 2> gpurun_out/r3h_w$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3h_w$w.json').read().strip().splitlines()[-1])
print('WIDE=$w', round(d['ms_per_step'],2), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if k.startswith('gemm')})
PY
grep "^\[gemm\]" gpurun_out/r3h_w$w.err | head -6
done
