timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for e in 1 0; do
CSAM_ATTN_REL_INKERNEL=$e timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3n.json 2> gpurun_out/r3n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3n.json').read().strip().splitlines()[-1])
print('INKERNEL=$e', round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if 'attention' in k})
PY
done
