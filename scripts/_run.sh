for cfg in "0 0" "1 0" "0 1"; do set -- $cfg
CSAM_GEMM_BN256=$1 CSAM_GEMM_PAIR=$2 timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r3g.json 2> gpurun_out/r3g.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3g.json').read().strip().splitlines()[-1])
print('BN256=$1 PAIR=$2', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if k.startswith('gemm')})
PY
done
