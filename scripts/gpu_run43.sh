#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 4 gpurun_out/$name.log | cut -c1-330; }
run tests_k python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -x
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
timeout 300 python scripts/bench_attn.py 2>&1 | head -3
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for r in d['rooflines']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in('kernel','achieved','frac','avg_launch_ms','share_of_step')})
PY
