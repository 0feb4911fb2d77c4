#!/bin/bash
# Validation of HEAD (third session of round 2, tagged r04): GPU tests, smoke, bench (eager and graph replay),
# launch list + ncu captures of the attention kernels changed since r03.  usage: gpurun -- bash scripts/gpu_validate_r04.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests -q -m gpu --timeout 900 -x --durations=8
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 10 --warmup 3
cp gpurun_out/bench.log gpurun_out/bench_r04.json
CSAM_GRAPHS=1 run bench_graphs python bench.py --steps 10 --warmup 3 --no-cpu-baseline
NB="--kernel-name-base demangled"
FULL="--set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_active.avg --clock-control none $NB -f"
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r04.csv python scripts/profile_step.py 2
run ncu_attn_win ncu $FULL -k 'regex:vit_attention_t' -s 0 -c 1 -o gpurun_out/prof_attn_win_r04 python scripts/profile_step.py 1
run ncu_attn_glob ncu $FULL -k 'regex:vit_attention_t' -s 5 -c 1 -o gpurun_out/prof_attn_glob_r04 python scripts/profile_step.py 1
python scripts/launch_summary.py gpurun_out/launches_r04.csv > gpurun_out/launches_r04_summary.csv
python scripts/ncu_summary.py gpurun_out/prof_*_r04.ncu-rep > gpurun_out/ncu_summary_r04.csv
for f in gpurun_out/prof_*_r04.ncu-rep; do ncu -i $f --page details > ${f%.ncu-rep}.details.txt 2>/dev/null; done
rm -f gpurun_out/prof_*_r04.ncu-rep
python - <<'PY'
import json
for n in ("bench", "bench_graphs"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 2), round(d["value"], 2), round(d["e2e"]["value"], 2), {k: round(v, 2) for k, v in d.get("kernel_ms_per_step", {}).items()})
    except Exception as e:
        print(n, "unparsed", e)
PY
head -30 gpurun_out/launches_r04_summary.csv
