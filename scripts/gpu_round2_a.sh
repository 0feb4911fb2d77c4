#!/bin/bash
# Round 2, first measurement pass: tests, bench A/B (pair GEMM on/off, graphs on/off), launch list.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest2.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_gputest2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_a.err
CSAM_GEMM_PAIR=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nopair.json 2> gpurun_out/r2_bench_nopair.err; tail -2 gpurun_out/r2_bench_nopair.err
CSAM_GRAPHS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nograph.json 2> gpurun_out/r2_bench_nograph.err; tail -2 gpurun_out/r2_bench_nograph.err
CSAM_GRAPHS=0 timeout 600 ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python scripts/profile_step.py 2 > gpurun_out/r2_launches.log 2>&1; tail -2 gpurun_out/r2_launches.log
python scripts/launch_summary.py gpurun_out/launches_r02.csv > gpurun_out/launches_r02_summary.csv; head -30 gpurun_out/launches_r02_summary.csv
python - <<'PY'
import json
for n in ("a", "nopair", "nograph"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "gemm", round(d["kernel_ms_per_step"].get("gemm", 0), 2), "attn", round(d["kernel_ms_per_step"].get("vit_attention", 0), 2),
              "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        for r in d["rooflines"]:
            print("   ", r["kernel"][:40], round(r["frac"], 3), round(r.get("avg_launch_ms", 0), 3))
        if d.get("cpu_baseline"):
            print("    cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
    except Exception as e:
        print(n, "failed", e)
PY
