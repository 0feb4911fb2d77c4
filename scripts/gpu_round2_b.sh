#!/bin/bash
# Round 2, second pass: bench with priming (graphs on/off), per-shape GEMM table pair vs single-CTA, ncu of the pair kernel.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-shapes > gpurun_out/r2b_pair.json 2> gpurun_out/r2b_pair.err; grep -E "resident|gemm\]" gpurun_out/r2b_pair.err | head -30
CSAM_GEMM_PAIR=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-shapes > gpurun_out/r2b_nopair.json 2> gpurun_out/r2b_nopair.err; grep -E "resident|gemm\]" gpurun_out/r2b_nopair.err | head -30
CSAM_GRAPHS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_nograph.json 2> gpurun_out/r2b_nograph.err; grep -E "resident" gpurun_out/r2b_nograph.err
CSAM_GRAPHS=0 timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -f --import-source on -k 'regex:gemm_pair_kernel' -s 40 -c 6 -o gpurun_out/prof_gemm_pair_r02 python scripts/profile_step.py 1 > gpurun_out/r2b_ncu_pair.log 2>&1; tail -2 gpurun_out/r2b_ncu_pair.log
ncu -i gpurun_out/prof_gemm_pair_r02.ncu-rep --page details > gpurun_out/ncu_details_gemm_pair_r02.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_gemm_pair_r02.ncu-rep > gpurun_out/ncu_summary_r02_pair.csv 2>/dev/null
rm -f gpurun_out/prof_gemm_pair_r02.ncu-rep
python - <<'PY'
import json
for n in ("pair", "nopair", "nograph"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "prof_pass", round(d["profiled_pass_ms_per_step"], 2), "gemm", round(d["kernel_ms_per_step"].get("gemm", 0), 2), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "failed", e)
PY
grep -E "Tensor|tensor|Duration|SM Frequency|L2 Cache Throughput|DRAM Throughput|SM Active|Elapsed Cycles|gemm_pair_kernel" gpurun_out/ncu_details_gemm_pair_r02.txt | head -60
