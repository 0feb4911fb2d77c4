#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest3.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_gputest3.log
for gx in 9 17 33 65; do CSAM_POST_GX=$gx python scripts/bench_post.py 5 2>&1 | tail -1; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-shapes > gpurun_out/r2c_default.json 2> gpurun_out/r2c_default.err; grep -E "resident" gpurun_out/r2c_default.err; grep -E "gemm\]" gpurun_out/r2c_default.err | head -12
CSAM_GEMM_PAIR=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_nopair.json 2> gpurun_out/r2c_nopair.err; grep -E "resident" gpurun_out/r2c_nopair.err
python - <<'PY'
import json
for n in ("default", "nopair"):
    d = json.loads(open(f"gpurun_out/r2c_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
          "gemm", round(d["kernel_ms_per_step"].get("gemm", 0), 2), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    for r in d["rooflines"]:
        print("   ", r["kernel"][:44], round(r["frac"], 3), round(r.get("avg_launch_ms", 0), 3))
PY
