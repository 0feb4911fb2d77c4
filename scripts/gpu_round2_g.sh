#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest4.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_gputest4.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/r2g_ref.json
CSAM_TEST_IMPLS=0 CSAM_TEST_ATTN_IMPLS=0 timeout 2400 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q --deselect tests/test_gpu_model.py::test_config2_vit_l_grid64_vs_reference --deselect tests/test_gpu_model.py::test_full_scale_vit_l_grid32_against_oracle_run > gpurun_out/memcheck_r02b.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02b.log | tail -4
(head -c 4000 gpurun_out/memcheck_r02b.log; echo "[...]"; tail -c 3000 gpurun_out/memcheck_r02b.log) > gpurun_out/memcheck_r02b.short; mv gpurun_out/memcheck_r02b.short gpurun_out/memcheck_r02b.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["roofline"].get("frac_executed"), "traffic", d["roofline"]["traffic"])
print("cpu", d["cpu_baseline"])
for r in d["rooflines"]:
    print("   ", r["kernel"][:44], round(r["frac"], 3), round(r.get("avg_launch_ms", 0), 3), r.get("traffic"))
PY
