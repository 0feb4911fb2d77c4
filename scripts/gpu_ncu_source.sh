#!/bin/bash
# Source-level stall sampling of one kernel: usage gpu_ncu_source.sh <kernel regex> <skip> <out tag>
mkdir -p gpurun_out
export PYTHONPATH=$PWD
K=$1; S=$2; T=$3
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -f -k "regex:$K" -s $S -c 1 -o gpurun_out/src_$T python scripts/profile_step.py 1 > gpurun_out/src_$T.log 2>&1
ncu -i gpurun_out/src_$T.ncu-rep --page source --csv > gpurun_out/src_$T.csv 2>/dev/null
rm -f gpurun_out/src_$T.ncu-rep
python - "$T" <<'PY'
import csv, sys, collections
t = sys.argv[1]
rows = list(csv.DictReader(open(f"gpurun_out/src_{t}.csv")))
print(len(rows), "rows; columns:", [c for c in rows[0].keys()][:14])
def num(r, k):
    try: return float(r[k].replace(",", ""))
    except Exception: return 0.0
key = next((c for c in rows[0] if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"), None)
key = key or next(c for c in rows[0] if "Samples" in c)
tot = sum(num(r, key) for r in rows)
rows.sort(key=lambda r: -num(r, key))
print("sampling column:", key, "total", tot)
for r in rows[:40]:
    print(f"{100 * num(r, key) / max(tot, 1):5.1f}%  {r.get('Source', '')[:150]}")
PY
