"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [--gemm-shapes]"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("csam::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
print("ms,share_pct,launches,avg_us,kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:.3f},{100 * v[1] / tot:.1f},{v[0]},{1e3 * v[1] / v[0]:.1f},{k[:100]}")
