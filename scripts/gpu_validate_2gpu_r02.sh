#!/bin/bash
# 2-GPU validation: bench at N = 2 (NCCL all-gather sized from the exchanged maximum) and the torchrun launcher that
# replaces tools/batch_eval.py against a single-process run of the same launcher.
mkdir -p gpurun_out
export PYTHONPATH=$PWD:$PWD/tests
python - <<'PY'
import os, sys
sys.path.insert(0, "tests")
import test_gpu_dropin as t
os.makedirs("/tmp/be", exist_ok=True)
print(t._write_fixture_files("/tmp/be"))
PY
OPTS="test.grid_size 8 test.pos_sim_thresh -1 test.max_prompts 64 test.filter_thresh 2.0"
python -m crowdsam_b200.batch_eval -c /tmp/be/crowdhuman.yaml -s /tmp/be/r1.json --coco /tmp/be/c1.json $OPTS > gpurun_out/be1.log 2>&1; echo "single rc=$?"; tail -2 gpurun_out/be1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 -m crowdsam_b200.batch_eval -c /tmp/be/crowdhuman.yaml -s /tmp/be/r2.json $OPTS > gpurun_out/be2.log 2>&1; echo "torchrun rc=$?"; tail -2 gpurun_out/be2.log
python - <<'PY'
import json
a, b = json.load(open("/tmp/be/r1.json")), json.load(open("/tmp/be/r2.json"))
assert [x["image_id"] for x in a] == [x["image_id"] for x in b] == [100, 101, 102], (a, b)
for x, y in zip(a, b):
    assert set(x) == set(y) == {"image_id", "num_gt", "boxes", "scores", "categories", "rles"}
    assert x["boxes"] == y["boxes"] and x["num_gt"] == y["num_gt"], (x["boxes"], y["boxes"])
    assert all(abs(p - q) <= 1e-5 * max(1.0, abs(p)) for p, q in zip(x["scores"], y["scores"]))
    assert [r["counts"] for r in x["rles"]] == [r["counts"] for r in y["rles"]]
c = json.load(open("/tmp/be/c1.json"))
print("launcher: 1 process == 2 ranks;", sum(len(x["boxes"]) for x in a), "detections,", len(c["annotations"]), "COCO annotations")
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r2_bench_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1]); print('2 GPUs', round(d['value'],2), 'images/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],2))"
