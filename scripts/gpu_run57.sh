#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 12 gpurun_out/$name.log | cut -c1-330; }
run tests_cc python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "small_regions" --timeout 600
run tests_model python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -x
python - <<'PY'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
from crowdsam_b200 import ops
g = torch.Generator().manual_seed(0)
m = (torch.nn.functional.interpolate(torch.randn(64, 1, 64, 64, generator=g), (1024, 1024), mode="bilinear")[:, 0] > 0.3)
m ^= torch.rand(64, 1024, 1024, generator=g) < 0.01
m = m.to(torch.uint8).cuda().contiguous()
for mode in ("holes", "islands"):
    x = m.clone(); ops.remove_small_regions(x, 100, mode); torch.cuda.synchronize()
    x = m.clone(); t0 = time.perf_counter(); c = ops.remove_small_regions(x, 100, mode); torch.cuda.synchronize()
    print(mode, "64 masks 1024x1024:", round((time.perf_counter() - t0) * 1e3, 2), "ms, changed", int(c.sum()))
PY
