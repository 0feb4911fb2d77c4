"""K-POST stats + write at P = 1024 (all-survive) on the instance-like fixture of bench.py, for timing sweeps
(CSAM_POST_GX = blocks per mask) and for the ncu traffic capture.  usage: bench_post.py [reps]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from crowdsam_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
low, sel = bench.post_fixture(1024, dev)
for _ in range(2):
    ops.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
    m, _ = ops.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ts = tw = 0.0
for _ in range(reps):
    ev[0].record()
    ops.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
    ev[1].record()
    m, _ = ops.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
    ev[2].record()
    torch.cuda.synchronize()
    ts += ev[0].elapsed_time(ev[1]) / reps
    tw += ev[1].elapsed_time(ev[2]) / reps
nb = 1024 * (2 * 262144.0 + 1048576.0)
print(f"post gx={os.environ.get('CSAM_POST_GX', 'default')} stats {ts:.3f} ms write {tw:.3f} ms combined {nb / ((ts + tw) * 1e-3) / 1e9:.0f} GB/s")
