"""Micro-benchmark of csam_vit_attention (tcgen05) for the three shapes of the workload."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

def run(name, groups, tokens, heads, S, split):
    torch.manual_seed(0)
    qkv = o.H16.from_f32(torch.randn(groups * tokens, 3 * heads * 64, device="cuda"), split)
    rel = (torch.randn(2 * S - 1, 64, device="cuda") * 0.1, torch.randn(2 * S - 1, 64, device="cuda") * 0.1) if S else (None, None)
    for _ in range(3):
        o.vit_attention(qkv, groups, tokens, heads, 64, 0.125, rel[0], rel[1], S, impl=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        o.vit_attention(qkv, groups, tokens, heads, 64, 0.125, rel[0], rel[1], S, impl=0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * groups * heads * tokens * tokens * 64
    print(f"{name:8s} split={split} {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (algorithmic)")

only = sys.argv[1] if len(sys.argv) > 1 else None
for split in (True, False):
    run("dino", 1, 5330, 16, 0, split)
    if only == "dino":
        continue
    run("global", 1, 4096, 16, 64, split)
    run("window", 25, 196, 16, 14, split)
