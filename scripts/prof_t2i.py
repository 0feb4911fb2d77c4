"""Profiling driver for the fused token->image decoder attention (csam_dec_t2i)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

dev = "cuda"
torch.manual_seed(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
x = o.H16.from_f32(torch.randn(P * 4096, 256, device=dev), True)
pek = o.H16.from_f32(torch.randn(4096, 128, device=dev), True)
qt = torch.randn(P, 7, 128, device=dev)
wk, wv = torch.randn(128, 256, device=dev) * 0.1, torch.randn(128, 256, device=dev) * 0.1
bv = torch.randn(128, device=dev)
b1 = o.dec_fold_t2i(qt, wk)
wv_t = wv.t().contiguous()
for _ in range(3):
    o.dec_t2i(x, False, pek, b1, P, wv_t, bv)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    o.dec_t2i(x, False, pek, b1, P, wv_t, bv)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"dec_t2i P={P}: {ms*1e3:.1f} us, {P*4096*1024/ms/1e9:.2f} TB/s (1 KB per row)")
