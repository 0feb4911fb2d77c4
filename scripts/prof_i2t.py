"""Profiling driver for the fused image->token decoder layer (csam_dec_i2t_layer) at P = 256 prompts."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

dev = "cuda"
torch.manual_seed(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shared = len(sys.argv) > 2 and sys.argv[2] == "shared"
x = o.H16.from_f32(torch.randn((1 if shared else P) * 4096, 256, device=dev), True)
peq = o.H16.from_f32(torch.randn(4096, 128, device=dev), True)
kt, vt = torch.randn(P, 7, 128, device=dev), torch.randn(P, 7, 128, device=dev)
wq, wo = torch.randn(128, 256, device=dev) * 0.08, torch.randn(256, 128, device=dev) * 0.1
bo, gam, bet = torch.randn(256, device=dev), torch.randn(256, device=dev), torch.randn(256, device=dev)
b1, b2 = o.dec_fold_i2t(kt, vt, wq, wo, bo)
out = o.H16.empty((P * 4096, 256), True, dev)
for _ in range(3):
    o.dec_i2t_layer(x, shared, peq, b1, b2, P, None, gam, bet, 1e-5, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    o.dec_i2t_layer(x, shared, peq, b1, b2, P, None, gam, bet, 1e-5, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"dec_i2t_layer P={P} shared={shared}: {ms*1e3:.1f} us, {P*4096*2048/ms/1e9:.2f} TB/s (2 KB per row)")
