"""Micro-benchmarks of the decoder-shaped GEMMs (isolates epilogue variants)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

dev = "cuda"
torch.manual_seed(0)

def timeit(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

M = 1 << 20
a128 = o.H16.from_f32(torch.randn(M, 128, device=dev), True)
a256 = o.H16.from_f32(torch.randn(M, 256, device=dev), True)
w256x128 = o.H16.from_f32(torch.randn(256, 128, device=dev) * 0.1, True)
w128x256 = o.H16.from_f32(torch.randn(128, 256, device=dev) * 0.1, True)
w256x256 = o.H16.from_f32(torch.randn(256, 256, device=dev) * 0.1, True)
bias = torch.randn(256, device=dev); gam = torch.randn(256, device=dev); bet = torch.randn(256, device=dev)
res_b = torch.randn(4096, 256, device=dev)
res_f = torch.randn(M, 256, device=dev)
res_h = o.H16.from_f32(torch.randn(M, 256, device=dev), True)
of32 = torch.empty(M, 256, device=dev)
oh = o.H16.empty((M, 256), True, dev)
oh128 = o.H16.empty((M, 128), True, dev)
of128 = torch.empty(M, 128, device=dev)
pe128 = torch.randn(4096, 128, device=dev)

def gb(bytes_): return bytes_ / 1e9
cases = [
 ("STD 1Mx256x128 -> f32", lambda: o.gemm(a128, w256x128, bias=bias, out_f32=of32), 0.5 + 4.0 * M * 256 / 1e9 * 1),
 ("STD 1Mx256x128 -> h16 pair", lambda: o.gemm(a128, w256x128, bias=bias, out_h16=oh), 0.5 + 1.0),
 ("STD 1Mx256x128 + res f32 -> f32", lambda: o.gemm(a128, w256x128, bias=bias, residual=res_f, out_f32=of32), 0.5 + 2.0),
 ("LN  1Mx256x128 no res -> h16", lambda: o.gemm(a128, w256x128, bias=bias, epi=1, gamma=gam, beta=bet, eps=1e-5, out_h16=oh), 0.5 + 1.0),
 ("LN  1Mx256x128 res bcast -> h16", lambda: o.gemm(a128, w256x128, bias=bias, residual=res_b, res_mod=4096, epi=1, gamma=gam, beta=bet, eps=1e-5, out_h16=oh), 0.5 + 1.0),
 ("LN  1Mx256x128 res pair -> h16", lambda: o.gemm(a128, w256x128, bias=bias, residual_h16=res_h, epi=1, gamma=gam, beta=bet, eps=1e-5, out_h16=oh), 0.5 + 1.0 + 1.0),
 ("LN  1Mx256x128 res f32 -> f32", lambda: o.gemm(a128, w256x128, bias=bias, residual=res_f, epi=1, gamma=gam, beta=bet, eps=1e-5, out_f32=of32), 0.5 + 2.0),
 ("STD 1Mx128x256 -> f32", lambda: o.gemm(a256, w128x256, out_f32=of128), 1.0 + 0.5),
 ("STD 1Mx128x256 + res bcast -> f32", lambda: o.gemm(a256, w128x256, residual=pe128, res_mod=4096, out_f32=of128), 1.0 + 0.5),
 ("STD 1Mx256x256 -> h16", lambda: o.gemm(a256, w256x256, bias=bias, out_h16=oh), 1.0 + 1.0),
]
for name, fn, gbytes in cases:
    ms = timeit(fn)
    print(f"{name:38s} {ms*1e3:8.1f} us   {gbytes/ (ms*1e-3) /1e3:6.2f} TB/s")
