"""Three launches for ncu source-level captures: DINOv2-shaped attention, the k|v|q projection GEMM and the LN GEMM."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o
dev = "cuda"
torch.manual_seed(0)
which = sys.argv[1]
if which == "attn":
    qkv = o.H16.from_f32(torch.randn(5330, 3 * 16 * 64, device=dev), True)
    for _ in range(3):
        o.vit_attention(qkv, 1, 5330, 16, 64, 0.125, None, None, 0, impl=0)
else:
    M = 1 << 20
    a256 = o.H16.from_f32(torch.randn(M, 256, device=dev), True)
    a128 = o.H16.from_f32(torch.randn(M, 128, device=dev), True)
    w384 = o.H16.from_f32(torch.randn(384, 256, device=dev) * 0.1, True)
    w256x128 = o.H16.from_f32(torch.randn(256, 128, device=dev) * 0.1, True)
    res = torch.randn(4096, 384, device=dev)
    out = torch.empty(M, 384, device=dev)
    bias = torch.randn(256, device=dev); gam = torch.randn(256, device=dev); bet = torch.randn(256, device=dev)
    res_h = o.H16.from_f32(torch.randn(M, 256, device=dev), True)
    oh = o.H16.empty((M, 256), True, dev)
    for _ in range(3):
        o.gemm(a256, w384, residual=res, res_mod=4096, out_f32=out)
        o.gemm(a128, w256x128, bias=bias, residual_h16=res_h, epi=1, gamma=gam, beta=bet, eps=1e-5, out_h16=oh)
torch.cuda.synchronize()
