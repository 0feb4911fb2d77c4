"""Do two persistent tcgen05 GEMMs on two streams fill each other's last partial wave?"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

dev = "cuda"
torch.manual_seed(0)
M, N, K = 5330, 1024, 1024          # 336 tiles = 2.27 waves on 148 SMs
a = [o.H16.from_f32(torch.randn(M, K, device=dev), True) for _ in range(2)]
w = [o.H16.from_f32(torch.randn(N, K, device=dev) * 0.05, True) for _ in range(2)]
out = [torch.empty(M, N, device=dev) for _ in range(2)]
s2 = torch.cuda.Stream()

def run(two_streams, n=40):
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    if two_streams:
        s2.wait_stream(main)
        for i in range(n // 2):
            o.gemm(a[0], w[0], out_f32=out[0])
            with torch.cuda.stream(s2):
                o.gemm(a[1], w[1], out_f32=out[1])
        main.wait_stream(s2)
    else:
        for i in range(n // 2):
            o.gemm(a[0], w[0], out_f32=out[0])
            o.gemm(a[1], w[1], out_f32=out[1])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for _ in range(2):
    run(False); run(True)
print(f"one stream : {run(False):.1f} us per GEMM")
print(f"two streams: {run(True):.1f} us per GEMM")
