"""Bring-up diagnostics for the tcgen05 GEMM: structured inputs whose wrong outputs reveal which
layout assumption (swizzle, descriptor stride, TMEM lane map) is broken.  Prints only; never asserts."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops as o

torch.manual_seed(0)
dev = "cuda"


def run(M, N, K, split, a=None, w=None, tag="", mn=False):
    a = torch.randn(M, K) if a is None else a
    w = (torch.randn(K, N) if mn else torch.randn(N, K)) if w is None else w
    ah, wh = o.H16.from_f32(a.to(dev), split), o.H16.from_f32(w.to(dev), split)
    ref = ah.float().cpu().double() @ (wh.float().cpu().double() if mn else wh.float().cpu().double().T)
    try:
        out, _ = o.gemm(ah, wh, want_f32=True, impl=0, b_mn_major=mn)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{tag}] M{M} N{N} K{K} split={split} mn={mn}: EXC {e}")
        return
    out = out.cpu().double()
    err = (out - ref).abs()
    print(f"[{tag}] M{M} N{N} K{K} split={split} mn={mn}: max err {err.max():.3e} ref max {ref.abs().max():.3e} "
          f"bad rows {int((err.max(1)[0] > 1e-3).sum())} bad cols {int((err.max(0)[0] > 1e-3).sum())}")
    if err.max() > 1e-3:
        print("   out[0,:8]", out[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        print("   out[:8,0]", out[:8, 0].tolist())
        print("   ref[:8,0]", ref[:8, 0].tolist())
        bad = (err > 1e-3).nonzero()[:6].tolist()
        print("   first bad idx", bad)


# identity-like A: C[m, n] = W[n, m]
A = torch.zeros(128, 64); A[torch.arange(64), torch.arange(64)] = 1.0
run(128, 128, 64, False, a=A, tag="identityA")
run(128, 128, 64, False, tag="k64")
run(128, 128, 16, False, tag="k16")
run(128, 128, 128, False, tag="k128")
run(128, 128, 512, False, tag="k512")
run(256, 256, 256, False, tag="4tiles")
run(128, 64, 64, False, tag="bn64")
run(128, 128, 64, True, tag="split")
run(1000, 1000, 1000, True, tag="ragged")
run(128, 128, 64, False, tag="mn", mn=True)
run(256, 256, 192, True, tag="mn-split", mn=True)
print("launches", __import__("crowdsam_b200.lib", fromlist=["x"]).launch_count())
