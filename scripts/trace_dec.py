"""Timeline of the fused decoder kernels from a -DCSAM_TRACE side build (CTA 0's clock64 stamps).

    CSAM_BUILD_OUT=$PWD/crowdsam_b200/_C_trace bash crowdsam_b200/csrc/build.sh -DCSAM_TRACE
    CSAM_LIB_PATH=$PWD/crowdsam_b200/_C_trace/libcsam_sm100.so python scripts/trace_dec.py [t2i|i2t] [P] [shared]
"""
import ctypes as C
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from crowdsam_b200 import lib, ops as o

which = sys.argv[1] if len(sys.argv) > 1 else "t2i"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 296
shared = len(sys.argv) > 3 and sys.argv[3] == "shared"
dev = "cuda"
torch.manual_seed(0)
L = lib.load()
raw = C.CDLL(lib.LIB_PATH)
raw.csam_debug_trace.restype = C.c_int
raw.csam_debug_trace.argtypes = [C.c_void_p, C.c_int]
TAGS, TILES = 160, 64
buf = np.zeros(TAGS * TILES, dtype=np.uint64)


def read():
    raw.csam_debug_trace(buf.ctypes.data, TAGS * TILES)
    tab = buf.reshape(TAGS, TILES).astype(np.int64)
    tag, val = np.nonzero(tab)
    return tab[tag, val], tag, val


xr = 4096 if shared else P * 4096
x = o.H16.from_f32(torch.randn(xr, 256, device=dev), True)
pe = o.H16.from_f32(torch.randn(4096, 128, device=dev), True)
if which == "t2i":
    qt = torch.randn(P, 7, 128, device=dev)
    wk, wv = torch.randn(128, 256, device=dev) * 0.1, torch.randn(128, 256, device=dev) * 0.1
    bv = torch.randn(128, device=dev)
    b1 = o.dec_fold_t2i(qt, wk)
    wv_t = wv.t().contiguous()
    run = lambda: o.dec_t2i(x, shared, pe, b1, P, wv_t, bv)
    names = {10: "x_empty0", 11: "x_empty1", 20: "x_full0", 21: "x_full1", 22: "x_full2", 23: "x_full3", 24: "pek_full0", 25: "pek_full1",
             **{50 + w: f"p_arrive w{w}" for w in range(16)}, **{70 + w: f"p_stored w{w}" for w in range(16)}, 30: "p_full(mma)", 40: "sm_start", 41: "s_full(sm)", 42: "sm_viol_done", 43: "pv_done(sm)"}
else:
    kt, vt = torch.randn(P, 7, 128, device=dev), torch.randn(P, 7, 128, device=dev)
    wq, wo = torch.randn(128, 256, device=dev) * 0.08, torch.randn(256, 128, device=dev) * 0.1
    bo, gam, bet = torch.randn(256, device=dev), torch.randn(256, device=dev), torch.randn(256, device=dev)
    b1, b2 = o.dec_fold_i2t(kt, vt, wq, wo, bo)
    out = o.H16.empty((P * 4096, 256), True, dev)
    run = lambda: o.dec_i2t_layer(x, shared, pe, b1, b2, P, None, gam, bet, 1e-5, out=out)
    names = {**{100 + k: f"empty_kb{k}(prod)" for k in range(6)}, **{110 + k: f"full_kb{k}(mma)" for k in range(6)},
             120: "p_full(mma)", 121: "o_empty(mma)", 130: "e1_start", 131: "s_full(epi)", 132: "resid_loaded", 133: "o_full(epi)",
             134: "o_drained", 135: "stored"}
for _ in range(2):
    run()
read()
run()
t, tag, val = read()
print(f"{which} P={P} shared={shared}: {len(t)} stamps")
order = np.argsort(t, kind="stable")
t, tag, val = t[order], tag[order], val[order]
lo, hi = 8, 12
base = None
for ti, tg, v in zip(t, tag, val):
    if lo <= v < hi:
        if base is None:
            base = ti
        print(f"{ti - base:8d}  tile {v:3d}  {names.get(int(tg), tg)}")
# per-tile period
key = 41 if which == "t2i" else 133
ts = t[tag == key]
if len(ts) > 4:
    d = np.diff(ts)
    print("period (clk) between tiles at", names[key], ": median", int(np.median(d)), "mean", int(d.mean()), "min", int(d.min()), "max", int(d.max()))
