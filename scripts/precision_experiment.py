"""Bounded experiment on the 3x MMA tax of the hi/lo split (round-1 VERDICT task 8), on the CPU, fp64 reference.

Question: can the two correction terms lo*hi + hi*lo of   A*B ~ Ah*Bh + Al*Bh + Ah*Bl   be issued as kind::f8f6f4
MMAs (twice the f16 rate: 2 f16-equivalents per k-step instead of 3) without leaving the 1e-3 parity bar with a >= 5x
margin?  kind::f8f6f4 takes BOTH operands in 8 bits or less, so the corrections become
   2^-s * ( e4m3(Al * 2^s) * e4m3(Bh)  +  e4m3(Ah) * e4m3(Bl * 2^s) )
with a power-of-two pre-scale s per tensor.  This script emulates the operand roundings (torch float8_e4m3fn casts,
exact fp64 accumulation, so only operand precision is measured) on encoder-shaped GEMMs and prints the relative error
of: x1 (fp16 operands), x3 (this repo: fp16 hi + fp16 lo, lo*lo dropped), x2-f8 (the candidate), x2-one-sided
(only one operand split).  Result recorded in DESIGN.md §2.
"""
import torch

torch.manual_seed(0)


def h16(x):
    return x.half().double()


def e4m3(x):
    return x.float().clamp(-448, 448).to(torch.float8_e4m3fn).float().double()


def pow2_scale(x):
    # largest power of two keeping max|x| * 2^s below the e4m3 maximum (448)
    m = x.abs().max().item()
    s = 0
    while m * 2.0 ** (s + 1) <= 448.0:
        s += 1
    return 2.0 ** s


def run(M, N, K, act_scale=1.0):
    a = (torch.randn(M, K, dtype=torch.float64) * act_scale)
    b = torch.randn(N, K, dtype=torch.float64) / K ** 0.5
    ref = a @ b.T
    ah, bh = h16(a), h16(b)
    al, bl = h16(a - ah), h16(b - bh)
    scale = ref.abs().max()

    def err(x):
        return float((x - ref).abs().max() / scale)

    x1 = ah @ bh.T
    x3 = ah @ bh.T + al @ bh.T + ah @ bl.T
    sa, sb = pow2_scale(al), pow2_scale(bl)
    x2f8 = ah @ bh.T + (e4m3(al * sa) @ e4m3(bh).T) / sa + (e4m3(ah) @ e4m3(bl * sb).T) / sb
    x2one = ah @ bh.T + al @ bh.T                      # split activations only
    return err(x1), err(x3), err(x2f8), err(x2one)


if __name__ == "__main__":
    print(f"{'shape':>22s} {'x1':>10s} {'x3':>10s} {'x2-f8':>10s} {'x2-one':>10s}   x1/x2-f8")
    for M, N, K in ((1024, 1024, 1024), (1024, 1024, 4096), (1024, 3072, 1024)):
        e1, e3, ef, eo = run(M, N, K)
        print(f"{M:6d}x{N:5d}x{K:5d}  {e1:10.2e} {e3:10.2e} {ef:10.2e} {eo:10.2e}   {e1 / ef:6.1f}x")
