"""Torch-tensor front end of the C-ABI kernels.  PyTorch is used here for device memory and
streams only; every computation is a libcsam_sm100 kernel."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import lib as L

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

# GEMM / attention implementation switches (validation only): 0 = tcgen05, 1 = SIMT
GEMM_IMPL = int(os.environ.get("CSAM_GEMM_IMPL", "0"))
ATTN_IMPL = int(os.environ.get("CSAM_ATTN_IMPL", "0"))
# P V of the ViT attention: -1 (default) = probabilities and values each as ONE fp16 (one MMA per k-step; measured end to end
# against the oracle: features 1.1e-5, class logits 1.3e-4 relative, the same as mode 0 to within 2e-6); 0 = values as hi+lo
# pair (two MMAs); 1 = probabilities as hi+lo pair too (three MMAs, SS-mode kernel, ~40% slower attention)
ATTN_PSPLIT = int(os.environ.get("CSAM_ATTN_PSPLIT", "-1"))


class Profiler:
    """Per-kernel-class CUDA-event timing on the launching stream (bench.py roofline numbers).
    Each record is (start_event, stop_event, work) with work = algorithmic FLOPs or bytes of that launch."""

    def __init__(self, detail: bool = False):
        self.records = {}
        self.detail = detail      # also key GEMM records by shape

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name: str, start, work: float):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.setdefault(name, []).append((start, e, work))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = [a.elapsed_time(b) for a, b, _ in recs]
            out[name] = {"launches": len(recs), "total_ms": sum(ms), "work": sum(w for _, _, w in recs)}
        return out


PROFILER: Optional[Profiler] = None


def _pb():
    return PROFILER.begin() if PROFILER is not None else None


def _pe(name, tok, work):
    if tok is not None:
        PROFILER.end(name, tok, work)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # the raw handle of torch's current stream on the current device.  torch.cuda.current_stream() builds a Python
    # Stream object through three helper layers (~15 us); with ~430 launches per image that was 6.6 ms of host time
    # per image, in front of every launch
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class H16:
    """fp16 hi (+ optional lo) pair: value = hi + lo (see include/csam.h)."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: Optional[torch.Tensor]):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(shape, split: bool, device="cuda") -> "H16":
        hi = torch.empty(shape, dtype=torch.float16, device=device)
        lo = torch.empty(shape, dtype=torch.float16, device=device) if split else None
        return H16(hi, lo)

    @staticmethod
    def from_f32(t: torch.Tensor, split: bool) -> "H16":
        """Weight preparation at load time (host-side plumbing, not on the hot path)."""
        t = t.detach().float().clamp(-65504.0, 65504.0)
        hi = t.half()
        lo = (t - hi.float()).half() if split else None
        return H16(hi.contiguous(), None if lo is None else lo.contiguous())

    @property
    def shape(self):
        return self.hi.shape

    def view(self, *shape) -> "H16":
        return H16(self.hi.view(*shape), None if self.lo is None else self.lo.view(*shape))

    def rows(self, a: int, b: int) -> "H16":
        return H16(self.hi[a:b], None if self.lo is None else self.lo[a:b])

    def float(self) -> torch.Tensor:
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()


def gemm(a: H16, w: H16, *, bias=None, act=ACT_NONE, residual=None, res_mod=0, row_map=None,
         row_scale=None, col_scale=None, out_f32=None, out_h16: Optional[H16] = None,
         want_f32=False, want_h16=False, impl=None, b_mn_major=False, M=None,
         epi=0, gamma=None, beta=None, eps=0.0, pe=None, pe_mod=0, out2: Optional[H16] = None, hyper=None, masks=None,
         residual_h16: Optional[H16] = None):
    """out = epilogue(a[M,K] @ w[N,K]^T); returns (fp32 or None, H16 or None).
    epi: 0 standard, 1 full-row LayerNorm (N == 256), 2 ConvT1+LN2d+GELU shuffle, 3 ConvT2+GELU+hypernet dot."""
    assert a.hi.dim() == 2 and w.hi.dim() == 2
    M = a.hi.shape[0] if M is None else M
    K = a.hi.shape[1]
    N = w.hi.shape[1] if b_mn_major else w.hi.shape[0]
    split = a.lo is not None and w.lo is not None
    dev = a.hi.device
    if row_map is not None:
        assert not (want_f32 and out_f32 is None) and not (want_h16 and out_h16 is None), \
            "row_map GEMMs need caller-provided outputs"
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((M, N), dtype=torch.float32, device=dev)
    if want_h16 and out_h16 is None:
        out_h16 = H16.empty((M, N), a.lo is not None, dev)
    g = L.GemmArgs()
    g.a_hi, g.a_lo = _p(a.hi), _p(a.lo if split else None)
    g.w_hi, g.w_lo = _p(w.hi), _p(w.lo if split else None)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw = a.hi.stride(0), w.hi.stride(0)
    g.bias, g.row_scale, g.col_scale, g.act = _p(bias), _p(row_scale), _p(col_scale), act
    g.residual, g.ldr, g.res_mod = _p(residual), (residual.stride(0) if residual is not None else 0), res_mod
    g.row_map = _p(row_map)
    g.out_f32, g.ldo = _p(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    if out_h16 is not None:
        g.out_hi, g.out_lo, g.ldh = _p(out_h16.hi), _p(out_h16.lo), out_h16.hi.stride(0)
    g.impl = GEMM_IMPL if impl is None else impl
    g.b_mn_major = 1 if b_mn_major else 0
    g.epi, g.gamma, g.beta, g.eps = epi, _p(gamma), _p(beta), eps
    g.pe, g.ldpe, g.pe_mod = _p(pe), (pe.stride(0) if pe is not None else 0), pe_mod
    if out2 is not None:
        g.out2_hi, g.out2_lo = _p(out2.hi), _p(out2.lo)
        if out_h16 is not None:
            assert out2.hi.stride(0) == out_h16.hi.stride(0)
        g.ldh = out2.hi.stride(0)
    g.hyper, g.masks = _p(hyper), _p(masks)
    if residual_h16 is not None:
        g.res_hi, g.res_lo, g.ldrh = _p(residual_h16.hi), _p(residual_h16.lo), residual_h16.hi.stride(0)
    if epi != 0:
        assert g.impl == 0, "fused epilogues exist on the tcgen05 path only"
    tok = _pb()
    L.check(L.load().csam_gemm(C.byref(g), _stream()), "csam_gemm")
    if tok is not None:
        # classify by arithmetic intensity: the big encoder / DINOv2 GEMMs are tensor-bound, the skinny
        # decoder ones (K <= 256, millions of rows) are HBM-bound and are accounted in bytes
        nop = 2 if split else 1
        nbytes = 2.0 * nop * (M * K + N * K) + (4.0 * M * N if out_f32 is not None else 0.0) \
            + (2.0 * M * N * (2 if (out_h16 is not None and out_h16.lo is not None) else 1) if out_h16 is not None else 0.0) \
            + (4.0 * M * N if (residual is not None and res_mod == 0) or residual_h16 is not None else 0.0) \
            + (4.0 * masks.numel() if masks is not None else 0.0)
        flops = 2.0 * M * N * K
        # >= 100 flop per byte: every encoder / DINOv2 linear layer incl. the attention output projections (AI ~ 160,
        # booked with the HBM-bound class in round 1); the decoder streams (K <= 256, or K = 65536 with M = 4P) sit at ~64
        if flops / nbytes >= 100.0:
            PROFILER.end("gemm_tensor", tok, flops)
        else:
            PROFILER.end("gemm_hbm", tok, nbytes)
        PROFILER.end("gemm", tok, flops)
        if PROFILER.detail:
            PROFILER.end(f"gemm {M}x{N}x{K}", tok, flops)
    return out_f32, out_h16


def patchify(img_u8: torch.Tensor, patch: int, n_side: int, resize_to: int, kpad: int, split: bool) -> H16:
    """img_u8: uint8 [3,h,w] (planar) or [h,w,3] (interleaved, as cv2 / PIL deliver it)."""
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and img_u8.is_contiguous()
    hwc = img_u8.shape[2] == 3 and img_u8.shape[0] != 3
    h, w = (img_u8.shape[0], img_u8.shape[1]) if hwc else (img_u8.shape[1], img_u8.shape[2])
    out = H16.empty((n_side * n_side, kpad), split, img_u8.device)
    L.check(L.load().csam_patchify(_p(img_u8), h, w, 1 if hwc else 0, patch, n_side, resize_to, _p(out.hi), _p(out.lo),
                                   kpad, _stream()), "csam_patchify")
    return out


def layernorm(x: torch.Tensor, gamma=None, beta=None, eps=1e-6, *, add=None, add_mod=0, row_map=None,
              rows_out=None, normalize=True, out_f32=None, want_f32=False, out_h16: Optional[H16] = None,
              want_h16=False, split=False, pe=None, pe_mod=0, out2: Optional[H16] = None, want_out2=False,
              act=ACT_NONE):
    assert x.dim() == 2 and x.dtype == torch.float32
    rows_in, cols = x.shape
    rows_out = (rows_in if row_map is None else row_map.numel()) if rows_out is None else rows_out
    dev = x.device
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((rows_out, cols), dtype=torch.float32, device=dev)
    if want_h16 and out_h16 is None:
        out_h16 = H16.empty((rows_out, cols), split, dev)
    if want_out2 and out2 is None:
        out2 = H16.empty((rows_out, cols), split, dev)
    a = L.LnArgs()
    a.x, a.ldx, a.rows_in = _p(x), x.stride(0), rows_in
    a.add, a.ldadd, a.add_mod = _p(add), (add.stride(0) if add is not None else 0), add_mod
    a.row_map, a.rows_out, a.cols = _p(row_map), rows_out, cols
    a.gamma, a.beta, a.eps, a.normalize = _p(gamma), _p(beta), eps, 1 if normalize else 0
    a.out_f32, a.ldo = _p(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    if out_h16 is not None:
        a.out_hi, a.out_lo, a.ldh = _p(out_h16.hi), _p(out_h16.lo), out_h16.hi.stride(0)
    if out2 is not None:
        assert pe is not None
        a.out2_hi, a.out2_lo = _p(out2.hi), _p(out2.lo)
        if out_h16 is not None:
            assert out2.hi.stride(0) == out_h16.hi.stride(0)
        a.ldh = out2.hi.stride(0)
    a.pe, a.ldpe, a.pe_mod = _p(pe), (pe.stride(0) if pe is not None else 0), pe_mod
    a.act = act
    tok = _pb()
    L.check(L.load().csam_layernorm(C.byref(a), _stream()), "csam_layernorm")
    _pe("layernorm", tok, 4.0 * rows_out * cols)
    return out_f32, out_h16, out2


_attn_scratch = {}


def vit_attention(qkv: H16, groups: int, tokens: int, heads: int, hd: int, scale: float, rel_h=None, rel_w=None,
                  S=0, impl=None, p_split=None) -> H16:
    dev = qkv.hi.device
    out = H16.empty((groups * tokens, heads * hd), qkv.lo is not None, dev)
    a = L.AttnArgs()
    a.qkv_hi, a.qkv_lo, a.ld_qkv = _p(qkv.hi), _p(qkv.lo), qkv.hi.stride(0)
    a.groups, a.tokens, a.heads, a.hd, a.scale = groups, tokens, heads, hd, scale
    a.rel_h, a.rel_w, a.S = _p(rel_h), _p(rel_w), S
    a.out_hi, a.out_lo, a.ld_out = _p(out.hi), _p(out.lo), out.hi.stride(0)
    if rel_h is not None:
        need = L.load().csam_vit_attention_scratch_bytes(groups, tokens, heads, hd, S)
        key = (dev, need)
        if key not in _attn_scratch:
            _attn_scratch[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        a.scratch, a.scratch_bytes = _p(_attn_scratch[key]), need
    a.impl = ATTN_IMPL if impl is None else impl
    a.p_split = ATTN_PSPLIT if p_split is None else int(p_split)
    if impl is None and (hd not in (64, 80) or (hd == 80 and a.p_split > 0)):
        a.impl = 1          # CUDA-core flash kernel; the tcgen05 kernels take head dim 64 and (without p_split) 80
    tok = _pb()
    L.check(L.load().csam_vit_attention(C.byref(a), _stream()), "csam_vit_attention")
    _pe("vit_attention", tok, 4.0 * groups * heads * tokens * tokens * hd)
    return out


def im2col3x3(x: H16, side: int, Cc: int) -> H16:
    out = H16.empty((side * side, 9 * Cc), x.lo is not None, x.hi.device)
    L.check(L.load().csam_im2col3x3(_p(x.hi), _p(x.lo), side, Cc, _p(out.hi), _p(out.lo), _stream()), "csam_im2col3x3")
    return out


def transpose_f32(x: torch.Tensor) -> torch.Tensor:
    rows, cols = x.shape
    out = torch.empty((cols, rows), dtype=torch.float32, device=x.device)
    L.check(L.load().csam_transpose_f32(_p(x), rows, cols, _p(out), _stream()), "csam_transpose_f32")
    return out


def bilinear(x: torch.Tensor, hout: int, wout: int, chlast: bool) -> torch.Tensor:
    """planes [n,h,w] -> [n,hout,wout]  or channels-last [h,w,n] -> [hout,wout,n]."""
    x = x.contiguous()
    if chlast:
        hin, win, n = x.shape
        out = torch.empty((hout, wout, n), dtype=torch.float32, device=x.device)
    else:
        n, hin, win = x.shape
        out = torch.empty((n, hout, wout), dtype=torch.float32, device=x.device)
    L.check(L.load().csam_bilinear(_p(x), n, hin, win, _p(out), hout, wout, 1 if chlast else 0, _stream()),
            "csam_bilinear")
    return out


def prompt_tokens(coords01, labels, gauss, tok5, point_emb, nap) -> torch.Tensor:
    P = coords01.shape[0]
    out = torch.empty((P, 7, 256), dtype=torch.float32, device=coords01.device)
    L.check(L.load().csam_prompt_tokens(_p(coords01), _p(labels), P, _p(gauss), _p(tok5), _p(point_emb), _p(nap),
                                        _p(out), _stream()), "csam_prompt_tokens")
    return out


def _dec_attn(fn_name, q, k, v, B, nq, nk, heads, hd, want_f32, want_h16, split):
    dev = q.device
    Cc = heads * hd
    a = L.DecAttnArgs()
    a.q, a.Bq = _p(q), (1 if q.shape[0] == 1 and B > 1 else B)
    a.k, a.v, a.Bk = _p(k), _p(v), (1 if k.shape[0] == 1 and B > 1 else B)
    a.B, a.nq, a.nk, a.heads, a.hd = B, nq, nk, heads, hd

    def ld(t, n):      # row stride of a [B, n, C] view (e.g. a column slice of a fused projection output)
        assert t.dim() == 3 and t.shape[1] == n and t.shape[2] == Cc and t.stride(2) == 1, "q/k/v must be [B, n, C] views"
        s = t.stride(1) if n > 1 else Cc
        assert t.shape[0] == 1 or t.stride(0) == n * s, "batch stride must equal n * row stride"
        return 0 if s == Cc else s

    a.ldq, a.ldk, a.ldv = ld(q, nq), ld(k, nk), ld(v, nk)
    of = torch.empty((B, nq, Cc), dtype=torch.float32, device=dev) if want_f32 else None
    oh = H16.empty((B, nq, Cc), split, dev) if want_h16 else None
    a.out_f32 = _p(of)
    if oh is not None:
        a.out_hi, a.out_lo = _p(oh.hi), _p(oh.lo)
    tok = _pb()
    L.check(getattr(L.load(), fn_name)(C.byref(a), _stream()), fn_name)
    _pe(fn_name[5:], tok, 4.0 * B * nq * nk * Cc)
    return of, oh


def attn_few_keys(q, k, v, B, nq, nk, heads, hd, want_f32=False, want_h16=False, split=False):
    return _dec_attn("csam_attn_few_keys", q, k, v, B, nq, nk, heads, hd, want_f32, want_h16, split)


def attn_few_queries(q, k, v, B, nq, nk, heads, hd, want_f32=False, want_h16=False, split=False):
    return _dec_attn("csam_attn_few_queries", q, k, v, B, nq, nk, heads, hd, want_f32, want_h16, split)


def dec_fold_i2t(kt: torch.Tensor, vt: torch.Tensor, wq: torch.Tensor, wo: torch.Tensor, bo: Optional[torch.Tensor] = None):
    """kt, vt fp32 [P,7,128]; wq fp32 [128,256]; wo fp32 [256,128] -> (B1 H16 [P*64,384], B2 H16 [P*256,64])
    (csam_dec_fold_i2t: the 7 prompt tokens folded into the operands of the fused image->token layer)."""
    P = kt.shape[0]
    assert kt.shape == (P, 7, 128) and vt.shape == (P, 7, 128) and kt.is_contiguous() and vt.is_contiguous()
    assert wq.shape == (128, 256) and wo.shape == (256, 128) and wq.is_contiguous() and wo.is_contiguous()
    b1 = H16.empty((P * 64, 384), True, kt.device)
    b2 = H16.empty((P * 256, 64), True, kt.device)
    tok = _pb()
    L.check(L.load().csam_dec_fold_i2t(_p(kt), _p(vt), P, _p(wq), _p(wo), _p(bo), _p(b1.hi), _p(b1.lo), _p(b2.hi), _p(b2.lo),
                                       _stream()), "csam_dec_fold_i2t")
    _pe("dec_fold_i2t", tok, 0.0)
    return b1, b2


def dec_i2t_layer(x: H16, x_shared: bool, peq: H16, b1: H16, b2: H16, P: int, bias, gamma, beta, eps: float,
                  out: Optional[H16] = None) -> H16:
    """x' = LayerNorm(x + out_proj(softmax(q_proj(x + pe) k_t^T / 4) v_t)) for every image token of every prompt
    (csam_dec_i2t_layer).  x: H16 [P*4096,256] or [4096,256] when x_shared.  Split (hi + lo) operands only."""
    assert x.lo is not None and peq.lo is not None and b1.lo is not None and b2.lo is not None
    assert x.hi.shape == ((4096 if x_shared else P * 4096), 256) and x.hi.is_contiguous() and x.lo.is_contiguous()
    if out is None:
        out = H16.empty((P * 4096, 256), True, x.hi.device)
    g = L.I2TArgs()
    g.x_hi, g.x_lo, g.x_shared = _p(x.hi), _p(x.lo), 1 if x_shared else 0
    g.peq_hi, g.peq_lo = _p(peq.hi), _p(peq.lo)
    g.b1_hi, g.b1_lo, g.b2_hi, g.b2_lo = _p(b1.hi), _p(b1.lo), _p(b2.hi), _p(b2.lo)
    g.P = P
    g.bias, g.gamma, g.beta, g.eps = _p(bias), _p(gamma), _p(beta), eps
    g.out_hi, g.out_lo = _p(out.hi), _p(out.lo)
    tok = _pb()
    L.check(L.load().csam_dec_i2t_layer(C.byref(g), _stream()), "csam_dec_i2t_layer")
    # algorithmic bytes: read x (hi+lo) once, write x' once
    # algorithmic bytes: the keys once in (unless all prompts share the 4096 input rows: layer 0) and once out
    _pe("dec_i2t_layer_shared" if x_shared else "dec_i2t_layer", tok, float(P) * 4096 * 256 * 4 * (1 if x_shared else 2))
    return out


def dec_fold_t2i(qt: torch.Tensor, wk: torch.Tensor) -> H16:
    """qt fp32 [P,7,128] = q_proj(tokens + pe), wk fp32 [128,256] -> B1 H16 [P*64,384] (csam_dec_fold_t2i)."""
    P = qt.shape[0]
    assert qt.shape == (P, 7, 128) and qt.is_contiguous() and wk.shape == (128, 256) and wk.is_contiguous()
    b1 = H16.empty((P * 64, 384), True, qt.device)
    tok = _pb()
    L.check(L.load().csam_dec_fold_t2i(_p(qt), P, _p(wk), _p(b1.hi), _p(b1.lo), _stream()), "csam_dec_fold_t2i")
    _pe("dec_fold_t2i", tok, 0.0)
    return b1


def dec_t2i(x: H16, x_shared: bool, pek: H16, b1: H16, P: int, wv_t: torch.Tensor, bv, want_f32=False, want_h16=True):
    """Token->image cross attention of P prompts over their 4096 image tokens with k_proj / v_proj folded away
    (csam_dec_t2i) -> (fp32 [P,7,128] or None, H16 [P,7,128] or None): the attention output before out_proj."""
    assert x.lo is not None and pek.lo is not None and b1.lo is not None
    assert x.hi.shape == ((4096 if x_shared else P * 4096), 256) and x.hi.is_contiguous() and x.lo.is_contiguous()
    assert wv_t.shape == (256, 128) and wv_t.is_contiguous()
    dev = x.hi.device
    xbar = torch.empty((P, 64, 256), dtype=torch.float32, device=dev)
    of = torch.empty((P, 7, 128), dtype=torch.float32, device=dev) if want_f32 else None
    oh = H16.empty((P, 7, 128), True, dev) if want_h16 else None
    g = L.T2IArgs()
    g.x_hi, g.x_lo, g.x_shared = _p(x.hi), _p(x.lo), 1 if x_shared else 0
    g.pek_hi, g.pek_lo, g.b1_hi, g.b1_lo = _p(pek.hi), _p(pek.lo), _p(b1.hi), _p(b1.lo)
    g.P, g.xbar, g.wv_t, g.bv = P, _p(xbar), _p(wv_t), _p(bv)
    g.out_f32 = _p(of)
    g.out_hi, g.out_lo = (_p(oh.hi), _p(oh.lo)) if oh is not None else (None, None)
    tok = _pb()
    L.check(L.load().csam_dec_t2i(C.byref(g), _stream()), "csam_dec_t2i")
    if x_shared:
        # layer 0: every prompt attends over the SAME 4096 keys (4 MB, L2-resident): no HBM stream to speak of, the
        # launch is bound by its MMAs -- work = algorithmic flops of scores (K = 384) + pooling per key row
        _pe("dec_t2i_shared", tok, float(P) * 4096 * (2 * 64 * 384 + 2 * 256 * 64))
    else:
        _pe("dec_t2i", tok, float(P) * 4096 * 256 * 4)     # algorithmic bytes: the keys, once
    return of, oh


def upscale_shuffle_ln_gelu(y1: torch.Tensor, P: int, gamma, beta, eps: float, split: bool) -> H16:
    out = H16.empty((P * 16384, 64), split, y1.device)
    L.check(L.load().csam_upscale_shuffle_ln_gelu(_p(y1), P, _p(gamma), _p(beta), eps, _p(out.hi), _p(out.lo),
                                                  _stream()), "csam_upscale_shuffle_ln_gelu")
    return out


def upscale_hyper_masks(y2: torch.Tensor, P: int, hyper_in: torch.Tensor, out=None) -> torch.Tensor:
    if out is None:
        out = torch.empty((P, 4, 256, 256), dtype=torch.float32, device=y2.device)
    L.check(L.load().csam_upscale_hyper_masks(_p(y2), P, _p(hyper_in), _p(out), _stream()), "csam_upscale_hyper_masks")
    return out


def softmax_weights(x: torch.Tensor, split: bool):
    R, n = x.shape
    e = H16.empty((R, n), split, x.device)
    inv = torch.empty((R,), dtype=torch.float32, device=x.device)
    L.check(L.load().csam_softmax_weights(_p(x), R, n, _p(e.hi), _p(e.lo), _p(inv), _stream()), "csam_softmax_weights")
    return e, inv


def select_candidates(iou: torch.Tensor, cls: torch.Tensor):
    P, ncls = iou.shape[0], cls.shape[-1]
    dev = iou.device
    score = torch.empty((P,), dtype=torch.float32, device=dev)
    sel = torch.empty((P,), dtype=torch.int32, device=dev)
    cat = torch.empty((P,), dtype=torch.int32, device=dev)
    L.check(L.load().csam_select_candidates(_p(iou), _p(cls), P, ncls, _p(score), _p(sel), _p(cat), _stream()),
            "csam_select_candidates")
    return score, sel, cat


def _post_args(low, sel, in_size, out_size, thr, off):
    a = L.PostArgs()
    low = low.contiguous()
    a.low, a.P, a.sel = _p(low), low.shape[0], _p(sel)
    a.planes = 4 if sel is not None else 1
    a.in_h, a.in_w = int(in_size[0]), int(in_size[1])
    a.out_h, a.out_w = int(out_size[0]), int(out_size[1])
    a.thr, a.off = float(thr), float(off)
    return a, low


def mask_post_stats(low, sel, in_size, out_size, thr=0.0, off=1.0):
    """-> counts int32 [P,3] (#>thr+off, #>thr-off, #>thr), boxes int32 [P,4]."""
    a, low = _post_args(low, sel, in_size, out_size, thr, off)
    P = low.shape[0]
    counts = torch.empty((P, 3), dtype=torch.int32, device=low.device)
    boxes = torch.empty((P, 4), dtype=torch.int32, device=low.device)
    a.counts, a.boxes = _p(counts), _p(boxes)
    tok = _pb()
    L.check(L.load().csam_mask_post_stats(C.byref(a), _stream()), "csam_mask_post_stats")
    _pe("mask_post_stats", tok, 262144.0 * P + 28.0 * P)      # read the selected 256x256 fp32 plane, write counts+box
    return counts, boxes


def mask_post_write(low, sel, keep, in_size, out_size, thr=0.0, want_masks=True, want_logits=False):
    a, low = _post_args(low, sel, in_size, out_size, thr, 0.0)
    n = low.shape[0] if keep is None else int(keep.numel())
    dev = low.device
    masks = torch.empty((n, a.out_h, a.out_w), dtype=torch.bool, device=dev) if want_masks else None
    logits = torch.empty((n, a.out_h, a.out_w), dtype=torch.float32, device=dev) if want_logits else None
    if n == 0:
        return masks, logits
    a.keep, a.n_keep = _p(keep), n
    a.masks, a.logits = _p(masks), _p(logits)
    tok = _pb()
    L.check(L.load().csam_mask_post_write(C.byref(a), _stream()), "csam_mask_post_write")
    # algorithmic bytes: read one 256x256 fp32 plane + write one bool mask (+ fp32 logits if requested)
    _pe("mask_post_write", tok, n * (262144.0 + a.out_h * a.out_w * (1.0 if want_masks else 0.0)
                                       + a.out_h * a.out_w * (4.0 if want_logits else 0.0)))
    return masks, logits


def box_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """torchvision.ops.nms semantics; returns kept indices (int64) in stable descending-score order."""
    n = boxes.shape[0]
    dev = boxes.device
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dev)
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    keep = torch.empty((n,), dtype=torch.int32, device=dev)
    nk = torch.zeros((1,), dtype=torch.int32, device=dev)
    need = L.load().csam_box_nms_scratch_bytes(n)
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    L.check(L.load().csam_box_nms(_p(boxes), _p(scores), n, float(thr), _p(keep), _p(nk), _p(scratch), need, _stream()),
            "csam_box_nms")
    k = int(nk.item())
    return keep[:k].long()


def mask_overlap(masks: torch.Tensor):
    n, h, w = masks.shape
    dev = masks.device
    inter = torch.empty((n, n), dtype=torch.int32, device=dev)
    area = torch.empty((n,), dtype=torch.int32, device=dev)
    need = L.load().csam_mask_overlap_scratch_bytes(n)
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    L.check(L.load().csam_mask_overlap(_p(masks.contiguous()), n, h, w, _p(inter), _p(area), _p(scratch), need,
                                       _stream()), "csam_mask_overlap")
    return inter, area


def points_occupied(masks: torch.Tensor, flag: torch.Tensor, pts_xy: torch.Tensor) -> torch.Tensor:
    n_pts = pts_xy.shape[0]
    occ = torch.zeros((n_pts,), dtype=torch.uint8, device=pts_xy.device)
    if n_pts == 0 or masks.shape[0] == 0:
        return occ
    n, h, w = masks.shape
    L.check(L.load().csam_points_occupied(_p(masks.contiguous()), n, h, w, _p(flag.contiguous()), _p(pts_xy.contiguous()),
                                          n_pts, _p(occ), _stream()), "csam_points_occupied")
    return occ


def rle_encode(masks: torch.Tensor):
    """Column-major uncompressed RLE of bool masks [n,h,w] -> list of int32 numpy arrays (run lengths)."""
    n, h, w = masks.shape
    if n == 0:
        return []
    masks = masks.contiguous()
    dev = masks.device
    lib = L.load()
    cnt = torch.empty((n,), dtype=torch.int32, device=dev)
    need = lib.csam_rle_scratch_bytes(n, h, w)
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    L.check(lib.csam_rle_count(_p(masks), n, h, w, _p(cnt), _p(scratch), need, _stream()), "csam_rle_count")
    o = np.zeros((n + 1,), dtype=np.int64)
    cnt_h = cnt.cpu().numpy()                       # host decision: the size of the run buffer
    np.cumsum(cnt_h, out=o[1:])
    total = int(o[-1])
    buf = torch.empty((2, total), dtype=torch.int32, device=dev)            # change positions | run lengths
    offs_d = torch.from_numpy(o[:-1]).to(dev)
    L.check(lib.csam_rle_fill(_p(masks), n, h, w, _p(offs_d), _p(cnt), int(cnt_h.max()), _p(buf[0]), _p(buf[1]),
                              _p(scratch), _stream()), "csam_rle_fill")
    runs_h = buf[1].cpu().numpy()
    return [runs_h[o[i]:o[i + 1]] for i in range(n)]


def remove_small_regions(masks: torch.Tensor, area_thresh: int, mode: str) -> torch.Tensor:
    """In-place device version of amg.remove_small_regions for a batch of masks (uint8 [n,h,w], values 0/1):
    mode "holes" fills background components smaller than area_thresh, mode "islands" drops foreground components
    smaller than it (keeping the largest if all are).  Returns changed uint8 [n] (any component below the threshold)."""
    assert masks.dtype == torch.uint8 and masks.dim() == 3 and masks.is_contiguous() and mode in ("holes", "islands")
    n, h, w = masks.shape
    changed = torch.zeros((n,), dtype=torch.uint8, device=masks.device)
    if n == 0:
        return changed
    need = L.load().csam_small_regions_scratch_bytes(n, h, w)
    scratch = torch.empty(need, dtype=torch.uint8, device=masks.device)
    tok = _pb()
    L.check(L.load().csam_remove_small_regions(_p(masks), n, h, w, int(area_thresh), 0 if mode == "holes" else 1,
                                               _p(changed), _p(scratch), need, _stream()), "csam_remove_small_regions")
    _pe("small_regions", tok, float(n) * h * w)
    return changed
