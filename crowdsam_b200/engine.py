"""Host-side orchestration of the Crowd-SAM hot path over libcsam_sm100 kernels.

Data layout in HBM: the residual stream is fp32 token-major [tokens, D];
every GEMM operand is an "h16 pair" (fp16 hi + optional lo) written by the producing kernel's
epilogue; weights are converted once at load.  Reference call sites are cited per method.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, H16

SD = Dict[str, torch.Tensor]


def default_split() -> bool:
    """Precision mode: 'x3' (hi/lo split, fp32-accurate, default) or 'x1' (single-pass fp16)."""
    return os.environ.get("CSAM_PRECISION", "x3") != "x1"


class _Lin:
    """nn.Linear weights prepared for K-GEMM."""

    def __init__(self, sd: SD, name: str, dev, split: bool, w: Optional[torch.Tensor] = None,
                 b: Optional[torch.Tensor] = None, kpad: int = 0):
        w = sd[name + ".weight"] if w is None else w
        w = w.detach().float().reshape(w.shape[0], -1)
        if kpad and w.shape[1] < kpad:
            w = torch.cat([w, w.new_zeros(w.shape[0], kpad - w.shape[1])], dim=1)
        if b is None and (name + ".bias") in sd:
            b = sd[name + ".bias"]
        self.w = H16.from_f32(w.to(dev), split)
        self.b = None if b is None else b.detach().float().contiguous().to(dev)

    def __call__(self, a: H16, **kw):
        return ops.gemm(a, self.w, bias=self.b, **kw)


def _f(sd: SD, name: str, dev) -> torch.Tensor:
    return sd[name].detach().float().contiguous().to(dev)


def run_steps(gen):
    """Exhaust a `forward_steps` generator on the current stream and return its result."""
    while True:
        try:
            next(gen)
        except StopIteration as stop:
            return stop.value


_side_streams: Dict[int, "torch.cuda.Stream"] = {}
_ORDER = (0, 1) if os.environ.get("CSAM_SIDE_FIRST", "1") == "0" else (1, 0)


def two_streams_enabled() -> bool:
    """The SAM encoder and the DINOv2 encoder of an image are independent until the decoder binds both
    (predictor.py:88-110 runs them back to back).  Default: enqueue them on two streams so that each fills the other's
    idle SMs -- last partial waves of the N = 1024 GEMMs (2.1-2.3 waves of tiles on 148 SMs), LayerNorm launches, kernel
    ramps.  CSAM_TWO_STREAMS=0 keeps one stream.  Not used while a kernel-class profiler is timing launches (per-launch
    event times are only meaningful serialised)."""
    return os.environ.get("CSAM_TWO_STREAMS", "1") != "0" and ops.PROFILER is None


def interleave_two_streams(gen_main, gen_side, dev):
    """Drive two `forward_steps` generators, `gen_main` on the current stream and `gen_side` on a per-device side stream,
    alternating block by block so that both streams stay fed; the current stream waits for the side stream at the end.
    -> (result_main, result_side).  Tensors the side generator returns were allocated on the side stream: the caller
    must `record_stream` them on the stream that consumes them."""
    main = torch.cuda.current_stream(dev)
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    side = _side_streams.get(idx)
    if side is None:
        side = _side_streams[idx] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("CSAM_SIDE_PRIO", "0")))
    side.wait_stream(main)
    res = [None, None]
    live = [True, True]
    gens = (gen_main, gen_side)
    while live[0] or live[1]:
        for k in _ORDER:                        # side first: its result is needed last
            if not live[k]:
                continue
            try:
                if k == 1:
                    with torch.cuda.stream(side):
                        next(gens[k])
                else:
                    next(gens[k])
            except StopIteration as stop:
                res[k], live[k] = stop.value, False
    main.wait_stream(side)
    return res[0], res[1]


# ============================================================================================
# SAM ViT image encoder                       (segment_anything_cs/modeling/image_encoder.py)
# ============================================================================================
class SamEncoder:
    def __init__(self, sd: SD, depth: int, heads: int, global_idx: Sequence[int], dev, split: bool,
                 prefix: str = "image_encoder"):
        p = prefix
        self.dev, self.split = dev, split
        self.depth, self.heads, self.glob = depth, heads, tuple(global_idx)
        D = sd[f"{p}.pos_embed"].shape[-1]
        self.D, self.hd = D, D // heads
        self.patch = _Lin(sd, f"{p}.patch_embed.proj", dev, split)           # :387-395
        self.pos = _f(sd, f"{p}.pos_embed", dev).reshape(4096, D)             # :108-109
        self.blocks = []
        for i in range(depth):
            b = f"{p}.blocks.{i}"
            self.blocks.append(dict(
                n1w=_f(sd, b + ".norm1.weight", dev), n1b=_f(sd, b + ".norm1.bias", dev),
                qkv=_Lin(sd, b + ".attn.qkv", dev, split), proj=_Lin(sd, b + ".attn.proj", dev, split),
                rel_h=_f(sd, b + ".attn.rel_pos_h", dev), rel_w=_f(sd, b + ".attn.rel_pos_w", dev),
                n2w=_f(sd, b + ".norm2.weight", dev), n2b=_f(sd, b + ".norm2.bias", dev),
                lin1=_Lin(sd, b + ".mlp.lin1", dev, split), lin2=_Lin(sd, b + ".mlp.lin2", dev, split)))
        self.neck0 = _Lin(sd, f"{p}.neck.0", dev, split)
        self.neck1 = (_f(sd, f"{p}.neck.1.weight", dev), _f(sd, f"{p}.neck.1.bias", dev))
        w2 = sd[f"{p}.neck.2.weight"].detach().float().permute(0, 2, 3, 1).reshape(256, 9 * 256)
        self.neck2 = _Lin(sd, f"{p}.neck.2", dev, split, w=w2)
        self.neck3 = (_f(sd, f"{p}.neck.3.weight", dev), _f(sd, f"{p}.neck.3.bias", dev))
        # window partition map (image_encoder.py:243-264): row (win, i, j) -> token or -1 (zero pad)
        win = 14
        nwin = 5
        idx = torch.full((nwin * nwin * win * win,), -1, dtype=torch.int32)
        r = 0
        for wy in range(nwin):
            for wx in range(nwin):
                for i in range(win):
                    for j in range(win):
                        y, x = wy * win + i, wx * win + j
                        if y < 64 and x < 64:
                            idx[r] = y * 64 + x
                        r += 1
        self.winmap = idx.to(dev)

    def forward(self, img_u8_chw: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """uint8 [3,h,w] on device -> (features fp32 [1,256,64,64], token-major fp32 [4096,256]).
        ImageEncoderViT.forward (image_encoder.py:106-116) fused with Sam.preprocess (sam.py:163-173)."""
        return run_steps(self.forward_steps(img_u8_chw))

    def forward_steps(self, img_u8_chw: torch.Tensor):
        """forward() as a generator that yields after every transformer block (its launches are enqueued on the stream
        that is current when the generator is resumed) and returns the result: lets a caller interleave the launches of
        two independent encoders on two streams (`interleave_two_streams`)."""
        D, heads, hd, split = self.D, self.heads, self.hd, self.split
        patches = ops.patchify(img_u8_chw, 16, 64, 0, 768, split)
        x, _ = self.patch(patches, residual=self.pos, want_f32=True)
        for i, blk in enumerate(self.blocks):
            yield
            is_glob = i in self.glob
            if is_glob:
                _, y, _ = ops.layernorm(x, blk["n1w"], blk["n1b"], 1e-6, want_h16=True, split=split)
                _, qkv = blk["qkv"](y, want_h16=True)
                a = ops.vit_attention(qkv, 1, 4096, heads, hd, hd ** -0.5, blk["rel_h"], blk["rel_w"], 64)
                blk["proj"](a, residual=x, out_f32=x)
            else:
                _, y, _ = ops.layernorm(x, blk["n1w"], blk["n1b"], 1e-6, row_map=self.winmap, want_h16=True, split=split)
                _, qkv = blk["qkv"](y, want_h16=True)
                a = ops.vit_attention(qkv, 25, 196, heads, hd, hd ** -0.5, blk["rel_h"], blk["rel_w"], 14)
                blk["proj"](a, residual=x, out_f32=x, row_map=self.winmap)
            _, y, _ = ops.layernorm(x, blk["n2w"], blk["n2b"], 1e-6, want_h16=True, split=split)
            _, h = blk["lin1"](y, act=ACT_GELU, want_h16=True)
            blk["lin2"](h, residual=x, out_f32=x)
        # neck (image_encoder.py:88-104): 1x1 conv -> LN2d -> 3x3 conv -> LN2d, all bias-free convs
        _, xh, _ = ops.layernorm(x, normalize=False, want_h16=True, split=split)
        z, _ = self.neck0(xh, want_f32=True)
        _, zh, _ = ops.layernorm(z, self.neck1[0], self.neck1[1], 1e-6, want_h16=True, split=split)
        cols = ops.im2col3x3(zh, 64, 256)
        z2, _ = self.neck2(cols, want_f32=True)
        feat_tok, _, _ = ops.layernorm(z2, self.neck3[0], self.neck3[1], 1e-6, want_f32=True)
        feats = ops.transpose_f32(feat_tok).view(1, 256, 64, 64)
        return feats, feat_tok


# ============================================================================================
# DINOv2 ViT-L/14 forward_features               (dinov2/dinov2/models/vision_transformer.py)
# ============================================================================================
class DinoEncoder:
    def __init__(self, sd: SD, depth: int, heads: int, dev, split: bool):
        self.dev, self.split, self.depth, self.heads = dev, split, depth, heads
        D = sd["cls_token"].shape[-1]
        self.D, self.hd = D, D // heads
        self.patch = _Lin(sd, "patch_embed.proj", dev, split, kpad=592)
        # interpolate_pos_encoding (vision_transformer.py:179-211): per-resolution constant, built once
        pe = sd["pos_embed"].detach().float().cpu()
        N = pe.shape[1] - 1
        M = int(math.sqrt(N))
        s = float(73 + 0.1) / M
        patch_pe = torch.nn.functional.interpolate(pe[:, 1:].reshape(1, M, M, D).permute(0, 3, 1, 2), mode="bicubic",
                                                   antialias=False, scale_factor=(s, s))
        assert patch_pe.shape[-2:] == (73, 73)
        self.pos_patch = patch_pe.permute(0, 2, 3, 1).reshape(73 * 73, D).contiguous().to(dev)
        self.cls_row = (sd["cls_token"].detach().float().cpu().reshape(1, D) + pe[0, :1]).contiguous().to(dev)
        self.blocks = []
        for i in range(depth):
            b = f"blocks.{i}"
            self.blocks.append(dict(
                n1w=_f(sd, b + ".norm1.weight", dev), n1b=_f(sd, b + ".norm1.bias", dev),
                qkv=_Lin(sd, b + ".attn.qkv", dev, split), proj=_Lin(sd, b + ".attn.proj", dev, split),
                ls1=_f(sd, b + ".ls1.gamma", dev),
                n2w=_f(sd, b + ".norm2.weight", dev), n2b=_f(sd, b + ".norm2.bias", dev),
                fc1=_Lin(sd, b + ".mlp.fc1", dev, split), fc2=_Lin(sd, b + ".mlp.fc2", dev, split),
                ls2=_f(sd, b + ".ls2.gamma", dev)))
        self.norm = (_f(sd, "norm.weight", dev), _f(sd, "norm.bias", dev))

    def forward(self, img_u8_chw: torch.Tensor) -> Tuple[torch.Tensor, H16]:
        """-> (x_norm_patchtokens fp32 [5329, D], same as h16 pair).  predictor.py:104-106: the
        SAM-normalised zero-padded image is resized bilinearly to 1022x1022 inside patchify."""
        return run_steps(self.forward_steps(img_u8_chw))

    def forward_steps(self, img_u8_chw: torch.Tensor):
        """forward() as a generator yielding after every block (see SamEncoder.forward_steps)."""
        D, heads, hd, split = self.D, self.heads, self.hd, self.split
        n = 73 * 73
        patches = ops.patchify(img_u8_chw, 14, 73, 1022, 592, split)
        x = torch.empty((n + 1, D), dtype=torch.float32, device=self.dev)
        x[:1].copy_(self.cls_row)
        self.patch(patches, residual=self.pos_patch, out_f32=x[1:])
        for blk in self.blocks:
            yield
            _, y, _ = ops.layernorm(x, blk["n1w"], blk["n1b"], 1e-6, want_h16=True, split=split)
            _, qkv = blk["qkv"](y, want_h16=True)
            a = ops.vit_attention(qkv, 1, n + 1, heads, hd, hd ** -0.5)
            blk["proj"](a, col_scale=blk["ls1"], residual=x, out_f32=x)
            _, y, _ = ops.layernorm(x, blk["n2w"], blk["n2b"], 1e-6, want_h16=True, split=split)
            _, h = blk["fc1"](y, act=ACT_GELU, want_h16=True)
            blk["fc2"](h, col_scale=blk["ls2"], residual=x, out_f32=x)
        out, outh, _ = ops.layernorm(x, self.norm[0], self.norm[1], 1e-6, want_f32=True, want_h16=True, split=split)
        return out[1:], outh.rows(1, n + 1)


# ============================================================================================
# Prompt encoder + two-way mask decoder + PWD-Net heads
#   (modeling/prompt_encoder.py, transformer.py, mask_decoder.py)
# ============================================================================================
class _Attn:
    def __init__(self, sd, name, dev, split):
        self.q = _Lin(sd, name + ".q_proj", dev, split)
        self.k = _Lin(sd, name + ".k_proj", dev, split)
        self.v = _Lin(sd, name + ".v_proj", dev, split)
        self.o = _Lin(sd, name + ".out_proj", dev, split)


class _MLP:
    def __init__(self, sd, name, n, dev, split):
        self.layers = [_Lin(sd, f"{name}.layers.{i}", dev, split) for i in range(n)]

    def __call__(self, a: H16, out_f32=None, residual=None):
        """ReLU between layers (mask_decoder.py:203-253, dropout inactive in eval)."""
        for i, lin in enumerate(self.layers):
            if i < len(self.layers) - 1:
                _, a = lin(a, act=ACT_RELU, want_h16=True)
            else:
                f, _ = lin(a, out_f32=out_f32, want_f32=out_f32 is None, residual=residual)
                return f


class MaskDecoderEngine:
    def __init__(self, sd: SD, dev, split: bool):
        self.dev, self.split = dev, split
        m, t, pe = "mask_decoder", "mask_decoder.transformer", "prompt_encoder"
        self.n_class = sd[f"{m}.point_classifier.layers.1.weight"].shape[0]
        self.gauss = _f(sd, f"{pe}.pe_layer.positional_encoding_gaussian_matrix", dev)
        self.point_emb = torch.cat([_f(sd, f"{pe}.point_embeddings.0.weight", dev),
                                    _f(sd, f"{pe}.point_embeddings.1.weight", dev)], dim=0).contiguous()
        self.nap = _f(sd, f"{pe}.not_a_point_embed.weight", dev).reshape(256)
        self.no_mask = _f(sd, f"{pe}.no_mask_embed.weight", dev).reshape(1, 256)
        self.tok5 = torch.cat([_f(sd, f"{m}.iou_token.weight", dev), _f(sd, f"{m}.mask_tokens.weight", dev)], 0).contiguous()
        # dense PE (prompt_encoder.py:64-73,198-209) is a constant of the weights: built once at load
        g = sd[f"{pe}.pe_layer.positional_encoding_gaussian_matrix"].detach().float().cpu()
        grid = torch.ones((64, 64), dtype=torch.float32)
        yy = (grid.cumsum(dim=0) - 0.5) / 64
        xx = (grid.cumsum(dim=1) - 0.5) / 64
        c = (2 * torch.stack([xx, yy], dim=-1) - 1) @ g
        c = 2 * math.pi * c
        pe_map = torch.cat([torch.sin(c), torch.cos(c)], dim=-1)                 # [64,64,256]
        self.pe_tok = pe_map.reshape(4096, 256).contiguous().to(dev)
        self.dense_pe = pe_map.permute(2, 0, 1).unsqueeze(0).contiguous().to(dev)
        self.layers = []
        for i in range(2):
            Lp = f"{t}.layers.{i}"
            self.layers.append(dict(
                sa=_Attn(sd, Lp + ".self_attn", dev, split),
                n1=(_f(sd, Lp + ".norm1.weight", dev), _f(sd, Lp + ".norm1.bias", dev)),
                t2i=_Attn(sd, Lp + ".cross_attn_token_to_image", dev, split),
                n2=(_f(sd, Lp + ".norm2.weight", dev), _f(sd, Lp + ".norm2.bias", dev)),
                lin1=_Lin(sd, Lp + ".mlp.lin1", dev, split), lin2=_Lin(sd, Lp + ".mlp.lin2", dev, split),
                n3=(_f(sd, Lp + ".norm3.weight", dev), _f(sd, Lp + ".norm3.bias", dev)),
                n4=(_f(sd, Lp + ".norm4.weight", dev), _f(sd, Lp + ".norm4.bias", dev)),
                i2t=_Attn(sd, Lp + ".cross_attn_image_to_token", dev, split)))
        self.final = _Attn(sd, f"{t}.final_attn_token_to_image", dev, split)
        self.nf = (_f(sd, f"{t}.norm_final_attn.weight", dev), _f(sd, f"{t}.norm_final_attn.bias", dev))
        # (keys + pe) only ever feeds linear projections, and pe is a constant of the weights, so
        # proj(keys + pe) = keys W^T + (pe W^T + b): the second term is precomputed here once and enters the
        # projection GEMM as a row-periodic residual.  The [P,4096,256] "keys + pe" operand never exists.
        def pe_proj(name):
            w = sd[name + ".weight"].detach().double().cpu()
            b = sd[name + ".bias"].detach().double().cpu()
            return (self.pe_tok.double().cpu() @ w.T + b).float().contiguous().to(dev)

        self.pe_k = [pe_proj(f"{t}.layers.{i}.cross_attn_token_to_image.k_proj") for i in range(2)]
        self.pe_qi = [pe_proj(f"{t}.layers.{i}.cross_attn_image_to_token.q_proj") for i in range(2)]
        self.pe_kf = pe_proj(f"{t}.final_attn_token_to_image.k_proj")
        # Projections that read the same keys are one GEMM: layer 1 needs k, v (token->image) and q (image->
        # token) of keys1, the final attention k, v of keys2.  One pass over the [P*4096, 256] operand
        # instead of three / two; the attention kernels read column slices of the fused output.
        # Row-periodic residual = (pe_k | b_v | pe_qi) resp. (pe_kf | b_v).
        def cat_w(names):
            return H16.from_f32(torch.cat([sd[n + ".weight"].detach().float() for n in names], 0).contiguous().to(dev), split)

        L1 = f"{t}.layers.1"
        fin = f"{t}.final_attn_token_to_image"
        self.w_kvq1 = cat_w([L1 + ".cross_attn_token_to_image.k_proj", L1 + ".cross_attn_token_to_image.v_proj",
                             L1 + ".cross_attn_image_to_token.q_proj"])
        bv1 = sd[L1 + ".cross_attn_token_to_image.v_proj.bias"].detach().float().to(dev)
        self.res_kvq1 = torch.cat([self.pe_k[1], bv1[None, :].expand(4096, -1), self.pe_qi[1]], 1).contiguous()
        # Fused image->token layer (csam_dec_i2t_layer, split precision only): q_proj / out_proj are folded into
        # per-prompt operands, so layer 1 needs only k | v of keys1 and the raw fp32 projection weights.
        self.fused_i2t = split and os.environ.get("CSAM_DEC_FUSED_I2T", "1") != "0"
        self.w_kv1 = cat_w([L1 + ".cross_attn_token_to_image.k_proj", L1 + ".cross_attn_token_to_image.v_proj"])
        self.res_kv1 = torch.cat([self.pe_k[1], bv1[None, :].expand(4096, -1)], 1).contiguous()
        for i, Lr in enumerate(self.layers):
            Lp = f"{t}.layers.{i}.cross_attn_image_to_token"
            Lr["i2t_wq"] = _f(sd, Lp + ".q_proj.weight", dev)
            Lr["i2t_wo"] = _f(sd, Lp + ".out_proj.weight", dev)
            Lr["peq_h"] = H16.from_f32(self.pe_qi[i], True)
        # Fused token->image attention (csam_dec_t2i): k_proj is folded into the query side, v_proj is applied to
        # the softmax-pooled keys, so no [P,4096,256] k | v stream is produced at all.
        self.fused_t2i = self.fused_i2t and os.environ.get("CSAM_DEC_FUSED_T2I", "1") != "0"
        self.t2i_fold = []
        for name, pek in ((f"{t}.layers.0.cross_attn_token_to_image", self.pe_k[0]),
                          (f"{t}.layers.1.cross_attn_token_to_image", self.pe_k[1]), (fin, self.pe_kf)):
            self.t2i_fold.append(dict(wk=_f(sd, name + ".k_proj.weight", dev),
                                      wv_t=_f(sd, name + ".v_proj.weight", dev).t().contiguous(),
                                      bv=_f(sd, name + ".v_proj.bias", dev), pek_h=H16.from_f32(pek, True)))
        self.w_kvf = cat_w([fin + ".k_proj", fin + ".v_proj"])
        bvf = sd[fin + ".v_proj.bias"].detach().float().to(dev)
        self.res_kvf = torch.cat([self.pe_kf, bvf[None, :].expand(4096, -1)], 1).contiguous()
        # ConvTranspose2d(k2,s2) as GEMM: N index = (dy*2+dx)*C_out + o   (mask_decoder.py:56-62)
        w1 = sd[f"{m}.output_upscaling.0.weight"].detach().float().permute(2, 3, 1, 0).reshape(256, 256)
        b1 = sd[f"{m}.output_upscaling.0.bias"].detach().float().repeat(4)
        self.ct1 = _Lin(sd, "", dev, split, w=w1, b=b1)
        self.up_ln = (_f(sd, f"{m}.output_upscaling.1.weight", dev), _f(sd, f"{m}.output_upscaling.1.bias", dev))
        w2 = sd[f"{m}.output_upscaling.3.weight"].detach().float().permute(2, 3, 1, 0).reshape(128, 64)
        b2 = sd[f"{m}.output_upscaling.3.bias"].detach().float().repeat(4)
        self.ct2 = _Lin(sd, "", dev, split, w=w2, b=b2)
        self.hyper = [_MLP(sd, f"{m}.output_hypernetworks_mlps.{i}", 3, dev, split) for i in range(4)]
        self.iou_head = _MLP(sd, f"{m}.iou_prediction_head", 3, dev, split)
        self.dino_proj = _Lin(sd, f"{m}.dino_proj", dev, split)
        self.par_iou = _MLP(sd, f"{m}.parallel_iou_head", 3, dev, split)
        self.point_cls = _MLP(sd, f"{m}.point_classifier", 2, dev, split)
        self._maps: Dict[int, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self.img = None
        self.img_gen = 0                  # bumped whenever the bound image changes (set_image or a graph replay)
        from .graphs import GraphCache

        self.graphs = GraphCache(2)       # captured set_image regions (per image shape), each owning its decode graphs

    # ---- per-image, prompt-independent work -------------------------------------------------
    def set_image(self, feat_tok: torch.Tensor, dino_tok_h: H16):
        """Hoists everything that does not depend on the prompt (SURVEY.md §8a A7 'redundancy'):
        src0 = features + no_mask_embed (mask_decoder.py:160-161), layer-0 k/v of token->image
        and q of image->token (identical for every prompt), dino_proj + bilinear 73->256
        (mask_decoder.py:187-188)."""
        split = self.split
        L0 = self.layers[0]
        keys0, keys0_h, _ = ops.layernorm(feat_tok, normalize=False, add=self.no_mask, add_mod=1,
                                          want_f32=True, want_h16=True, split=split)
        k0 = v0 = q0 = None
        if not self.fused_t2i:
            k0, _ = ops.gemm(keys0_h, L0["t2i"].k.w, residual=self.pe_k[0], want_f32=True)
            v0, _ = L0["t2i"].v(keys0_h, want_f32=True)
            k0, v0 = k0.view(1, 4096, 128), v0.view(1, 4096, 128)
        if not self.fused_i2t:
            q0, _ = ops.gemm(keys0_h, L0["i2t"].q.w, residual=self.pe_qi[0], want_f32=True)
            q0 = q0.view(1, 4096, 128)
        dproj_h = dmap_h = None
        if dino_tok_h is not None:
            dproj, dproj_h = self.dino_proj(dino_tok_h, want_f32=True, want_h16=True)  # [5329,256]
            planes = ops.transpose_f32(dproj).view(256, 73, 73)
            dmap = ops.bilinear(planes, 256, 256, chlast=False)                        # [256,256,256]
            _, dmap_h, _ = ops.layernorm(dmap.view(256 * 32, 2048), normalize=False, want_h16=True, split=split)
            dmap_h = dmap_h.view(256, 65536)
        self.img = dict(keys0=keys0, keys0_h=keys0_h, k0=k0, v0=v0, q0=q0, dproj_h=dproj_h, dmap_h=dmap_h)
        self.img_gen += 1

    def fg_logits(self) -> torch.Tensor:
        """predict_fg_map (predictor.py:113-121) -> fp32 [1,n_class,256,256]."""
        if self.img["dproj_h"] is None:
            raise RuntimeError("predict_fg_map needs DINOv2 features (the predictor was built without dino_model)")
        logits = self.point_cls(self.img["dproj_h"])                                   # [5329,n_class]
        planes = ops.transpose_f32(logits).view(self.n_class, 73, 73)
        return ops.bilinear(planes, 256, 256, chlast=False).view(1, self.n_class, 256, 256)

    def _gather_maps(self, P: int):
        if P not in self._maps:
            p = torch.arange(P, dtype=torch.int32).repeat_interleave(4)
            l = torch.arange(4, dtype=torch.int32).repeat(P)
            self._maps[P] = ((p * 7).to(self.dev), (p * 7 + 1 + l).to(self.dev))
        return self._maps[P]

    # ---- per-prompt work -------------------------------------------------------------------------
    def decode(self, coords01: torch.Tensor, labels: torch.Tensor):
        """coords01 fp32 [P,2] = (pt+0.5)/1024 (fp64 math on the host), labels int32 [P].
        -> low-res masks fp32 [P,4,256,256], iou [P,4], cls [P,4,n_class]
        (MaskDecoder.predict_masks, mask_decoder.py:138-199; TwoWayTransformer transformer.py:62-192)."""
        assert self.img is not None, "set_image first"
        split, I = self.split, self.img
        P = coords01.shape[0]
        T = 7 * P
        ln = ops.layernorm
        tokens = ops.prompt_tokens(coords01, labels, self.gauss, self.tok5, self.point_emb, self.nap).view(T, 256)
        _, tok_h, _ = ln(tokens, normalize=False, want_h16=True, split=split)
        queries, q_h, q_pe_h = tokens, tok_h, tok_h
        keys_f32, keys_h = None, None
        for li, Lr in enumerate(self.layers):
            # (1) token self-attention (transformer.py:163-169)
            sa = Lr["sa"]
            qs, _ = sa.q(q_pe_h if li else q_h, want_f32=True)
            ks, _ = sa.k(q_pe_h if li else q_h, want_f32=True)
            vs, _ = sa.v(q_h, want_f32=True)
            _, a = ops.attn_few_keys(qs.view(P, 7, 256), ks.view(P, 7, 256), vs.view(P, 7, 256), P, 7, 7, 8, 32,
                                     want_h16=True, split=split)
            pre, _ = sa.o(a.view(T, 256), residual=(queries if li else None), want_f32=True)
            queries, q_h, q_pe_h = ln(pre, Lr["n1"][0], Lr["n1"][1], 1e-5, want_f32=True, want_h16=True, split=split,
                                      pe=tokens, want_out2=True)
            # (2) token -> image cross attention (transformer.py:171-176)
            ta = Lr["t2i"]
            qc, _ = ta.q(q_pe_h, want_f32=True)
            if self.fused_t2i:
                F = self.t2i_fold[li]
                b1t = ops.dec_fold_t2i(qc.view(P, 7, 128), F["wk"])
                _, a = ops.dec_t2i(I["keys0_h"] if li == 0 else keys_h, li == 0, F["pek_h"], b1t, P, F["wv_t"], F["bv"])
            elif li == 0:
                kc, vc = I["k0"], I["v0"]
            elif self.fused_i2t:
                kv1, _ = ops.gemm(keys_h, self.w_kv1, residual=self.res_kv1, res_mod=4096, want_f32=True)
                kv1 = kv1.view(P, 4096, 256)
                kc, vc = kv1[:, :, 0:128], kv1[:, :, 128:256]
            else:
                kvq, _ = ops.gemm(keys_h, self.w_kvq1, residual=self.res_kvq1, res_mod=4096, want_f32=True)
                kvq = kvq.view(P, 4096, 384)
                kc, vc, qi1 = kvq[:, :, 0:128], kvq[:, :, 128:256], kvq[:, :, 256:384]
            if not self.fused_t2i:
                _, a = ops.attn_few_queries(qc.view(P, 7, 128), kc, vc, P, 7, 4096, 8, 16, want_h16=True, split=split)
            pre, _ = ta.o(a.view(T, 128), residual=queries, want_f32=True)
            queries, q_h, _ = ln(pre, Lr["n2"][0], Lr["n2"][1], 1e-5, want_f32=True, want_h16=True, split=split)
            # (3) MLP (transformer.py:178-182)
            _, h = Lr["lin1"](q_h, act=ACT_RELU, want_h16=True)
            pre, _ = Lr["lin2"](h, residual=queries, want_f32=True)
            queries, q_h, q_pe_h = ln(pre, Lr["n3"][0], Lr["n3"][1], 1e-5, want_f32=True, want_h16=True, split=split,
                                      pe=tokens, want_out2=True)
            # (4) image -> token cross attention (transformer.py:184-190)
            ia = Lr["i2t"]
            kt, _ = ia.k(q_pe_h, want_f32=True)
            vt, _ = ia.v(q_h, want_f32=True)
            if self.fused_i2t:
                # q_proj, the 7-key softmax, out_proj, the residual and norm4 in ONE kernel: reads the keys once,
                # writes the new keys once (as the h16 pair that is both next operand and next residual)
                b1, b2 = ops.dec_fold_i2t(kt.view(P, 7, 128), vt.view(P, 7, 128), Lr["i2t_wq"], Lr["i2t_wo"], ia.o.b)
                keys_h = ops.dec_i2t_layer(I["keys0_h"] if li == 0 else keys_h, li == 0, Lr["peq_h"], b1, b2, P,
                                           None, Lr["n4"][0], Lr["n4"][1], 1e-5)     # out_proj.bias is folded into b2
                keys_f32 = None
                continue
            qi = I["q0"] if li == 0 else qi1
            _, a = ops.attn_few_keys(qi, kt.view(P, 7, 128), vt.view(P, 7, 128), P, 4096, 7, 8, 16,
                                     want_h16=True, split=split)
            # out_proj + residual + norm4 in one GEMM epilogue.  The new keys are stored once, as an h16 pair:
            # it is the operand of the next projections / upscaling AND (hi + lo = fp32-accurate) the next
            # layer's residual, so no fp32 copy of the [P,4096,256] stream is written in split mode.
            new_h = H16.empty((P * 4096, 256), split, self.dev)
            nxt = None
            if li == 0 and not split:
                nxt = torch.empty((P * 4096, 256), dtype=torch.float32, device=self.dev)
            if li == 0:
                res_kw = dict(residual=I["keys0"], res_mod=4096)
            elif split:
                res_kw = dict(residual_h16=keys_h)
            else:
                res_kw = dict(residual=keys_f32)
            ops.gemm(a.view(P * 4096, 128), ia.o.w, bias=ia.o.b, epi=1, gamma=Lr["n4"][0], beta=Lr["n4"][1], eps=1e-5,
                     out_f32=nxt, out_h16=new_h, **res_kw)
            keys_h, keys_f32 = new_h, nxt
        # final token -> image attention (transformer.py:104-112)
        fa = self.final
        qc, _ = fa.q(q_pe_h, want_f32=True)
        if self.fused_t2i:
            F = self.t2i_fold[2]
            b1t = ops.dec_fold_t2i(qc.view(P, 7, 128), F["wk"])
            _, a = ops.dec_t2i(keys_h, False, F["pek_h"], b1t, P, F["wv_t"], F["bv"])
        else:
            kv, _ = ops.gemm(keys_h, self.w_kvf, residual=self.res_kvf, res_mod=4096, want_f32=True)
            kv = kv.view(P, 4096, 256)
            kc, vc = kv[:, :, 0:128], kv[:, :, 128:256]
            _, a = ops.attn_few_queries(qc.view(P, 7, 128), kc, vc, P, 7, 4096, 8, 16, want_h16=True, split=split)
            del kc, vc, kv
        pre, _ = fa.o(a.view(T, 128), residual=queries, want_f32=True)
        hs, hs_h, _ = ln(pre, self.nf[0], self.nf[1], 1e-5, want_f32=True, want_h16=True, split=split)
        del keys_f32, pre
        hs2 = hs_h.view(P, 7 * 256)

        def cols(h: H16, c0: int, c1: int) -> H16:
            return H16(h.hi[:, c0:c1], None if h.lo is None else h.lo[:, c0:c1])

        # hypernetwork MLPs first: their output is consumed by the fused upscaling epilogue
        hyper = torch.empty((P, 4, 32), dtype=torch.float32, device=self.dev)
        for l in range(4):
            self.hyper[l](cols(hs2, (1 + l) * 256, (2 + l) * 256), out_f32=hyper[:, l, :])
        # upscaling (mask_decoder.py:172-181): ConvT1 (+LN2d+GELU, pixel shuffle) and ConvT2 (+GELU + hypernet
        # dot) are GEMMs with fused epilogues; the [P,32,256,256] upscaled embedding never exists in HBM
        up1 = H16.empty((P * 16384, 64), split, self.dev)
        ops.gemm(keys_h, self.ct1.w, bias=self.ct1.b, epi=2, gamma=self.up_ln[0], beta=self.up_ln[1], eps=1e-6,
                 out_h16=up1)
        masks = torch.empty((P, 4, 256, 256), dtype=torch.float32, device=self.dev)
        ops.gemm(up1, self.ct2.w, bias=self.ct2.b, epi=3, hyper=hyper, masks=masks)
        del up1
        # IoU head (mask_decoder.py:184)
        iou = self.iou_head(cols(hs2, 0, 256))                                         # [P,4]
        # PWD-Net (mask_decoder.py:187-198): softmax-weighted pooling of the DINO map as one GEMM
        if I["dmap_h"] is None:
            cls = torch.zeros((P, 4, self.n_class), dtype=torch.float32, device=self.dev)   # no DINOv2 features bound
        else:
            e, inv = ops.softmax_weights(masks.view(4 * P, 65536), split)
            _, pooled = ops.gemm(e, I["dmap_h"], row_scale=inv, want_h16=True)
            del e
            cls = self.point_cls(pooled).view(P, 4, self.n_class)
        # parallel IoU head on cat(iou_token, mask_token) (mask_decoder.py:194-198); the residual add
        # iou_pred + res_iou_pred is the epilogue of its last GEMM
        m_iou, m_tok = self._gather_maps(P)
        fused = H16.empty((4 * P, 512), split, self.dev)
        ln(hs, normalize=False, row_map=m_iou, out_h16=cols(fused, 0, 256))
        ln(hs, normalize=False, row_map=m_tok, out_h16=cols(fused, 256, 512))
        iou_out = self.par_iou(fused, residual=iou.view(4 * P, 1)).view(P, 4)
        return masks, iou_out, cls
