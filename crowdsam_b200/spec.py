"""Parameter specification (names and shapes) of the models on the hot path.

The names are the reference's state_dict keys, so reference checkpoints
(`sam_vit_l_0b3195.pth`, the adapter file written by tools/train.py:312,
`dinov2_vitl14_pretrain.pth`) load unchanged.  Architectures: build_sam.py:14-45.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

Shape = Tuple[int, ...]

SAM_ARCHS = {
    "vit_b": (768, 12, 12, (2, 5, 8, 11)),
    "vit_l": (1024, 24, 16, (5, 11, 17, 23)),
    "vit_h": (1280, 32, 16, (7, 15, 23, 31)),
}
DINO_ARCHS = {"dinov2_vitl14": (1024, 24, 16), "dinov2_vitb14": (768, 12, 12), "dinov2_vits14": (384, 12, 6)}


def _lin(d: Dict[str, Shape], name: str, out_f: int, in_f: int):
    d[name + ".weight"] = (out_f, in_f)
    d[name + ".bias"] = (out_f,)


def _norm(d, name, n):
    d[name + ".weight"] = (n,)
    d[name + ".bias"] = (n,)


def image_encoder_spec(D: int, depth: int, heads: int, global_idx: Sequence[int]) -> Dict[str, Shape]:
    hd = D // heads
    d: Dict[str, Shape] = {"pos_embed": (1, 64, 64, D), "patch_embed.proj.weight": (D, 3, 16, 16),
                           "patch_embed.proj.bias": (D,)}
    for i in range(depth):
        S = 64 if i in global_idx else 14
        b = f"blocks.{i}"
        _norm(d, b + ".norm1", D)
        _lin(d, b + ".attn.qkv", 3 * D, D)
        _lin(d, b + ".attn.proj", D, D)
        d[b + ".attn.rel_pos_h"] = (2 * S - 1, hd)
        d[b + ".attn.rel_pos_w"] = (2 * S - 1, hd)
        _norm(d, b + ".norm2", D)
        _lin(d, b + ".mlp.lin1", 4 * D, D)
        _lin(d, b + ".mlp.lin2", D, 4 * D)
    d["neck.0.weight"] = (256, D, 1, 1)
    _norm(d, "neck.1", 256)
    d["neck.2.weight"] = (256, 256, 3, 3)
    _norm(d, "neck.3", 256)
    return d


def prompt_encoder_spec() -> Tuple[Dict[str, Shape], Dict[str, Shape]]:
    """(parameters, buffers)."""
    d: Dict[str, Shape] = {}
    for i in range(4):
        d[f"point_embeddings.{i}.weight"] = (1, 256)
    d["not_a_point_embed.weight"] = (1, 256)
    d["mask_downscaling.0.weight"] = (4, 1, 2, 2)
    d["mask_downscaling.0.bias"] = (4,)
    _norm(d, "mask_downscaling.1", 4)
    d["mask_downscaling.3.weight"] = (16, 4, 2, 2)
    d["mask_downscaling.3.bias"] = (16,)
    _norm(d, "mask_downscaling.4", 16)
    d["mask_downscaling.6.weight"] = (256, 16, 1, 1)
    d["mask_downscaling.6.bias"] = (256,)
    d["no_mask_embed.weight"] = (1, 256)
    return d, {"pe_layer.positional_encoding_gaussian_matrix": (2, 128)}


def _attn(d, name, dim, internal):
    for p in ("q_proj", "k_proj", "v_proj"):
        _lin(d, f"{name}.{p}", internal, dim)
    _lin(d, f"{name}.out_proj", dim, internal)


def _mlp(d, name, dims):
    for i in range(len(dims) - 1):
        _lin(d, f"{name}.layers.{i}", dims[i + 1], dims[i])


def mask_decoder_spec(n_class: int = 1) -> Dict[str, Shape]:
    d: Dict[str, Shape] = {}
    for i in range(2):
        L = f"transformer.layers.{i}"
        _attn(d, L + ".self_attn", 256, 256)
        _norm(d, L + ".norm1", 256)
        _attn(d, L + ".cross_attn_token_to_image", 256, 128)
        _norm(d, L + ".norm2", 256)
        _lin(d, L + ".mlp.lin1", 2048, 256)
        _lin(d, L + ".mlp.lin2", 256, 2048)
        _norm(d, L + ".norm3", 256)
        _norm(d, L + ".norm4", 256)
        _attn(d, L + ".cross_attn_image_to_token", 256, 128)
    _attn(d, "transformer.final_attn_token_to_image", 256, 128)
    _norm(d, "transformer.norm_final_attn", 256)
    d["iou_token.weight"] = (1, 256)
    d["mask_tokens.weight"] = (4, 256)
    d["output_upscaling.0.weight"] = (256, 64, 2, 2)
    d["output_upscaling.0.bias"] = (64,)
    _norm(d, "output_upscaling.1", 64)
    d["output_upscaling.3.weight"] = (64, 32, 2, 2)
    d["output_upscaling.3.bias"] = (32,)
    for i in range(5):   # 5 allocated, 4 used (kept for checkpoint compatibility)
        _mlp(d, f"output_hypernetworks_mlps.{i}", (256, 256, 256, 32))
    _mlp(d, "iou_prediction_head", (256, 256, 256, 4))
    _lin(d, "dino_proj", 256, 1024)
    _mlp(d, "parallel_iou_head", (512, 256, 256, 1))
    _mlp(d, "point_classifier", (256, 256, n_class))
    return d


def dino_spec(D: int, depth: int) -> Dict[str, Shape]:
    d: Dict[str, Shape] = {"cls_token": (1, 1, D), "pos_embed": (1, 37 * 37 + 1, D), "mask_token": (1, D),
                           "patch_embed.proj.weight": (D, 3, 14, 14), "patch_embed.proj.bias": (D,)}
    for i in range(depth):
        b = f"blocks.{i}"
        _norm(d, b + ".norm1", D)
        _lin(d, b + ".attn.qkv", 3 * D, D)
        _lin(d, b + ".attn.proj", D, D)
        d[b + ".ls1.gamma"] = (D,)
        _norm(d, b + ".norm2", D)
        _lin(d, b + ".mlp.fc1", 4 * D, D)
        _lin(d, b + ".mlp.fc2", D, 4 * D)
        d[b + ".ls2.gamma"] = (D,)
    _norm(d, "norm", D)
    return d
