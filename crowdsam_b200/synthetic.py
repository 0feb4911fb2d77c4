"""Synthetic workload generator: weights ("recipe v1 / v2"), images and the reference's
default test configuration.  Pure data, no compute path: bench.py, smoke() and the tests build their
inputs from here (no network, so no checkpoint or dataset can be fetched), and oracle/weights.py re-exports
it so that the real reference, the oracle restatement and the CUDA path all load the SAME state_dict.

No checkpoint ships with the reference, so every parity / bench run uses the
same deterministic random state_dict on all sides.  Key names and shapes follow the reference modules so
`load_state_dict(strict=True)` succeeds on them
(segment_anything_cs/build_sam.py:104-158, modeling/*.py,
dinov2/dinov2/models/vision_transformer.py:60-177).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

SAM_ARCHS = {
    # name: (embed_dim, depth, heads, global attention block indexes)  build_sam.py:14-45
    "vit_b": (768, 12, 12, (2, 5, 8, 11)),
    "vit_l": (1024, 24, 16, (5, 11, 17, 23)),
    "vit_h": (1280, 32, 16, (7, 15, 23, 31)),
    # reduced-depth variants with full-size per-layer shapes, for fast CPU tests
    "tiny": (128, 2, 2, (1,)),
    "tiny_l": (1024, 2, 16, (1,)),
}

DINO_ARCHS = {
    # name: (embed_dim, depth, heads)  vision_transformer.py:340-379
    "dinov2_vitl14": (1024, 24, 16),
    "tiny": (1024, 2, 16),
}


def _lin(g: torch.Generator, out_f: int, in_f: int, bias: bool = True):
    """nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both."""
    k = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * k
    b = (torch.rand(out_f, generator=g) * 2 - 1) * k if bias else None
    return w, b


def _put_lin(sd, g, name, out_f, in_f, bias=True):
    w, b = _lin(g, out_f, in_f, bias)
    sd[name + ".weight"] = w
    if bias:
        sd[name + ".bias"] = b


def _put_ln(sd, g, name, dim):
    # affine parameters away from (1, 0) so the kernels' gamma/beta paths are exercised
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
    sd[name + ".bias"] = 0.1 * torch.randn(dim, generator=g)


def _put_attn(sd, g, name, dim, internal):
    for p in ("q_proj", "k_proj", "v_proj"):
        _put_lin(sd, g, f"{name}.{p}", internal, dim)
    _put_lin(sd, g, f"{name}.out_proj", dim, internal)


def _put_mlp(sd, g, name, dims: Sequence[int]):
    for i in range(len(dims) - 1):
        _put_lin(sd, g, f"{name}.layers.{i}", dims[i + 1], dims[i])


def make_sam_state(arch: str = "vit_l", n_class: int = 1, seed: int = 0,
                   recipe_v1: bool = True, recipe: str = "") -> Dict[str, torch.Tensor]:
    """State dict for `Sam` (segment_anything_cs/modeling/sam.py:16-47).
    recipe "v1" = SURVEY.md §8d (used for the reduced-depth golden archs); "v2" (default for the full-depth
    archs) additionally zeroes the last hypernetwork bias and centres the classifier logit, because with
    24-layer encoders v1 gives masks with a large negative mean (stability 0.60, 3% pass) and classifier
    logits around -6 (scores < 0.1), i.e. nothing reaches NMS [measured with the oracle on ViT-L]."""
    if not recipe:
        recipe = "v1" if arch.startswith("tiny") else "v2"
    D, depth, heads, glob = SAM_ARCHS[arch]
    hd = D // heads
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    # ---- image encoder (image_encoder.py:17-104)
    e = "image_encoder"
    sd[f"{e}.pos_embed"] = 0.02 * torch.randn(1, 64, 64, D, generator=g)
    k = 1.0 / math.sqrt(3 * 16 * 16)
    sd[f"{e}.patch_embed.proj.weight"] = (torch.rand(D, 3, 16, 16, generator=g) * 2 - 1) * k
    sd[f"{e}.patch_embed.proj.bias"] = (torch.rand(D, generator=g) * 2 - 1) * k
    for i in range(depth):
        b = f"{e}.blocks.{i}"
        S = 64 if i in glob else 14
        _put_ln(sd, g, f"{b}.norm1", D)
        _put_lin(sd, g, f"{b}.attn.qkv", 3 * D, D)
        _put_lin(sd, g, f"{b}.attn.proj", D, D)
        sd[f"{b}.attn.rel_pos_h"] = 0.02 * torch.randn(2 * S - 1, hd, generator=g)
        sd[f"{b}.attn.rel_pos_w"] = 0.02 * torch.randn(2 * S - 1, hd, generator=g)
        _put_ln(sd, g, f"{b}.norm2", D)
        _put_lin(sd, g, f"{b}.mlp.lin1", 4 * D, D)
        _put_lin(sd, g, f"{b}.mlp.lin2", D, 4 * D)
    sd[f"{e}.neck.0.weight"] = ((torch.rand(256, D, 1, 1, generator=g) * 2 - 1) / math.sqrt(D))
    _put_ln(sd, g, f"{e}.neck.1", 256)
    sd[f"{e}.neck.2.weight"] = ((torch.rand(256, 256, 3, 3, generator=g) * 2 - 1) / math.sqrt(256 * 9))
    _put_ln(sd, g, f"{e}.neck.3", 256)
    # ---- prompt encoder (prompt_encoder.py:16-60)
    p = "prompt_encoder"
    sd[f"{p}.pe_layer.positional_encoding_gaussian_matrix"] = torch.randn(2, 128, generator=g)
    for i in range(4):
        sd[f"{p}.point_embeddings.{i}.weight"] = torch.randn(1, 256, generator=g)
    sd[f"{p}.not_a_point_embed.weight"] = torch.randn(1, 256, generator=g)
    sd[f"{p}.no_mask_embed.weight"] = torch.randn(1, 256, generator=g)
    # mask_downscaling (unused on the point-prompt path, kept for state_dict compatibility)
    for idx, (co, ci, ks) in {0: (4, 1, 2), 3: (16, 4, 2), 6: (256, 16, 1)}.items():
        k = 1.0 / math.sqrt(ci * ks * ks)
        sd[f"{p}.mask_downscaling.{idx}.weight"] = (torch.rand(co, ci, ks, ks, generator=g) * 2 - 1) * k
        sd[f"{p}.mask_downscaling.{idx}.bias"] = (torch.rand(co, generator=g) * 2 - 1) * k
    _put_ln(sd, g, f"{p}.mask_downscaling.1", 4)
    _put_ln(sd, g, f"{p}.mask_downscaling.4", 16)
    # ---- mask decoder (mask_decoder.py:19-75, transformer.py:16-61,117-158)
    m = "mask_decoder"
    t = f"{m}.transformer"
    for i in range(2):
        L = f"{t}.layers.{i}"
        _put_attn(sd, g, f"{L}.self_attn", 256, 256)
        _put_ln(sd, g, f"{L}.norm1", 256)
        _put_attn(sd, g, f"{L}.cross_attn_token_to_image", 256, 128)
        _put_ln(sd, g, f"{L}.norm2", 256)
        _put_lin(sd, g, f"{L}.mlp.lin1", 2048, 256)
        _put_lin(sd, g, f"{L}.mlp.lin2", 256, 2048)
        _put_ln(sd, g, f"{L}.norm3", 256)
        _put_ln(sd, g, f"{L}.norm4", 256)
        _put_attn(sd, g, f"{L}.cross_attn_image_to_token", 256, 128)
    _put_attn(sd, g, f"{t}.final_attn_token_to_image", 256, 128)
    _put_ln(sd, g, f"{t}.norm_final_attn", 256)
    sd[f"{m}.iou_token.weight"] = torch.randn(1, 256, generator=g)
    sd[f"{m}.mask_tokens.weight"] = torch.randn(4, 256, generator=g)
    # ConvTranspose2d weight layout is [in, out, kh, kw]
    k = 1.0 / math.sqrt(64 * 4)
    sd[f"{m}.output_upscaling.0.weight"] = (torch.rand(256, 64, 2, 2, generator=g) * 2 - 1) * k
    sd[f"{m}.output_upscaling.0.bias"] = (torch.rand(64, generator=g) * 2 - 1) * k
    _put_ln(sd, g, f"{m}.output_upscaling.1", 64)
    k = 1.0 / math.sqrt(32 * 4)
    sd[f"{m}.output_upscaling.3.weight"] = (torch.rand(64, 32, 2, 2, generator=g) * 2 - 1) * k
    sd[f"{m}.output_upscaling.3.bias"] = (torch.rand(32, generator=g) * 2 - 1) * k
    for i in range(5):  # 5 allocated, 4 used (mask_decoder.py:63-68,177-178)
        _put_mlp(sd, g, f"{m}.output_hypernetworks_mlps.{i}", (256, 256, 256, 32))
    _put_mlp(sd, g, f"{m}.iou_prediction_head", (256, 256, 256, 4))
    _put_lin(sd, g, f"{m}.dino_proj", 256, 1024)
    _put_mlp(sd, g, f"{m}.parallel_iou_head", (512, 256, 256, 1))
    _put_mlp(sd, g, f"{m}.point_classifier", (256, 256, n_class))
    if recipe_v1:
        # SURVEY.md §8d (ii),(iii): spread logits / scores so filters see both outcomes
        for i in range(4):
            sd[f"{m}.output_hypernetworks_mlps.{i}.layers.2.weight"] *= 200.0
            sd[f"{m}.output_hypernetworks_mlps.{i}.layers.2.bias"] *= 200.0
        sd[f"{m}.iou_prediction_head.layers.2.weight"] *= 10.0
        sd[f"{m}.iou_prediction_head.layers.2.bias"] = torch.full((4,), 0.5)
        sd[f"{m}.point_classifier.layers.1.weight"] *= 50.0
        if recipe == "v2":
            for i in range(4):
                sd[f"{m}.output_hypernetworks_mlps.{i}.layers.2.bias"].zero_()
            sd[f"{m}.point_classifier.layers.1.bias"] = torch.full((n_class,), 6.0)
    return sd


def make_dino_state(arch: str = "dinov2_vitl14", seed: int = 1) -> Dict[str, torch.Tensor]:
    """State dict for DINOv2 `DinoVisionTransformer` built by the hub entry
    (hub/backbones.py:18-62: img_size 518, patch 14, init_values 1.0, block_chunks 0)."""
    D, depth, heads = DINO_ARCHS[arch]
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["cls_token"] = 0.02 * torch.randn(1, 1, D, generator=g)
    sd["pos_embed"] = 0.02 * torch.randn(1, 37 * 37 + 1, D, generator=g)
    sd["mask_token"] = torch.zeros(1, D)
    k = 1.0 / math.sqrt(3 * 14 * 14)
    sd["patch_embed.proj.weight"] = (torch.rand(D, 3, 14, 14, generator=g) * 2 - 1) * k
    sd["patch_embed.proj.bias"] = (torch.rand(D, generator=g) * 2 - 1) * k

    def tn(*shape):
        return (0.02 * torch.randn(*shape, generator=g)).clamp_(-0.04, 0.04)

    for i in range(depth):
        b = f"blocks.{i}"
        _put_ln(sd, g, f"{b}.norm1", D)
        sd[f"{b}.attn.qkv.weight"] = tn(3 * D, D)
        sd[f"{b}.attn.qkv.bias"] = 0.02 * torch.randn(3 * D, generator=g)
        sd[f"{b}.attn.proj.weight"] = tn(D, D)
        sd[f"{b}.attn.proj.bias"] = 0.02 * torch.randn(D, generator=g)
        sd[f"{b}.ls1.gamma"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        _put_ln(sd, g, f"{b}.norm2", D)
        sd[f"{b}.mlp.fc1.weight"] = tn(4 * D, D)
        sd[f"{b}.mlp.fc1.bias"] = 0.02 * torch.randn(4 * D, generator=g)
        sd[f"{b}.mlp.fc2.weight"] = tn(D, 4 * D)
        sd[f"{b}.mlp.fc2.bias"] = 0.02 * torch.randn(D, generator=g)
        sd[f"{b}.ls2.gamma"] = 1.0 + 0.1 * torch.randn(D, generator=g)
    _put_ln(sd, g, "norm", D)
    return sd


def synthetic_image(index: int, h: int = 1024, w: int = 1024):
    """uniform-random uint8 HWC image, seed = image index."""
    import numpy as np

    return np.random.default_rng(index).integers(0, 256, (h, w, 3), dtype=np.uint8)


# configs/crowdhuman.yaml:34-60 (`test:` section of the reference's only shipped config)
DEFAULT_TEST_CFG = dict(
    crop_n_layers=0, crop_nms_thresh=0.7, crop_overlap_ratio=0.341, pos_sim_thresh=0.5,
    grid_size=192, max_prompts=500, filter_thresh=0.7, points_per_batch=32,
    mask_selection="max_iou", max_size=1024, min_mask_region_area=100, box_nms_thresh=0.65,
    stability_score_thresh=0.8, stability_score_offset=1, pred_iou_thresh=0.1)


def grid_points(G: int, size: int = 1024):
    """Prompt grid produced by the reference config overrides of SURVEY.md §8d: pixel (j*size/G, i*size/G) for every
    grid cell, row-major (model.py:200-223)."""
    import numpy as np

    ii, jj = np.meshgrid(np.arange(G), np.arange(G), indexing="ij")
    return np.stack([jj.reshape(-1) * (size / G), ii.reshape(-1) * (size / G)], axis=1).astype(int)
