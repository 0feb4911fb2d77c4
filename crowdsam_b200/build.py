"""Model registry with the reference's call convention (segment_anything_cs/build_sam.py:95-101):
`sam_model_registry[name](checkpoint=path_or_None, n_class=int)`.

The reference's `vit_b` / `vit_h` / `default` entries raise TypeError (they omit `n_class`) and `vit_t`
raises NameError; here every ViT entry accepts `n_class=1` and `vit_t` raises a
clear NotImplementedError.
"""
from __future__ import annotations

from .modules import Sam, build_sam_model
from .spec import SAM_ARCHS


def _build_sam(encoder_embed_dim, encoder_depth, encoder_num_heads, n_class, encoder_global_attn_indexes,
               checkpoint=None) -> Sam:
    return build_sam_model(encoder_embed_dim, encoder_depth, encoder_num_heads, n_class,
                           tuple(encoder_global_attn_indexes), checkpoint)


def _entry(name):
    D, depth, heads, glob = SAM_ARCHS[name]

    def build(checkpoint=None, n_class=1):
        return _build_sam(D, depth, heads, n_class, glob, checkpoint)

    build.__name__ = f"build_sam_{name}"
    return build


build_sam_vit_h = _entry("vit_h")
build_sam_vit_l = _entry("vit_l")
build_sam_vit_b = _entry("vit_b")
build_sam = build_sam_vit_h


def build_sam_vit_t(checkpoint=None, n_class=1):
    raise NotImplementedError("vit_t (MobileSAM TinyViT) is not defined in the reference either (build_sam.py:47-93)")


sam_model_registry = {
    "default": build_sam_vit_h,
    "vit_h": build_sam_vit_h,
    "vit_l": build_sam_vit_l,
    "vit_b": build_sam_vit_b,
    "vit_t": build_sam_vit_t,
}
