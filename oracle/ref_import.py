"""Import the REAL reference (/root/reference) on CPU with the three shims of SURVEY.md §8c.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); the GPU box does not have it, so nothing under `-m gpu`, smoke() or
bench.py may call this.  Used by tests/golden/make_golden.py to produce the
committed golden vectors and by the `-m "not gpu"` live cross-check.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("CROWDSAM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "segment_anything_cs"))


def _install_shims():
    import torch

    # shim 1: matplotlib is imported at crowdsam/utils.py:17 but only used for drawing
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    # shim 2: hard-coded .cuda() at predictor.py:105 on a CPU-only host
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # shim 3: pycocotools.mask.frPyObjects (amg.py:294-300) -> our restatement of rleToString
    if "pycocotools" not in sys.modules:
        try:
            import pycocotools  # noqa: F401
        except ImportError:
            from . import restate

            pkg = types.ModuleType("pycocotools")
            msk = types.ModuleType("pycocotools.mask")

            def frPyObjects(rle, h, w):
                return {"size": [h, w], "counts": restate.coco_rle_string(rle["counts"]).encode("utf-8")}

            msk.frPyObjects = frPyObjects
            pkg.mask = msk
            sys.modules["pycocotools"] = pkg
            sys.modules["pycocotools.mask"] = msk


def load():
    """Returns the reference modules (segment_anything_cs, crowdsam.model, crowdsam.utils)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import segment_anything_cs as sacs
    import crowdsam.model as cmodel
    import crowdsam.utils as cutils

    assert sacs.__file__.startswith(REF_ROOT), sacs.__file__
    return sacs, cmodel, cutils


def build_sam(sam_state, arch):
    """Real reference `Sam` with our synthetic state_dict loaded strictly.
    Uses _build_sam directly: the vit_b / vit_h registry entries raise TypeError
    (build_sam.py:14-21,38-45)."""
    from .weights import SAM_ARCHS

    load()
    from segment_anything_cs.build_sam import _build_sam

    D, depth, heads, glob = SAM_ARCHS[arch]
    sam = _build_sam(D, depth, heads, 1, list(glob))
    missing = sam.load_state_dict(sam_state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return sam.eval()


def build_dino(dino_state, arch):
    """Real DINOv2 ViT built like hub/backbones.py:18-62 (img 518, patch 14, init_values 1.0,
    block_chunks 0) at the requested depth."""
    from .weights import DINO_ARCHS

    load()
    dino_root = os.path.join(REF_ROOT, "dinov2")
    if dino_root not in sys.path:
        sys.path.insert(0, dino_root)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from dinov2.models import vision_transformer as vits
        from dinov2.layers import MemEffAttention, NestedTensorBlock as Block
    from functools import partial

    D, depth, heads = DINO_ARCHS[arch]
    model = vits.DinoVisionTransformer(
        img_size=518, patch_size=14, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4,
        block_fn=partial(Block, attn_class=MemEffAttention), init_values=1.0, ffn_layer="mlp",
        block_chunks=0, num_register_tokens=0, interpolate_antialias=False, interpolate_offset=0.1)
    res = model.load_state_dict(dino_state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model


def build_crowdsam(sam, dino, test_cfg):
    """CrowdSAM object without touching checkpoints: __new__ + the attributes __init__ sets
    (crowdsam/model.py:27-64)."""
    import torch

    sacs, cmodel, _ = load()
    m = cmodel.CrowdSAM.__new__(cmodel.CrowdSAM)
    m.device = torch.device("cpu")
    m.train_free = False
    m.predictor = sacs.SamPredictor(sam, dino)
    for k in ("mask_selection", "max_prompts", "filter_thresh", "max_size", "grid_size",
              "pred_iou_thresh", "stability_score_thresh", "stability_score_offset",
              "box_nms_thresh", "points_per_batch", "crop_n_layers", "crop_nms_thresh",
              "crop_overlap_ratio", "min_mask_region_area", "pos_sim_thresh"):
        setattr(m, k, test_cfg[k])
    m.apply_box_offsets = False
    m.fuse_simmap = False
    m.output_rles = True
    return m


def coco_string(rle) -> str:
    """COCO compressed string of an uncompressed RLE dict of the reference (through the pycocotools stand-in)."""
    load()
    from segment_anything_cs.utils.amg import coco_encode_rle

    return coco_encode_rle(dict(rle))["counts"]
