"""nn.Module shells that mirror the reference's module tree (same attribute paths, same
state_dict keys) while computing through the libcsam_sm100 engines.

Reference surface mirrored: `Sam` (modeling/sam.py:16-52,132-173),
`ImageEncoderViT.forward/img_size` (image_encoder.py:106-116), `PromptEncoder.forward/
get_dense_pe` (prompt_encoder.py:64-73,130-172), `MaskDecoder.forward` returning 3 tensors
with `dino_proj / point_classifier / parallel_iou_head` sub-modules (mask_decoder.py:72-74,
92-137), DINOv2 `forward_features(x)['x_norm_patchtokens']` (vision_transformer.py:254-270).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn

from . import engine as E
from . import ops, spec


class ParamTree(nn.Module):
    """A module whose parameter tree is generated from {dotted name: shape}; intermediate names
    become child ParamTree modules so attribute paths like `mask_decoder.dino_proj` exist and
    have `.parameters()` (tools/train.py:294-303)."""

    def __init__(self, shapes: Optional[Dict[str, Tuple[int, ...]]] = None, buffers=None):
        super().__init__()
        for name, shape in (shapes or {}).items():
            self._add(name.split("."), shape, False)
        for name, shape in (buffers or {}).items():
            self._add(name.split("."), shape, True)

    def _add(self, parts, shape, is_buffer):
        node = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamTree())
            node = node._modules[p]
        t = torch.zeros(*shape)
        if is_buffer:
            node.register_buffer(parts[-1], t)
        else:
            node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))

    def __getitem__(self, i):   # ModuleList-style access: layers[0], point_embeddings[1]
        return self._modules[str(i)]


def _default_init_(module: nn.Module, seed: int = 0):
    """Deterministic default initialisation (fan-in uniform for matrices, ones/zeros for norms)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in list(module.named_parameters()) + list(module.named_buffers()):
        leaf = name.split(".")[-1]
        is_norm = any(k in name for k in ("norm", "neck.1", "neck.3", "output_upscaling.1", "mask_downscaling.1",
                                          "mask_downscaling.4"))
        with torch.no_grad():
            if "gaussian_matrix" in name:
                p.copy_(torch.randn(p.shape, generator=g))
            elif leaf == "gamma" or (is_norm and leaf == "weight"):
                p.fill_(1.0)
            elif is_norm and leaf == "bias":
                p.zero_()
            elif leaf in ("pos_embed", "rel_pos_h", "rel_pos_w", "mask_token"):
                p.zero_()
            elif leaf == "cls_token":
                p.copy_(1e-6 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2 and "embed" not in name and "token" not in name:
                fan_in = p[0].numel() if "output_upscaling" not in name else p.shape[1] * p.shape[2] * p.shape[3]
                k = 1.0 / math.sqrt(max(fan_in, 1))
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * k)
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.05)


class _Engined(nn.Module):
    """Keeps a lazily-built engine in sync with the parameters (rebuilt after load_state_dict / .to)."""

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None

    def _params(self):
        # nn.Module.parameters() walks the module tree through several Python generators: ~1 ms for a ViT-L, five times
        # per image, all of it BEFORE the first kernel of the image is launched (measured: 3 ms of the gap between the
        # end-to-end and the device-resident leg).  The flat list is cached and dropped wherever the tree can change.
        pl = self.__dict__.get("_plist")
        if pl is None:
            pl = list(self.parameters())
            self.__dict__["_plist"] = pl
        return pl

    def _state_key(self):
        pl = self._params()
        return (pl[0].device, sum(q._version for q in pl), E.default_split())

    def load_state_dict(self, *a, **k):
        self._engine = None
        self.__dict__["_plist"] = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        self.__dict__["_plist"] = None
        return super()._apply(fn, *a, **k)

    def register_parameter(self, name, param):
        self.__dict__["_plist"] = None
        return super().register_parameter(name, param)

    def add_module(self, name, module):
        self.__dict__["_plist"] = None
        return super().add_module(name, module)

    def _flat(self) -> Dict[str, torch.Tensor]:
        sd = {k: v for k, v in self.named_parameters()}
        sd.update({k: v for k, v in self.named_buffers()})
        return sd

    def engine(self):
        key = self._state_key()
        if self._engine is None or self._engine_key != key:
            if key[0].type != "cuda":
                raise RuntimeError("crowdsam_b200 computes on CUDA only (no CPU fallback): move the model to a B200")
            self._engine = self._build(key[0])
            self._engine_key = key
        return self._engine


class ImageEncoderViT(_Engined):
    def __init__(self, embed_dim: int, depth: int, num_heads: int, global_attn_indexes: Sequence[int],
                 img_size: int = 1024):
        super().__init__()
        assert img_size == 1024
        self.img_size = img_size
        self.depth, self.num_heads, self.global_attn_indexes = depth, num_heads, tuple(global_attn_indexes)
        tree = ParamTree(spec.image_encoder_spec(embed_dim, depth, num_heads, global_attn_indexes))
        for n, m in tree._modules.items():
            self.add_module(n, m)
        for n, p in tree._parameters.items():
            self.register_parameter(n, p)

    def _build(self, dev):
        sd = {"image_encoder." + k: v for k, v in self._flat().items()}
        return E.SamEncoder(sd, self.depth, self.num_heads, self.global_attn_indexes, dev, E.default_split())

    @torch.no_grad()
    def forward_u8(self, img_u8_chw: torch.Tensor):
        """uint8 [3,h,w] -> (features [1,256,64,64], token-major [4096,256]); normalisation and padding
        (Sam.preprocess) are fused into the patch gather."""
        return self.engine().forward(img_u8_chw)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "ImageEncoderViT.forward(float image) is not part of the B200 path; use SamPredictor.set_image / "
            "forward_u8 (preprocess is fused into the encoder's first kernel)")


class PromptEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed_dim = 256
        self.input_image_size = (1024, 1024)
        self.image_embedding_size = (64, 64)
        params, bufs = spec.prompt_encoder_spec()
        tree = ParamTree(params, bufs)
        for n, m in tree._modules.items():
            self.add_module(n, m)
        self._owner = None   # set by Sam: the decoder engine owns the fused token kernel

    def get_dense_pe(self) -> torch.Tensor:
        return self._owner[0].mask_decoder.engine().dense_pe

    def forward(self, points, boxes, masks):
        """Returns (sparse [B,n+1,256], dense view [B,256,64,64]) for point prompts
        (prompt_encoder.py:130-172).  Boxes / mask inputs are outside the hot path."""
        if boxes is not None or masks is not None or points is None:
            raise NotImplementedError("only point prompts are on the B200 hot path")
        coords, labels = points
        if coords.shape[1] != 1:
            raise NotImplementedError("one point per prompt on the B200 hot path")
        eng = self._owner[0].mask_decoder.engine()
        c01 = ((coords[:, 0, :].double() + 0.5) / 1024.0).float().contiguous().to(eng.dev)
        lab = labels[:, 0].to(torch.int32).contiguous().to(eng.dev)
        tokens = ops.prompt_tokens(c01, lab, eng.gauss, eng.tok5, eng.point_emb, eng.nap)
        dense = eng.no_mask.reshape(1, -1, 1, 1).expand(coords.shape[0], -1, 64, 64)
        return tokens[:, 5:, :], dense


class MaskDecoder(_Engined):
    def __init__(self, n_class: int = 1):
        super().__init__()
        self.transformer_dim = 256
        self.num_multimask_outputs = 3
        self.num_mask_tokens = 4
        self.n_class = n_class
        tree = ParamTree(spec.mask_decoder_spec(n_class))
        for n, m in tree._modules.items():
            self.add_module(n, m)
        self._owner = None

    def _flat(self):
        sd = {"mask_decoder." + k: v for k, v in super()._flat().items()}
        pe = self._owner[0].prompt_encoder
        sd.update({"prompt_encoder." + k: v for k, v in list(pe.named_parameters()) + list(pe.named_buffers())})
        return sd

    def _state_key(self):
        pe = self._owner[0].prompt_encoder
        base = super()._state_key()
        pl = self.__dict__.get("_pe_plist")
        if pl is None or self.__dict__.get("_pe_owner") is not pe:
            pl = list(pe.parameters())
            self.__dict__["_pe_plist"], self.__dict__["_pe_owner"] = pl, pe
        return base + (sum(q._version for q in pl),)

    def _build(self, dev):
        return E.MaskDecoderEngine(self._flat(), dev, E.default_split())

    def forward(self, image_embeddings=None, image_pe=None, sparse_prompt_embeddings=None,
                dense_prompt_embeddings=None, multimask_output=True, attn_sim=None, target_embedding=None,
                dino_feats=None):
        raise NotImplementedError(
            "MaskDecoder.forward on raw embeddings is served through SamPredictor.predict_torch on the B200 path "
            "(the decoder engine consumes prompt coordinates directly and hoists prompt-independent work)")


class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder: ImageEncoderViT, prompt_encoder: PromptEncoder, mask_decoder: MaskDecoder):
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1), False)
        # non-module back references (a list hides them from nn.Module registration)
        prompt_encoder._owner = [self]
        mask_decoder._owner = [self]

    @property
    def device(self):
        return self.pixel_mean.device

    def load_state_dict(self, *a, **k):
        self.image_encoder._engine = None
        self.mask_decoder._engine = None
        return super().load_state_dict(*a, **k)

    def forward(self, *a, **k):
        raise NotImplementedError("Sam.forward is broken in the reference fork (sam.py:110-116) and unused; "
                                  "use SamPredictor")


def build_sam_model(embed_dim, depth, heads, n_class, global_idx, checkpoint=None) -> Sam:
    sam = Sam(ImageEncoderViT(embed_dim, depth, heads, global_idx), PromptEncoder(), MaskDecoder(n_class))
    _default_init_(sam)
    sam.eval()
    if checkpoint is not None:
        with open(checkpoint, "rb") as f:
            state = torch.load(f, map_location="cpu")
        sam.load_state_dict(state, strict=False)    # build_sam.py:154-157
    return sam


class DinoVisionTransformer(_Engined):
    """DINOv2 backbone shell: parameters + forward_features through the engine."""

    def __init__(self, embed_dim=1024, depth=24, num_heads=16):
        super().__init__()
        self.embed_dim, self.depth, self.num_heads, self.patch_size = embed_dim, depth, num_heads, 14
        tree = ParamTree(spec.dino_spec(embed_dim, depth))
        for n, m in tree._modules.items():
            self.add_module(n, m)
        for n, p in tree._parameters.items():
            self.register_parameter(n, p)
        _default_init_(self, seed=1)

    def _build(self, dev):
        return E.DinoEncoder(self._flat(), self.depth, self.num_heads, dev, E.default_split())

    @torch.no_grad()
    def forward_features_u8(self, img_u8_chw: torch.Tensor):
        return self.engine().forward(img_u8_chw)

    def forward_features(self, x, masks=None):
        raise NotImplementedError("DINOv2 consumes the uint8 image through SamPredictor.set_image on the B200 path")
