"""Plain PyTorch fp32 (CPU) restatement of the Crowd-SAM inference hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Functional style over a
flat ``state_dict``; every function cites the reference file:line it follows
(paths relative to /root/reference).  Pinned against the real reference by
tests/golden (tests/test_oracle_golden.py).

All arithmetic is fp32 like the reference; only point coordinates pass through
fp64 on the host (transforms.py:42-44, prompt_encoder.py:82,215-218).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

T = torch.Tensor
SD = Dict[str, torch.Tensor]

PIXEL_MEAN = (123.675, 116.28, 103.53)   # build_sam.py:148
PIXEL_STD = (58.395, 57.12, 57.375)      # build_sam.py:149
IMG = 1024                               # build_sam.py:113
WIN = 14                                 # build_sam.py:128


def _lin(sd: SD, name: str, x: T) -> T:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd: SD, name: str, x: T, eps: float) -> T:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


# --------------------------------------------------------------------------------------
# A1  Sam.preprocess                                                    sam.py:163-173
# --------------------------------------------------------------------------------------
def preprocess(img_u8: T) -> T:
    """uint8 [1,3,h,w] -> fp32 [1,3,1024,1024]; normalise then zero-pad right/bottom."""
    mean = torch.tensor(PIXEL_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(PIXEL_STD).view(1, 3, 1, 1)
    x = (img_u8 - mean) / std
    h, w = x.shape[-2:]
    return F.pad(x, (0, IMG - w, 0, IMG - h))


# --------------------------------------------------------------------------------------
# A2  SAM ViT image encoder                                   image_encoder.py:106-395
# --------------------------------------------------------------------------------------
def _rel_table(rel: T, S: int) -> T:
    """get_rel_pos for q_size == k_size == S (image_encoder.py:292-322): R[q,k] = rel[q-k+S-1]."""
    assert rel.shape[0] == 2 * S - 1
    idx = torch.arange(S)[:, None] - torch.arange(S)[None, :] + (S - 1)
    return rel[idx]


def _vit_attention(sd: SD, pre: str, x: T, heads: int) -> T:
    """Attention.forward with decomposed rel-pos (image_encoder.py:224-240,325-361).
    x: [B,S,S,D].  Scores use q*scale; the bias uses the UNSCALED q (:231-234)."""
    B, S, _, D = x.shape
    hd = D // heads
    qkv = _lin(sd, pre + ".qkv", x).reshape(B, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * heads, S * S, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh = _rel_table(sd[pre + ".rel_pos_h"], S)
    Rw = _rel_table(sd[pre + ".rel_pos_w"], S)
    rq = q.reshape(B * heads, S, S, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(B * heads, S, S, S, S) + rel_h[..., :, None] + rel_w[..., None, :]).view(
        B * heads, S * S, S * S)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).view(B, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(B, S, S, D)
    return _lin(sd, pre + ".proj", out)


def _vit_block(sd: SD, pre: str, x: T, heads: int, window: int) -> T:
    """Block.forward (image_encoder.py:166-182) with window_partition/unpartition (:243-289).
    Padding to 70x70 happens AFTER norm1, so padded tokens are exact zeros going into qkv."""
    short = x
    x = _ln(sd, pre + ".norm1", x, 1e-6)
    if window > 0:
        B, H, W, D = x.shape
        ph, pw = (-H) % window, (-W) % window
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        x = x.view(B, Hp // window, window, Wp // window, window, D).permute(0, 1, 3, 2, 4, 5)
        x = x.reshape(-1, window, window, D)
    x = _vit_attention(sd, pre + ".attn", x, heads)
    if window > 0:
        x = x.view(B, Hp // window, Wp // window, window, window, D).permute(0, 1, 3, 2, 4, 5)
        x = x.reshape(B, Hp, Wp, D)[:, :H, :W, :]
    x = short + x
    y = _ln(sd, pre + ".norm2", x, 1e-6)
    y = _lin(sd, pre + ".mlp.lin2", F.gelu(_lin(sd, pre + ".mlp.lin1", y)))   # common.py:25-26
    return x + y


def _ln2d(sd: SD, name: str, x: T, eps: float = 1e-6) -> T:
    """LayerNorm2d over channels, biased variance (common.py:38-43)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[name + ".weight"][:, None, None] * x + sd[name + ".bias"][:, None, None]


def sam_encoder(sd: SD, x: T, depth: int, heads: int, global_idx: Sequence[int],
                pre: str = "image_encoder") -> T:
    """ImageEncoderViT.forward (image_encoder.py:106-116): fp32 [1,3,1024,1024] -> [1,256,64,64]."""
    x = F.conv2d(x, sd[pre + ".patch_embed.proj.weight"], sd[pre + ".patch_embed.proj.bias"], stride=16)
    x = x.permute(0, 2, 3, 1) + sd[pre + ".pos_embed"]
    for i in range(depth):
        x = _vit_block(sd, f"{pre}.blocks.{i}", x, heads, 0 if i in global_idx else WIN)
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[pre + ".neck.0.weight"])
    x = _ln2d(sd, pre + ".neck.1", x)
    x = F.conv2d(x, sd[pre + ".neck.2.weight"], padding=1)
    return _ln2d(sd, pre + ".neck.3", x)


# --------------------------------------------------------------------------------------
# A3  DINOv2 forward_features                 dinov2/models/vision_transformer.py:179-270
# --------------------------------------------------------------------------------------
def dino_pos_embed(sd: SD, n_side: int = 73) -> T:
    """interpolate_pos_encoding (vision_transformer.py:179-211): bicubic 37x37 -> 73x73 via
    scale_factor=(n+0.1)/37, class token embedding kept."""
    pe = sd["pos_embed"].float()
    N = pe.shape[1] - 1
    M = int(math.sqrt(N))
    D = pe.shape[-1]
    s = float(n_side + 0.1) / M
    patch = F.interpolate(pe[:, 1:].reshape(1, M, M, D).permute(0, 3, 1, 2), mode="bicubic",
                          antialias=False, scale_factor=(s, s))
    assert patch.shape[-2:] == (n_side, n_side)
    patch = patch.permute(0, 2, 3, 1).reshape(1, -1, D)
    return torch.cat((pe[:, :1], patch), dim=1)


def dino_forward(sd: SD, x: T, depth: int, heads: int) -> T:
    """forward_features (vision_transformer.py:254-270) -> x_norm_patchtokens [1,N,D].
    Block: x += ls1(attn(norm1 x)); x += ls2(mlp(norm2 x)) (layers/block.py:89-115);
    attention scales q first (layers/attention.py:56-69)."""
    B, _, H, W = x.shape
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=14)
    n_side = t.shape[-1]
    t = t.flatten(2).transpose(1, 2)                                   # patch_embed.py:68-81
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    t = t + dino_pos_embed(sd, n_side)
    D = t.shape[-1]
    hd = D // heads
    for i in range(depth):
        b = f"blocks.{i}"
        y = _ln(sd, b + ".norm1", t, 1e-6)
        N = y.shape[1]
        qkv = _lin(sd, b + ".attn.qkv", y).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        y = (a @ v).transpose(1, 2).reshape(B, N, D)
        t = t + _lin(sd, b + ".attn.proj", y) * sd[b + ".ls1.gamma"]
        y = _ln(sd, b + ".norm2", t, 1e-6)
        y = _lin(sd, b + ".mlp.fc2", F.gelu(_lin(sd, b + ".mlp.fc1", y)))
        t = t + y * sd[b + ".ls2.gamma"]
    t = _ln(sd, "norm", t, 1e-6)
    return t[:, 1:]


def set_image(sam_sd: SD, dino_sd: SD, img_u8_chw: T, sam_cfg, dino_cfg) -> Tuple[T, T]:
    """SamPredictor.set_torch_image (predictor.py:96-108): SAM features [1,256,64,64] and DINOv2
    patch tokens viewed [1,73,73,1024].  DINOv2 sees the SAM-normalised, zero-padded image
    resized bilinearly to 1022x1022 (:104)."""
    x = preprocess(img_u8_chw)
    feats = sam_encoder(sam_sd, x, *sam_cfg)
    x2 = F.interpolate(x, (1022, 1022), mode="bilinear")
    dino = dino_forward(dino_sd, x2, *dino_cfg).view(1, 73, 73, -1)
    return feats, dino


# --------------------------------------------------------------------------------------
# A6  Prompt encoder                                             prompt_encoder.py:64-218
# --------------------------------------------------------------------------------------
def _pe_encode(sd: SD, coords01: T) -> T:
    """PositionEmbeddingRandom._pe_encoding (prompt_encoder.py:189-196)."""
    g = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    c = (2 * coords01 - 1) @ g
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd: SD) -> T:
    """get_dense_pe (prompt_encoder.py:64-73,198-209) -> [1,256,64,64]."""
    grid = torch.ones((64, 64), dtype=torch.float32)
    y = (grid.cumsum(dim=0) - 0.5) / 64
    x = (grid.cumsum(dim=1) - 0.5) / 64
    return _pe_encode(sd, torch.stack([x, y], dim=-1)).permute(2, 0, 1).unsqueeze(0)


def embed_points(sd: SD, coords: T, labels: T) -> T:
    """_embed_points with pad=True (prompt_encoder.py:75-93,211-218).  coords [P,n,2] keep their
    incoming dtype (float64 from apply_coords) through +0.5 and /1024, then cast to fp32."""
    P = coords.shape[0]
    pts = coords + 0.5
    pts = torch.cat([pts, torch.zeros((P, 1, 2), dtype=pts.dtype)], dim=1)
    lab = torch.cat([labels, -torch.ones((P, 1), dtype=labels.dtype)], dim=1)
    c = pts.clone()
    c[:, :, 0] = c[:, :, 0] / IMG
    c[:, :, 1] = c[:, :, 1] / IMG
    emb = _pe_encode(sd, c.to(torch.float))
    emb[lab == -1] = 0.0
    emb[lab == -1] += sd["prompt_encoder.not_a_point_embed.weight"]
    emb[lab == 0] += sd["prompt_encoder.point_embeddings.0.weight"]
    emb[lab == 1] += sd["prompt_encoder.point_embeddings.1.weight"]
    return emb


# --------------------------------------------------------------------------------------
# A7  Mask decoder + PWD-Net heads            mask_decoder.py:138-199, transformer.py:62-254
# --------------------------------------------------------------------------------------
def _dec_attention(sd: SD, pre: str, q: T, k: T, v: T, heads: int = 8) -> T:
    """transformer.py:228-254: project, split heads, softmax(QK^T/sqrt(hd)) V, out_proj."""
    q, k, v = _lin(sd, pre + ".q_proj", q), _lin(sd, pre + ".k_proj", k), _lin(sd, pre + ".v_proj", v)
    B, _, C = q.shape
    hd = C // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, hd).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = torch.softmax((q @ k.permute(0, 1, 3, 2)) / math.sqrt(hd), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, -1, C)
    return _lin(sd, pre + ".out_proj", o)


def two_way_transformer(sd: SD, src: T, pos: T, tokens: T, pre: str = "mask_decoder.transformer"):
    """TwoWayTransformer.forward (transformer.py:62-114) + TwoWayAttentionBlock (:160-192).
    LayerNorm eps is the nn.LayerNorm default 1e-5 here."""
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos.flatten(2).permute(0, 2, 1)
    queries, qpe = tokens, tokens
    for i in range(2):
        L = f"{pre}.layers.{i}"
        if i == 0:   # skip_first_layer_pe: self-attention REPLACES the queries (:164-165)
            queries = _dec_attention(sd, L + ".self_attn", queries, queries, queries)
        else:
            q = queries + qpe
            queries = queries + _dec_attention(sd, L + ".self_attn", q, q, queries)
        queries = _ln(sd, L + ".norm1", queries, 1e-5)
        queries = queries + _dec_attention(sd, L + ".cross_attn_token_to_image", queries + qpe, keys + kpe, keys)
        queries = _ln(sd, L + ".norm2", queries, 1e-5)
        mlp = _lin(sd, L + ".mlp.lin2", F.relu(_lin(sd, L + ".mlp.lin1", queries)))
        queries = _ln(sd, L + ".norm3", queries + mlp, 1e-5)
        keys = keys + _dec_attention(sd, L + ".cross_attn_image_to_token", keys + kpe, queries + qpe, queries)
        keys = _ln(sd, L + ".norm4", keys, 1e-5)
    queries = queries + _dec_attention(sd, pre + ".final_attn_token_to_image", queries + qpe, keys + kpe, keys)
    return _ln(sd, pre + ".norm_final_attn", queries, 1e-5), keys


def _mlp(sd: SD, pre: str, x: T, n: int) -> T:
    """MLP / DropMLP in eval mode (mask_decoder.py:203-253): ReLU between layers."""
    for i in range(n):
        x = _lin(sd, f"{pre}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def mask_decoder(sd: SD, feats: T, pe: T, sparse: T, dino_feats: T, pre: str = "mask_decoder"):
    """MaskDecoder.predict_masks with multimask_output=True returning ALL 4 masks
    (mask_decoder.py:129-130,138-199).  -> masks [P,4,256,256], iou [P,4], cls [P,4,n_class]."""
    P = sparse.shape[0]
    out_tok = torch.cat([sd[pre + ".iou_token.weight"], sd[pre + ".mask_tokens.weight"]], dim=0)
    tokens = torch.cat((out_tok.unsqueeze(0).expand(P, -1, -1), sparse), dim=1)
    dense = sd["prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1)   # prompt_encoder.py:168-170
    src = torch.repeat_interleave(feats, P, dim=0) + dense
    pos = torch.repeat_interleave(pe, P, dim=0)
    hs, src = two_way_transformer(sd, src, pos, tokens)
    iou_tok, mask_tok = hs[:, 0, :], hs[:, 1:5, :]
    src = src.transpose(1, 2).reshape(P, 256, 64, 64)
    up = F.conv_transpose2d(src, sd[pre + ".output_upscaling.0.weight"], sd[pre + ".output_upscaling.0.bias"], stride=2)
    up = F.gelu(_ln2d(sd, pre + ".output_upscaling.1", up))
    up = F.gelu(F.conv_transpose2d(up, sd[pre + ".output_upscaling.3.weight"], sd[pre + ".output_upscaling.3.bias"], stride=2))
    hyper = torch.stack([_mlp(sd, f"{pre}.output_hypernetworks_mlps.{i}", mask_tok[:, i, :], 3) for i in range(4)], dim=1)
    masks = (hyper @ up.view(P, 32, 256 * 256)).view(P, 4, 256, 256)
    iou = _mlp(sd, pre + ".iou_prediction_head", iou_tok, 3)
    # PWD-Net (mask_decoder.py:187-198)
    dmap = _lin(sd, pre + ".dino_proj", dino_feats)
    dmap = F.interpolate(dmap.permute(0, 3, 1, 2), (256, 256), mode="bilinear")
    wgt = masks.flatten(2).softmax(-1).reshape(P, 4, 256, 256)
    pooled = torch.einsum("blhw,chw->blc", wgt, dmap[0])
    cls = _mlp(sd, pre + ".point_classifier", pooled, 2)
    fused = torch.cat([iou_tok.unsqueeze(1).repeat(1, 4, 1), mask_tok], dim=-1)
    iou = iou + _mlp(sd, pre + ".parallel_iou_head", fused, 3).squeeze(2)
    return masks, iou, cls


def fg_map(sd: SD, dino_feats: T, pre: str = "mask_decoder") -> T:
    """SamPredictor.predict_fg_map (predictor.py:113-121) -> [1,n_class,256,256] logits."""
    d = _lin(sd, pre + ".dino_proj", dino_feats)
    logits = _mlp(sd, pre + ".point_classifier", d, 2).permute(0, 3, 1, 2)
    return F.interpolate(logits, (256, 256), mode="bilinear")


# --------------------------------------------------------------------------------------
# A8  postprocess_masks                                                  sam.py:132-161
# --------------------------------------------------------------------------------------
def postprocess_masks(masks: T, input_size, original_size) -> T:
    m = F.interpolate(masks, (IMG, IMG), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)


def preprocess_shape(oldh: int, oldw: int, long_side: int = IMG) -> Tuple[int, int]:
    """ResizeLongestSide.get_preprocess_shape (transforms.py:93-102)."""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def apply_coords(coords: np.ndarray, original_size) -> np.ndarray:
    """ResizeLongestSide.apply_coords (transforms.py:33-45) -> float64."""
    oh, ow = original_size
    nh, nw = preprocess_shape(oh, ow)
    c = np.array(coords, dtype=float, copy=True)
    c[..., 0] = c[..., 0] * (nw / ow)
    c[..., 1] = c[..., 1] * (nh / oh)
    return c


# --------------------------------------------------------------------------------------
# A10/A11  stability score, boxes                              amg.py:156-176,303-346
# --------------------------------------------------------------------------------------
def stability_score(masks: T, thr: float, off: float) -> T:
    inter = (masks > (thr + off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    union = (masks > (thr - off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return inter / union


def mask_to_box(masks: T) -> T:
    """bool [n,H,W] -> int64 [n,4] inclusive XYXY, empty -> zeros (amg.py:303-346)."""
    n, H, W = masks.shape
    if masks.numel() == 0:
        return torch.zeros(n, 4)
    rows = masks.any(dim=2)
    cols = masks.any(dim=1)
    ar_h, ar_w = torch.arange(H), torch.arange(W)
    bottom = (rows * ar_h).max(dim=1).values
    top = (rows * ar_h + H * (~rows)).min(dim=1).values
    right = (cols * ar_w).max(dim=1).values
    left = (cols * ar_w + W * (~cols)).min(dim=1).values
    empty = (right < left) | (bottom < top)
    out = torch.stack([left, top, right, bottom], dim=-1)
    return out * (~empty).unsqueeze(-1)


# --------------------------------------------------------------------------------------
# A12  NMS (torchvision.ops.nms semantics)               model.py:257-263
# --------------------------------------------------------------------------------------
def nms_reference(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Greedy box NMS with torchvision's arithmetic, restated from the published torchvision
    CPU kernel (torchvision 0.26.0 csrc/ops/cpu/nms_kernel.cpp; dependency absent from the
    reference tree; call sites model.py:171,257,429):
      order = stable descending sort of the scores (torch.sort semantics: NaN first, -0.0 == 0.0);
      area = (x2-x1)*(y2-y1) (no +1);
      suppress j when inter/(area_i+area_j-inter) > thr  (fp32; 0/0 = NaN never suppresses).
    Returns kept ORIGINAL indices in stable descending-score order."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = torch.sort(torch.as_tensor(scores), descending=True, stable=True)[1].numpy()
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep: List[int] = []
    thr32 = np.float32(thr)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(n):
            i = order[a]
            if dead[i]:
                continue
            keep.append(int(i))
            rest = order[a + 1:]
            xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(np.float32(0), xx2 - xx1)
            h = np.maximum(np.float32(0), yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            dead[rest[ovr > thr32]] = True
    return np.asarray(keep, dtype=np.int64)


# --------------------------------------------------------------------------------------
# A14  RLE                                                      amg.py:107-135,294-300
# --------------------------------------------------------------------------------------
def mask_to_rle(mask: np.ndarray) -> Dict:
    """Column-major (Fortran) run lengths starting with the count of zeros (amg.py:107-135)."""
    h, w = mask.shape
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    idx = np.concatenate([[0], change, [h * w]])
    counts = np.diff(idx).tolist()
    if flat[0]:
        counts = [0] + counts
    return {"size": [h, w], "counts": counts}


def coco_rle_string(counts: Sequence[int]) -> str:
    """COCO API rleToString (pycocotools maskApi.c, dependency absent here: PARITY UNPINNED):
    delta-code counts beyond the second against counts[i-2], then 5-bit little-endian groups
    with a continuation bit, chr(48 + group)."""
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def coco_encode_rle(rle: Dict) -> Dict:
    h, w = rle["size"]
    return {"size": [h, w], "counts": coco_rle_string(rle["counts"])}


def coco_rle_decode(s: str, size) -> np.ndarray:
    """Inverse of coco_rle_string (COCO API rleFrString) followed by rle_to_mask."""
    counts: List[int] = []
    p, m = 0, 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if m > 2:
            x += counts[m - 2]
        counts.append(x)
        m += 1
    return rle_to_mask({"size": list(size), "counts": counts})


def rle_to_mask(rle: Dict) -> np.ndarray:
    """Inverse of mask_to_rle (amg.py:138-150)."""
    h, w = rle["size"]
    m = np.empty(h * w, dtype=bool)
    i, par = 0, False
    for c in rle["counts"]:
        m[i:i + c] = par
        i += c
        par = not par
    return m.reshape(w, h).T


# --------------------------------------------------------------------------------------
# A15  mask-overlap NMS (dead code in the reference)         crowdsam/utils.py:422-479
# --------------------------------------------------------------------------------------
def mask_iou_nms(scores: np.ndarray, masks: T, thr: float) -> np.ndarray:
    if masks.numel() == 0:
        return np.zeros((0,), dtype=np.int64)
    m = F.interpolate(masks.float().unsqueeze(0), (150, 150))[0].bool()
    order = np.argsort(-scores).tolist()
    keep: List[int] = []
    for i in order:
        if keep:
            a = m[i].unsqueeze(0)
            b = m[keep]
            inter = (a * b).sum([-1, -2])
            cov = torch.maximum(inter / a.sum([-1, -2]), inter / b.sum([-1, -2]))
            if torch.any(cov > thr):
                continue
        keep.append(i)
    return np.asarray(keep, dtype=np.int64)


# --------------------------------------------------------------------------------------
# small-region cleanup                                     amg.py:267-291, model.py:395-443
# --------------------------------------------------------------------------------------
def remove_small_regions(mask: np.ndarray, area_thresh: float, mode: str):
    import cv2

    holes = mode == "holes"
    work = (holes ^ mask).astype(np.uint8)
    n, regions, stats, _ = cv2.connectedComponentsWithStats(work, 8)
    sizes = stats[:, -1][1:]
    small = [i + 1 for i, s in enumerate(sizes) if s < area_thresh]
    if not small:
        return mask, False
    fill = [0] + small
    if not holes:
        fill = [i for i in range(n) if i not in fill]
        if not fill:
            fill = [int(np.argmax(sizes)) + 1]
    return np.isin(regions, fill), True


# --------------------------------------------------------------------------------------
# L4 pipeline: CrowdSAM.generate                                        model.py:134-449
# --------------------------------------------------------------------------------------
from crowdsam_b200.synthetic import DEFAULT_TEST_CFG  # noqa: E402  configs/crowdhuman.yaml:34-60 (shared input data)


def resize_image(image: np.ndarray, max_size: int):
    """crowdsam/utils.py:141-156 (cv2.resize default INTER_LINEAR; may up-scale)."""
    import cv2

    h, w = image.shape[:2]
    r = min(max_size / w, max_size / h)
    h, w = int(r * h), int(r * w)
    return cv2.resize(image, (w, h)), r


def crop_boxes_for(im_size, n_layers: int, overlap_ratio: float):
    """generate_crop_boxes (amg.py:200-234)."""
    im_h, im_w = im_size
    boxes = [[0, 0, im_w, im_h]]
    short = min(im_h, im_w)
    for layer in range(n_layers):
        n = 2 ** (layer + 1)
        ov = int(overlap_ratio * short * (2 / n))
        cw = int(math.ceil((ov * (n - 1) + im_w) / n))
        ch = int(math.ceil((ov * (n - 1) + im_h) / n))
        xs = [int((cw - ov) * i) for i in range(n)]
        ys = [int((ch - ov) * i) for i in range(n)]
        for x0 in xs:
            for y0 in ys:
                boxes.append([x0, y0, min(x0 + cw, im_w), min(y0 + ch, im_h)])
    return boxes


def near_crop_edge(boxes: T, crop_box, orig_box, downscale: float, atol: float = 20.0) -> T:
    """crowdsam/utils.py:213-223."""
    cb = torch.as_tensor(crop_box, dtype=torch.float)
    ob = torch.as_tensor(orig_box, dtype=torch.float)
    x0, y0 = crop_box[0], crop_box[1]
    b = (boxes / downscale + torch.tensor([[x0, y0, x0, y0]])).float()
    near_c = torch.isclose(b, cb[None, :], atol=atol, rtol=0)
    near_i = torch.isclose(b, ob[None, :], atol=atol, rtol=0)
    return torch.any(near_c & ~near_i, dim=1)


class OracleCrowdSAM:
    """Restatement of crowdsam.model.CrowdSAM (model.py:24-449) for n_class = 1, trainfree=False."""

    def __init__(self, sam_sd: SD, dino_sd: SD, sam_cfg, dino_cfg, test_cfg: Optional[dict] = None):
        self.sam_sd, self.dino_sd = sam_sd, dino_sd
        self.sam_cfg, self.dino_cfg = sam_cfg, dino_cfg
        self.cfg = dict(DEFAULT_TEST_CFG)
        if test_cfg:
            self.cfg.update(test_cfg)
        self.pe = dense_pe(sam_sd)
        self.stage_times: Dict[str, float] = {}

    # ---- predict_torch (predictor.py:214-292)
    def predict(self, feats, dino, coords: T, labels: T, input_size, original_size):
        sparse = embed_points(self.sam_sd, coords, labels)
        low, iou, cls = mask_decoder(self.sam_sd, feats, self.pe, sparse, dino)
        return postprocess_masks(low, input_size, original_size), iou, cls, low

    def select(self, masks: T, iou: T):
        """select_mask (model.py:318-331)."""
        mode = self.cfg["mask_selection"]
        binm = masks > 0.0
        if mode == "max_area":
            ind = binm.sum(dim=[-1, -2]).max(dim=-1)[1]
        elif mode == "min_area":
            ind = binm.sum(dim=[-1, -2]).min(dim=-1)[1]
        elif mode == "max_iou":
            ind = iou.max(dim=-1)[1]
        else:
            raise NotImplementedError
        return torch.arange(len(masks)), ind

    def process_batch(self, feats, dino, points: np.ndarray, input_size, original_size, crop_box,
                      orig_hw, downscale):
        """_process_batch (model.py:334-390)."""
        c = self.cfg
        tp = apply_coords(points, original_size)
        coords = torch.as_tensor(tp)[:, None, :]
        labels = torch.ones(coords.shape[0], dtype=torch.int)[:, None]
        masks, iou, cls, _ = self.predict(feats, dino, coords, labels, input_size, original_size)
        iou = torch.clamp(iou, 0.0) * cls.squeeze(2).sigmoid()
        idx = self.select(masks, iou)
        cats = cls.max(dim=-1)[1]
        d = dict(masks=masks[idx], iou_preds=iou[idx], points=torch.as_tensor(points), categories=cats[idx])

        def filt(keep):
            for k in d:
                d[k] = d[k][keep]

        if c["pred_iou_thresh"] > 0.0:
            filt(d["iou_preds"] > c["pred_iou_thresh"])
        d["stability_score"] = stability_score(d["masks"], 0.0, c["stability_score_offset"])
        if c["stability_score_thresh"] > 0.0:
            filt(d["stability_score"] >= c["stability_score_thresh"])
        d["masks"] = d["masks"] > 0.0
        d["boxes"] = mask_to_box(d["masks"])
        keep = ~near_crop_edge(d["boxes"], crop_box, [0, 0, orig_hw[1], orig_hw[0]], downscale)
        if not torch.all(keep):
            filt(keep)
        return d

    def process_crop(self, image: np.ndarray, crop_box):
        """_process_crop (model.py:191-305)."""
        import time

        c = self.cfg
        x0, y0, x1, y1 = crop_box
        orig_hw = image.shape[:2]
        img, r = resize_image(image[y0:y1, x0:x1, :], c["max_size"])
        t0 = time.perf_counter()
        # predictor.set_image (predictor.py:32-69): PIL bilinear resize to long side 1024
        ih, iw = preprocess_shape(img.shape[0], img.shape[1])
        if (ih, iw) != img.shape[:2]:
            from PIL import Image

            inp = np.array(Image.fromarray(img).resize((iw, ih), Image.BILINEAR))
        else:
            inp = img
        t = torch.as_tensor(inp).permute(2, 0, 1).contiguous()[None]
        feats, dino = set_image(self.sam_sd, self.dino_sd, t, self.sam_cfg, self.dino_cfg)
        self.stage_times["set_image"] = self.stage_times.get("set_image", 0.0) + time.perf_counter() - t0
        input_size, original_size = (ih, iw), img.shape[:2]
        # foreground prior -> candidate points (model.py:196-223, 445-449)
        G = c["grid_size"]
        img_size = torch.tensor(img.shape[:2])
        feat_size = (img_size * min(G / img_size)).int()
        sim = fg_map(self.sam_sd, dino)
        sim = F.interpolate(sim, (G, G), mode="bilinear").sigmoid().max(dim=1)[0]
        sim = sim[0, : feat_size[0], : feat_size[1]]
        coords = (sim > c["pos_sim_thresh"]).nonzero()[:, [1, 0]]
        inv = torch.tensor([feat_size[1] / img.shape[1], feat_size[0] / img.shape[0]])
        pts = (coords / inv).numpy()
        # EPS iterator (model.py:229-248)
        data: Dict[str, T] = {}
        occupy = torch.zeros(*img.shape[:2], dtype=torch.bool)
        pts = pts.astype("int")
        np.random.shuffle(pts)
        count, bs = 0, c["points_per_batch"]
        t0 = time.perf_counter()
        while len(pts) > 0 and count < c["max_prompts"]:
            bs = min(len(pts), bs)
            sel, pts = pts[:bs], pts[bs:]
            d = self.process_batch(feats, dino, sel, input_size, original_size, crop_box, orig_hw, r)
            occupy = d["masks"][d["iou_preds"] > c["filter_thresh"]].any(0)
            for k, v in d.items():
                data[k] = v if k not in data else torch.cat([data[k], v], dim=0)
            keep = (~occupy[pts[:, 1], pts[:, 0]]).numpy()
            pts = pts[keep]
            count += bs
        self.stage_times["decode"] = self.stage_times.get("decode", 0.0) + time.perf_counter() - t0
        if not data or len(data["masks"]) == 0:
            return None
        t0 = time.perf_counter()
        keep = torch.as_tensor(nms_reference(data["boxes"].float().numpy(), data["iou_preds"].numpy(),
                                             c["box_nms_thresh"]))
        for k in data:
            data[k] = data[k][keep]
        if c["min_mask_region_area"] > 0:
            data = self.small_regions(data, c["min_mask_region_area"],
                                      max(c["box_nms_thresh"], c["crop_nms_thresh"]))
        data["scores"] = data["iou_preds"]
        rles = [mask_to_rle(m) for m in data["masks"].numpy()]
        del data["masks"]
        off = torch.tensor([[x0, y0, x0, y0]])
        data["boxes"] = data["boxes"] / r + off
        data["points"] = data["points"] / r + off[:, :2]
        data["crop_boxes"] = torch.tensor([crop_box for _ in range(len(data["boxes"]))])
        data["fboxes"] = data["boxes"]
        out = dict(data)
        out["rles"] = rles
        out["rles_info"] = [crop_box, [orig_hw[0], orig_hw[1]]]
        self.stage_times["select"] = self.stage_times.get("select", 0.0) + time.perf_counter() - t0
        return out

    @staticmethod
    def small_regions(data, min_area, nms_thresh):
        """postprocess_small_regions (model.py:395-443)."""
        if len(data["masks"]) == 0:
            return data
        new, scores = [], []
        for m in data["masks"].numpy():
            m, ch1 = remove_small_regions(m, min_area, "holes")
            m, ch2 = remove_small_regions(m, min_area, "islands")
            new.append(torch.as_tensor(m).unsqueeze(0))
            scores.append(float(not ch1 and not ch2))
        masks = torch.cat(new, dim=0)
        boxes = mask_to_box(masks)
        keep = nms_reference(boxes.float().numpy(), np.asarray(scores, dtype=np.float32), nms_thresh)
        for i in keep:
            if scores[i] == 0.0:
                data["boxes"][i] = boxes[i]
                data["masks"][i] = masks[i]
        keep = torch.as_tensor(keep)
        for k in data:
            data[k] = data[k][keep]
        return data

    def generate(self, image: np.ndarray) -> Dict:
        """_generate_masks (model.py:151-189)."""
        c = self.cfg
        image = np.asarray(image, dtype=np.uint8)
        boxes = crop_boxes_for(image.shape[:2], c["crop_n_layers"], c["crop_overlap_ratio"])
        data: Dict = {}
        for cb in boxes:
            d = self.process_crop(image, cb)
            if d is None:
                continue
            for k, v in d.items():
                if k not in data:
                    data[k] = v
                elif isinstance(v, torch.Tensor):
                    data[k] = torch.cat([data[k], v], dim=0)
                else:
                    data[k] = data[k] + v
        if len(boxes) > 1 and "crop_boxes" in data and len(data["crop_boxes"]) > 0:
            cbx = data["crop_boxes"]
            sc = 1.0 / ((cbx[:, 2] - cbx[:, 0]) * (cbx[:, 3] - cbx[:, 1])).float()
            keep = nms_reference(data["boxes"].float().numpy(), sc.numpy(), c["crop_nms_thresh"])
            for k, v in list(data.items()):
                if isinstance(v, torch.Tensor):
                    data[k] = v[torch.as_tensor(keep)]
                elif k == "rles":
                    data[k] = [v[i] for i in keep]
            del data["crop_boxes"]
        if data:
            del data["iou_preds"]
        else:
            data["boxes"] = torch.zeros(0, 4)
            data["scores"] = torch.zeros(0, 4)
        data["rles"] = [coco_encode_rle(r) for r in data.get("rles", [])]
        return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
