"""Seeded synthetic stage-level inputs shared by golden generation and the parity tests.

TEST INFRASTRUCTURE ONLY.  Random weights give spatial-noise masks whose boxes are always
the full image, so box / stability / NMS decisions are exercised with
per-prompt Gaussian-blob logits injected at the low-res-logit boundary.
"""
from __future__ import annotations

import numpy as np
import torch


def blob_logits(P: int, seed: int = 0, lowres: int = 256, four: bool = True):
    """Low-res logits [P,4,256,256] (or [P,256,256]), fp32: +6 blob over a -6 floor, sigma in
    U[8,48] low-res px, N(0,0.5) noise; the four candidates of a prompt share a centre with
    growing sigma.  Also returns scores [P,4] in U(0,1) with ~10% exact ties and cls logits."""
    g = torch.Generator().manual_seed(1000 + seed)
    cx = torch.rand(P, generator=g) * (lowres - 1)
    cy = torch.rand(P, generator=g) * (lowres - 1)
    sig = 8 + 40 * torch.rand(P, generator=g)
    yy, xx = torch.meshgrid(torch.arange(lowres, dtype=torch.float32),
                            torch.arange(lowres, dtype=torch.float32), indexing="ij")
    C = 4 if four else 1
    out = torch.empty(P, C, lowres, lowres)
    for c in range(C):
        s = (sig * (0.6 + 0.3 * c)).view(P, 1, 1)
        d2 = (xx[None] - cx.view(P, 1, 1)) ** 2 + (yy[None] - cy.view(P, 1, 1)) ** 2
        out[:, c] = -6.0 + 12.0 * torch.exp(-d2 / (2 * s * s))
    out += 0.5 * torch.randn(out.shape, generator=g)
    iou = torch.rand(P, C, generator=g)
    tie = torch.rand(P, generator=g) < 0.1
    iou[tie] = (iou[tie] * 4).round() / 4          # exact ties across prompts
    cls = 3.0 * torch.randn(P, C, 1, generator=g)
    if not four:
        out = out[:, 0]
    return out, iou, cls


def random_boxes(n: int, seed: int = 0, extent: float = 1024.0, tie_frac: float = 0.1,
                 binary_scores: bool = False):
    """Boxes [n,4] fp32 XYXY clustered so that many pairs overlap, scores with exact ties,
    a few zero-area and duplicate boxes (NaN IoU path)."""
    rng = np.random.default_rng(2000 + seed)
    n_c = max(1, n // 8)
    centres = rng.uniform(0, extent, (n_c, 2))
    which = rng.integers(0, n_c, n)
    c = centres[which] + rng.normal(0, 12, (n, 2))
    wh = rng.uniform(8, 160, (n, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    boxes = np.clip(np.round(boxes), 0, extent - 1).astype(np.float32)
    if n >= 16:
        boxes[3] = boxes[2]                       # exact duplicate
        boxes[5, 2:] = boxes[5, :2]               # zero area
        boxes[6] = boxes[5]                       # duplicate zero-area pair -> 0/0
    if binary_scores:
        scores = (rng.uniform(0, 1, n) < 0.5).astype(np.float32)
    else:
        scores = rng.uniform(0, 1, n).astype(np.float32)
        tie = rng.uniform(0, 1, n) < tie_frac
        scores[tie] = np.round(scores[tie] * 8) / 8
    return boxes, scores


def grid_points(G: int, size: int = 1024) -> np.ndarray:
    """Prompt grid produced by the reference config overrides of SURVEY.md §8d: pixel
    (j*size/G, i*size/G) for every grid cell, row-major (model.py:200-223)."""
    ii, jj = np.meshgrid(np.arange(G), np.arange(G), indexing="ij")
    return np.stack([jj.reshape(-1) * (size / G), ii.reshape(-1) * (size / G)], axis=1).astype(int)


def cluster_points(n: int, seed: int = 0, n_clusters: int = 12, spread: float = 30.0, size: int = 1024) -> np.ndarray:
    """n integer prompt points (x, y) around a few cluster centres, so that the injected instances overlap heavily."""
    rng = np.random.default_rng(3000 + seed)
    centres = rng.uniform(100, size - 100, (n_clusters, 2))
    pts = centres[rng.integers(0, n_clusters, n)] + rng.normal(0, spread, (n, 2))
    return np.clip(np.round(pts), 0, size - 1).astype(np.float64)


# ------------------------------------------------------------------------------------------------
# Decoder-output injection keyed by the prompt point (multi-detection end-to-end parity)
# ------------------------------------------------------------------------------------------------
def injected_decoder_outputs(coords_xy: np.ndarray, seed: int = 0, r_lo: float = 4.0, r_hi: float = 22.0,
                             lowres: int = 256):
    """Deterministic stand-in for the mask decoder's three outputs, a pure function of each prompt's
    input-frame coordinates (the `apply_coords` output both the reference's `predict_torch` and this
    repo's `decode_low_res` receive): low-res logits [P,4,256,256] = 6*tanh((r - d)/tau) for an elliptical
    distance d from a centre near the prompt (plateau +6 inside, -6 outside, edge width tau in U[0.4,3] low-res
    px so that the stability score spreads over ~[0.55, 0.97]), about half of the prompts with a small hole
    and / or a small satellite island (areas around the 100 px cleanup threshold at full resolution), plus
    N(0,0.5) pixel noise (ragged edges); the four candidates share the centre with growing radius.
    iou [P,4] in U(0,1) and cls logits [P,4,1]; about 10 % of the prompts have both quantised (exact score
    ties across prompts).  Injected on both sides at the low-res-logit boundary so that the whole
    selection chain (PWD score, select_mask, filters, boxes, EPS pruning, NMS, small-region cleanup, RLE,
    crop merge) of the REAL reference runs on many distinct instances (tests/golden/make_golden.py) and is
    compared with the CUDA path on identical inputs."""
    pts = np.asarray(coords_xy, dtype=np.float64).reshape(-1, 2)
    P = pts.shape[0]
    yy, xx = np.meshgrid(np.arange(lowres, dtype=np.float32), np.arange(lowres, dtype=np.float32), indexing="ij")
    low = np.empty((P, 4, lowres, lowres), dtype=np.float32)
    iou = np.empty((P, 4), dtype=np.float32)
    cls = np.empty((P, 4, 1), dtype=np.float32)
    six = np.float32(6.0)
    for p in range(P):
        kx, ky = int(round(pts[p, 0] * 4)), int(round(pts[p, 1] * 4))
        rng = np.random.default_rng([int(seed), kx & 0xFFFFFFFF, ky & 0xFFFFFFFF])
        cx = np.float32(pts[p, 0] / 4.0 + rng.normal(0.0, 1.5))
        cy = np.float32(pts[p, 1] / 4.0 + rng.normal(0.0, 1.5))
        r = np.float32(rng.uniform(r_lo, r_hi))
        tau = np.float32(rng.uniform(0.4, 3.0))
        aspect = np.float32(rng.uniform(0.6, 1.8))              # people are taller than wide
        d = np.sqrt((xx - cx) ** 2 + ((yy - cy) / aspect) ** 2)
        extras = []
        for sign in (-1.0, 1.0):                                # a hole inside / an island outside
            if rng.uniform() < 0.5:
                rr = np.float32(rng.uniform(1.0, 2.6))
                ang, rad = rng.uniform(0, 2 * np.pi), (rng.uniform(0.0, 0.5) if sign < 0 else rng.uniform(1.5, 2.0))
                ex = cx + np.float32(rad * np.cos(ang)) * r
                ey = cy + np.float32(rad * np.sin(ang)) * r * aspect
                de = np.sqrt((xx - ex) ** 2 + (yy - ey) ** 2)
                extras.append((np.float32(sign), np.tanh((rr - de) / np.float32(0.5))))
        for c in range(4):
            rc = r * np.float32(0.6 + 0.3 * c)
            m = six * np.tanh((rc - d) / tau)
            for sign, e in extras:                               # e = +1 inside the feature, -1 outside
                m = np.where(e > 0, sign * six * e, m) if c == 0 or sign < 0 else m
            low[p, c] = m
        low[p] += np.float32(0.5) * rng.standard_normal(low[p].shape, dtype=np.float32)
        q = rng.uniform(0.0, 1.0, 4).astype(np.float32)
        k = (3.0 * rng.standard_normal(4)).astype(np.float32)
        if rng.uniform() < 0.1:
            q, k = np.round(q * 4) / 4, np.round(k)
        iou[p] = q
        cls[p, :, 0] = k
    return torch.from_numpy(low), torch.from_numpy(iou), torch.from_numpy(cls)
