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
