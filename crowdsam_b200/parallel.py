"""Multi-GPU plumbing: images are independent units, one process per GPU.

Reference behaviour replaced: tools/batch_eval.py:80-103 fans `tools/test.py` out as one subprocess per GPU
over contiguous index slices (last rank takes the remainder, :82-89) and merges `temp_result_{rank}.json`
files.  Here the slices are the same and the merge is ONE all-gather of a padded detection buffer + counts
(NCCL on GPUs; the same code runs over gloo on CPU tensors for tests).  No collective touches the data path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of rank `rank` (batch_eval.py:82-89)."""
    per = n_items // world
    start = rank * per
    end = n_items if rank == world - 1 else (rank + 1) * per
    return start, end


def pack_detections(dets: Sequence[Optional[Dict]], nmax: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_images, nmax, 6] = (x0, y0, x1, y1, score, category), zero padded, + int32 counts."""
    buf = torch.zeros((len(dets), nmax, 6), dtype=torch.float32, device=device)
    cnt = torch.zeros((len(dets),), dtype=torch.int32, device=device)
    for i, d in enumerate(dets):
        if d is None or len(d["boxes"]) == 0:
            continue
        n = min(len(d["boxes"]), nmax)
        buf[i, :n, :4] = torch.as_tensor(d["boxes"][:n], dtype=torch.float32, device=device)
        buf[i, :n, 4] = torch.as_tensor(d["scores"][:n], dtype=torch.float32, device=device)
        buf[i, :n, 5] = torch.as_tensor(d["categories"][:n], dtype=torch.float32, device=device)
        cnt[i] = n
    return buf, cnt


def gather_detections(dets: Sequence[Optional[Dict]], nmax: int = 64, device="cpu", group=None) -> List[List[Dict]]:
    """All-gather every rank's detections.  Ranks may hold different numbers of images (the last slice is
    longer): image counts are exchanged first and buffers padded to the maximum.
    Returns, on every rank, a list over ranks of per-image dicts {boxes [n,4], scores [n], categories [n]}."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    buf, cnt = pack_detections(dets, nmax, device)
    if world == 1:
        bufs, cnts, nimg = [buf], [cnt], [len(dets)]
    else:
        n_local = torch.tensor([len(dets)], dtype=torch.int64, device=device)
        n_all = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(n_all, n_local, group=group)
        nimg = [int(x.item()) for x in n_all]
        top = max(nimg)
        pbuf = torch.zeros((top, nmax, 6), dtype=torch.float32, device=device)
        pcnt = torch.zeros((top,), dtype=torch.int32, device=device)
        pbuf[: len(dets)] = buf
        pcnt[: len(dets)] = cnt
        bufs = [torch.empty_like(pbuf) for _ in range(world)]
        cnts = [torch.empty_like(pcnt) for _ in range(world)]
        dist.all_gather(bufs, pbuf, group=group)
        dist.all_gather(cnts, pcnt, group=group)
    out: List[List[Dict]] = []
    for r in range(world):
        per_rank = []
        for i in range(nimg[r]):
            n = int(cnts[r][i])
            b = bufs[r][i, :n].cpu()
            per_rank.append({"boxes": b[:, :4].numpy(), "scores": b[:, 4].numpy(), "categories": b[:, 5].long().numpy()})
        out.append(per_rank)
    return out
