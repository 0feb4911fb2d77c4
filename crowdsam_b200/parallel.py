"""Multi-GPU plumbing: images are independent units, one process per GPU.

Reference behaviour replaced: tools/batch_eval.py:80-103 fans `tools/test.py` out as one subprocess per GPU
over contiguous index slices (last rank takes the remainder, :82-89) and merges `temp_result_{rank}.json`
files.  Here the slices are the same and the merge is ONE all-gather of a padded detection buffer + counts
(NCCL on GPUs; the same code runs over gloo on CPU tensors for tests).  No collective touches the data path.

Nothing is ever truncated: the padded buffer is sized by the exchanged global maximum of the per-image detection
counts (CrowdHuman images carry hundreds of persons; the reference's JSON merge keeps all of them).  A caller that
passes an explicit `nmax` smaller than a local count gets a ValueError, never a silently shortened list.
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


def _count(d: Optional[Dict]) -> int:
    return 0 if d is None else len(d["boxes"])


def pack_detections(dets: Sequence[Optional[Dict]], nmax: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_images, nmax, 6] = (x0, y0, x1, y1, score, category), zero padded, + int32 counts.
    Raises ValueError if an image holds more than `nmax` detections (no truncation)."""
    top = max([_count(d) for d in dets], default=0)
    if top > nmax:
        raise ValueError(f"pack_detections: an image holds {top} detections but the buffer takes {nmax}; "
                         "pass nmax=None to gather_detections to size it from the global maximum")
    buf = torch.zeros((len(dets), nmax, 6), dtype=torch.float32, device=device)
    cnt = torch.zeros((len(dets),), dtype=torch.int32, device=device)
    for i, d in enumerate(dets):
        n = _count(d)
        if n == 0:
            continue
        buf[i, :n, :4] = torch.as_tensor(d["boxes"], dtype=torch.float32, device=device).reshape(n, 4)
        buf[i, :n, 4] = torch.as_tensor(d["scores"], dtype=torch.float32, device=device).reshape(n)
        buf[i, :n, 5] = torch.as_tensor(d["categories"], dtype=torch.float32, device=device).reshape(n)
        cnt[i] = n
    return buf, cnt


def gather_detections(dets: Sequence[Optional[Dict]], nmax: Optional[int] = None, device="cpu",
                      group=None) -> List[List[Dict]]:
    """All-gather every rank's detections.  Ranks may hold different numbers of images (the last slice is
    longer) and images different numbers of detections: (image count, max detection count) are exchanged first
    and the buffers padded to the global maxima.  `nmax=None` (default) sizes the buffer from that exchange.
    Returns, on every rank, a list over ranks of per-image dicts {boxes [n,4], scores [n], categories [n]}."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_top = max([_count(d) for d in dets], default=0)
    if nmax is not None and local_top > nmax:
        raise ValueError(f"gather_detections: {local_top} detections in one image exceed nmax={nmax}")
    if world == 1:
        width = max(local_top, 1) if nmax is None else nmax
        buf, cnt = pack_detections(dets, width, device)
        bufs, cnts, nimg = [buf], [cnt], [len(dets)]
    else:
        meta = torch.tensor([len(dets), local_top], dtype=torch.int64, device=device)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        nimg = [int(m[0].item()) for m in metas]
        width = max(max(int(m[1].item()) for m in metas), 1) if nmax is None else nmax
        buf, cnt = pack_detections(dets, width, device)
        top = max(nimg)
        pbuf = torch.zeros((top, width, 6), dtype=torch.float32, device=device)
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
        cr = cnts[r].cpu()
        br = bufs[r].cpu()
        for i in range(nimg[r]):
            n = int(cr[i])
            b = br[i, :n]
            per_rank.append({"boxes": b[:, :4].numpy(), "scores": b[:, 4].numpy(), "categories": b[:, 5].long().numpy()})
        out.append(per_rank)
    return out


def gather_objects(local: List, group=None) -> List:
    """Rank-ordered concatenation of per-rank Python lists (the per-image result dicts with their COCO RLE strings):
    the variable-length payload of `result.json` that the reference merges through temp files (batch_eval.py:17-28)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return list(local)
    parts: List = [None] * world
    dist.all_gather_object(parts, list(local), group=group)
    return [x for p in parts for x in p]
