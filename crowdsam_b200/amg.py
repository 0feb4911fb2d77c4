"""Mask-data container and selection utilities of the hot path, backed by the CUDA kernels.

Mirrors the names of segment_anything_cs/utils/amg.py that callers use (MaskData :16-75,
calculate_stability_score :156-176, batched_mask_to_box :303-346, mask_to_rle_pytorch :107-135,
coco_encode_rle :294-300, generate_crop_boxes :200-234, remove_small_regions :267-291,
build_point_grid :179-186, batch_iterator :97-104).
"""
from __future__ import annotations

import math
from typing import Any, Dict, Iterator, List, Sequence, Tuple

import numpy as np
import torch

from . import ops

_OK = (list, np.ndarray, torch.Tensor)


class MaskData:
    """Column store of per-mask fields with row filtering / concatenation."""

    def __init__(self, **fields) -> None:
        self._stats: Dict[str, Any] = {}
        for k, v in fields.items():
            self[k] = v

    def __setitem__(self, key: str, item: Any) -> None:
        assert isinstance(item, _OK), "MaskData only supports list, numpy arrays, and torch tensors."
        self._stats[key] = item

    def __getitem__(self, key: str) -> Any:
        return self._stats[key]

    def __delitem__(self, key: str) -> None:
        del self._stats[key]

    def items(self):
        return self._stats.items()

    def filter(self, keep: torch.Tensor) -> None:
        is_bool = keep.dtype == torch.bool
        for k, v in self._stats.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                self._stats[k] = v[torch.as_tensor(keep, device=v.device)]
            elif isinstance(v, np.ndarray):
                self._stats[k] = v[keep.detach().cpu().numpy()]
            elif isinstance(v, list):
                idx = [i for i, f in enumerate(keep.tolist()) if f] if is_bool else [int(i) for i in keep]
                self._stats[k] = [v[i] for i in idx]
            else:
                raise TypeError(f"MaskData key {k} has an unsupported type {type(v)}.")

    def cat(self, other: "MaskData") -> None:
        from copy import deepcopy

        for k, v in other.items():
            cur = self._stats.get(k)
            if cur is None:
                self._stats[k] = v if isinstance(v, torch.Tensor) else deepcopy(v)
            elif isinstance(v, torch.Tensor):
                self._stats[k] = torch.cat([cur, v], dim=0)
            elif isinstance(v, np.ndarray):
                self._stats[k] = np.concatenate([cur, v], axis=0)
            elif isinstance(v, list):
                self._stats[k] = cur + deepcopy(v)
            else:
                raise TypeError(f"MaskData key {k} has an unsupported type {type(v)}.")

    @staticmethod
    def merged(parts: "List[MaskData]") -> "MaskData":
        """Concatenate many batches with ONE copy per field (cat() in a loop re-copies the growing
        [N,H,W] mask tensor every batch)."""
        out = MaskData()
        if not parts:
            return out
        for k in parts[0]._stats:
            vals = [p._stats[k] for p in parts]
            if isinstance(vals[0], torch.Tensor):
                out._stats[k] = vals[0] if len(vals) == 1 else torch.cat(vals, dim=0)
            elif isinstance(vals[0], np.ndarray):
                out._stats[k] = np.concatenate(vals, axis=0)
            else:
                out._stats[k] = [x for v in vals for x in v]
        return out

    def to_numpy(self) -> None:
        dev = [k for k, v in self._stats.items() if isinstance(v, torch.Tensor) and v.is_cuda]
        if len(dev) > 1:
            # one device-to-host read for all device columns instead of one synchronising copy per column
            for k, a in zip(dev, tensors_to_numpy_packed([self._stats[k] for k in dev])):
                self._stats[k] = a
        for k, v in self._stats.items():
            if isinstance(v, torch.Tensor):
                self._stats[k] = v.detach().cpu().numpy()


def tensors_to_numpy_packed(tensors: List[torch.Tensor]) -> List[np.ndarray]:
    """[t.cpu().numpy() for t in tensors] through ONE transfer: the tensors' bytes are concatenated on their device
    (each padded to 8 bytes), copied once and cut up again on the host.  Same dtypes, shapes and values."""
    parts, meta, off = [], [], 0
    for t in tensors:
        t = t.detach().contiguous()
        b = t.reshape(-1).view(torch.uint8)
        pad = (-b.numel()) % 8
        parts.append(b)
        if pad:
            parts.append(torch.zeros(pad, dtype=torch.uint8, device=t.device))
        meta.append((off, b.numel(), tuple(t.shape), torch.empty(0, dtype=t.dtype).numpy().dtype))
        off += b.numel() + pad
    flat = torch.cat(parts).cpu().numpy() if off else np.zeros(0, dtype=np.uint8)
    return [flat[o:o + n].copy().view(dt).reshape(shape) for o, n, shape, dt in meta]


def batch_iterator(batch_size: int, *args) -> Iterator[List[Any]]:
    assert args and all(len(a) == len(args[0]) for a in args), "Batched iteration must have inputs of all the same size."
    for s in range(0, len(args[0]), batch_size):
        yield [a[s:s + batch_size] for a in args]


def build_point_grid(n_per_side: int) -> np.ndarray:
    off = 1 / (2 * n_per_side)
    side = np.linspace(off, 1 - off, n_per_side)
    xs, ys = np.meshgrid(side, side)
    return np.stack([xs, ys], axis=-1).reshape(-1, 2)


def generate_crop_boxes(im_size: Tuple[int, ...], n_layers: int, overlap_ratio: float):
    im_h, im_w = im_size
    short = min(im_h, im_w)
    boxes, layers = [[0, 0, im_w, im_h]], [0]
    for li in range(n_layers):
        n = 2 ** (li + 1)
        ov = int(overlap_ratio * short * (2 / n))
        cw = int(math.ceil((ov * (n - 1) + im_w) / n))
        ch = int(math.ceil((ov * (n - 1) + im_h) / n))
        for x0 in (int((cw - ov) * i) for i in range(n)):
            for y0 in (int((ch - ov) * i) for i in range(n)):
                boxes.append([x0, y0, min(x0 + cw, im_w), min(y0 + ch, im_h)])
                layers.append(li + 1)
    return boxes, layers


def calculate_stability_score(masks: torch.Tensor, mask_threshold: float, threshold_offset: float) -> torch.Tensor:
    """Stability of already materialised full-size logits [n,H,W] (K-POST counts, identity geometry is not
    assumed: this entry point re-counts on the given tensor with the library's low-res kernel when the
    input is 256x256, otherwise with exact integer counts)."""
    hi = (masks > (mask_threshold + threshold_offset)).flatten(1).sum(1, dtype=torch.int32)
    lo = (masks > (mask_threshold - threshold_offset)).flatten(1).sum(1, dtype=torch.int32)
    return hi / lo


def batched_mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    """Inclusive XYXY boxes of bool masks [...,H,W]; empty -> zeros."""
    if masks.numel() == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    shape = masks.shape
    h, w = shape[-2:]
    m = masks.reshape(-1, h, w)
    rows, cols = m.any(dim=2), m.any(dim=1)
    ys = torch.arange(h, device=m.device)
    xs = torch.arange(w, device=m.device)
    bottom = (rows * ys).amax(1)
    top = (rows * ys + h * (~rows)).amin(1)
    right = (cols * xs).amax(1)
    left = (cols * xs + w * (~cols)).amin(1)
    out = torch.stack([left, top, right, bottom], dim=-1)
    out = out * (~((right < left) | (bottom < top))).unsqueeze(-1)
    return out.reshape(*shape[:-2], 4)


def mask_to_rle_pytorch(tensor: torch.Tensor) -> List[Dict[str, Any]]:
    """Uncompressed column-major RLE via the csam_rle kernels (one D2H of the run lengths)."""
    n, h, w = tensor.shape
    if n == 0:
        return []
    runs = ops.rle_encode(tensor.to(torch.bool))
    return [{"size": [h, w], "counts": r.tolist()} for r in runs]


def mask_to_rle_arrays(tensor: torch.Tensor) -> List[Dict[str, Any]]:
    """Same as mask_to_rle_pytorch but keeps the run lengths as numpy arrays (no per-element Python objects);
    used inside the pipeline where the counts go straight into the COCO string encoder."""
    n, h, w = tensor.shape
    if n == 0:
        return []
    return [{"size": [h, w], "counts": r} for r in ops.rle_encode(tensor.to(torch.bool))]


def rle_to_mask(rle: Dict[str, Any]) -> np.ndarray:
    h, w = rle["size"]
    vals = np.zeros(len(rle["counts"]), dtype=bool)
    vals[1::2] = True
    return np.repeat(vals, rle["counts"]).reshape(w, h).T


def area_from_rle(rle: Dict[str, Any]) -> int:
    return sum(rle["counts"][1::2])


def _coco_string(counts: Sequence[int]) -> str:
    """COCO API run-length string (maskApi.c rleToString): counts beyond the second are delta coded
    against counts[i-2]; each value is emitted as 5-bit groups (little endian), bit 5 = continuation, + 48.
    Vectorised over all counts: one numpy pass per 5-bit group (at most 7 for int32 values)."""
    c = np.asarray(counts)
    # run lengths of a mask fit 32 bits (h * w < 2^31), and so do their pairwise differences: half the bytes to move
    c = c.astype(np.int32 if (c.size == 0 or int(c.max(initial=0)) < (1 << 30)) else np.int64, copy=False)
    n = c.shape[0]
    if n == 0:
        return ""
    x = c.copy()
    if n > 3:
        x[3:] -= c[1:-2]
    # pass k emits the k-th character of every value that still has one.  Nearly all deltas of a mask fit one
    # character, so the first pass is plain whole-array arithmetic and the later passes run on the small remainder.
    def one_pass(v):
        low = v & 0x1F
        v = v >> 5                                   # arithmetic shift, like the C code on a signed long
        done = v == -((low >> 4) & 1)                # sign bit set: finished when the rest is -1, else when it is 0
        return v, done, (low | ((~done).astype(v.dtype) << 5)) + 48

    x, done, ch0 = one_pass(x)
    ch0 = ch0.astype(np.uint8)
    if done.all():
        return ch0.tobytes().decode("ascii")
    idx = np.flatnonzero(~done)
    x = x[idx]
    n_chars = np.ones(n, dtype=np.int64)
    groups = []          # later passes: (indices of the values that emit a character, the char codes)
    while idx.size:
        x, done, ch = one_pass(x)
        groups.append((idx, ch.astype(np.uint8)))
        n_chars[idx] += 1
        keep = ~done
        idx, x = idx[keep], x[keep]
    starts = np.cumsum(n_chars) - n_chars
    out = np.empty(int(n_chars.sum()), dtype=np.uint8)
    out[starts] = ch0
    for k, (ii, ch) in enumerate(groups):
        out[starts[ii] + k + 1] = ch
    return out.tobytes().decode("ascii")


_PYCOCO = None


def _have_pycocotools() -> bool:
    """Probed once: a failing `import pycocotools` walks every sys.path entry again on each attempt (~0.2 ms per image)."""
    global _PYCOCO
    if _PYCOCO is None:
        try:
            from pycocotools import mask as mask_utils  # type: ignore  # noqa: F401

            _PYCOCO = True
        except ImportError:
            _PYCOCO = False
    return _PYCOCO


def coco_encode_rles(rles: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """coco_encode_rle for a whole list in ONE library call (csam_coco_rle_strings, host C): the per-mask interpreter
    overhead of hundreds of instances per crowd image disappears.  Same strings as `_coco_string` / pycocotools."""
    if not rles:
        return []
    if _have_pycocotools():
        return [coco_encode_rle(r) for r in rles]       # the reference's own dependency, when it is installed
    import ctypes as C

    from . import lib as L

    cnt = [np.ascontiguousarray(r["counts"], dtype=np.int32) for r in rles]
    flat = np.concatenate(cnt) if len(cnt) > 1 else cnt[0]
    offs = np.zeros(len(cnt) + 1, dtype=np.int64)
    np.cumsum([c.size for c in cnt], out=offs[1:])
    out = np.empty(max(7 * int(flat.size), 1), dtype=np.uint8)
    out_offs = np.empty(len(cnt) + 1, dtype=np.int64)
    L.check(L.load().csam_coco_rle_strings(flat.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), len(cnt),
                                           out.ctypes.data_as(C.c_void_p), out.size, out_offs.ctypes.data_as(C.c_void_p)),
            "csam_coco_rle_strings")
    buf = out[: int(out_offs[-1])].tobytes().decode("ascii")
    return [{"size": [int(r["size"][0]), int(r["size"][1])], "counts": buf[int(out_offs[i]):int(out_offs[i + 1])]}
            for i, r in enumerate(rles)]


def coco_encode_rle(uncompressed_rle: Dict[str, Any]) -> Dict[str, Any]:
    """Uses pycocotools when importable (as the reference does), else the built-in encoder."""
    h, w = uncompressed_rle["size"]
    try:
        from pycocotools import mask as mask_utils  # type: ignore

        rle = mask_utils.frPyObjects({"size": [h, w], "counts": [int(c) for c in uncompressed_rle["counts"]]}, h, w)
        rle["counts"] = rle["counts"].decode("utf-8")
        return rle
    except ImportError:
        return {"size": [h, w], "counts": _coco_string(uncompressed_rle["counts"])}


def remove_small_regions(mask: np.ndarray, area_thresh: float, mode: str):
    """Host-side OpenCV connected components exactly as the reference (out of the CUDA scope)."""
    import cv2  # type: ignore

    assert mode in ["holes", "islands"]
    holes = mode == "holes"
    n, regions, stats, _ = cv2.connectedComponentsWithStats((holes ^ mask).astype(np.uint8), 8)
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


def uncrop_boxes_xyxy(boxes: torch.Tensor, crop_box: List[int]) -> torch.Tensor:
    x0, y0 = crop_box[0], crop_box[1]
    off = torch.tensor([[x0, y0, x0, y0]], device=boxes.device)
    return boxes + (off.unsqueeze(1) if boxes.dim() == 3 else off)


def uncrop_points(points: torch.Tensor, crop_box: List[int]) -> torch.Tensor:
    off = torch.tensor([[crop_box[0], crop_box[1]]], device=points.device)
    return points + (off.unsqueeze(1) if points.dim() == 3 else off)
