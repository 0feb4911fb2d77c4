"""`crowdsam.utils` names that the reference's callers import (tools/test.py:8-10, batch_eval.py:7),
restricted to the hot-path subset of crowdsam/utils.py (config :31-58, image/box helpers :141-223,
dataset loading :370-390).  Drawing / evaluation helpers are out of scope."""
from __future__ import annotations

import os
import time
from datetime import datetime
from typing import List

import numpy as np
import torch
import yaml

from crowdsam_b200.pipeline import resize_image, _near_crop_edge  # noqa: F401

data_meta = {"crowdhuman": ["./datasets/crowdhuman", 1, {1: "person"}],
             "occhuman": ["./datasets/OCHuman", 1, {1: "person"}]}


def load_config(config_file):
    with open(config_file, "r") as f:
        return yaml.safe_load(f)


def convert_value(value: str):
    if value.lower() in ("true", "false"):
        return value.lower() == "true"
    for cast in (int, float):
        try:
            return cast(value)
        except ValueError:
            pass
    return value


def modify_config(config, options: List[str]):
    """Positional `a.b.c value` overrides (tools/test.py:23,27-28)."""
    assert len(options) % 2 == 0
    for key, value in zip(options[0::2], options[1::2]):
        node = config
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = convert_value(value)
    return config


def setup_logger(save_path, quiet=False):
    from loguru import logger
    import sys

    logger.remove()
    os.makedirs(save_path, exist_ok=True)
    stamp = datetime.fromtimestamp(time.time()).strftime("%Y-%m-%d %H:%M:%S")
    logger.add(f"{save_path}/{stamp}.log", format="{time}-{level}-{message}", filter="my_module", retention="10 days",
               level="DEBUG")
    logger.add(sys.stdout, format="{time}-{level}-{message}", filter="my_module", level="INFO")
    return logger


def load_img_and_annotation(dataset_path, annots, dataset, id=0):
    import cv2

    meta = annots["images"][id]
    sub = {"crowdhuman": "Images", "coco": "val2017", "occhuman": "images", "mineapple": "images"}
    if dataset == "coco_occ":
        path = os.path.join(dataset_path, "occ2017", meta["file_name"].split("/")[-1])
    elif dataset in sub:
        path = os.path.join(dataset_path, sub[dataset], meta["file_name"])
    else:
        raise NotImplementedError
    image = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
    boxes = np.array([a["bbox"] for a in annots["annotations"] if a["image_id"] == meta["id"]])
    boxes[..., 2:] += boxes[..., :2]
    return image, boxes, meta["id"]


def uncrop_boxes_xyxy(boxes: torch.Tensor, crop_box, downscale: float) -> torch.Tensor:
    x0, y0 = crop_box[0], crop_box[1]
    off = torch.tensor([[x0, y0, x0, y0]], device=boxes.device)
    return boxes / downscale + (off.unsqueeze(1) if boxes.dim() == 3 else off)


def uncrop_points(points: torch.Tensor, crop_box, downscale: float) -> torch.Tensor:
    off = torch.tensor([[crop_box[0], crop_box[1]]], device=points.device)
    return points / downscale + (off.unsqueeze(1) if points.dim() == 3 else off)


def is_box_near_crop_edge(boxes, crop_box, orig_box, downscale, atol: float = 20.0):
    return _near_crop_edge(boxes, crop_box, orig_box, downscale, atol)


def mask_iou_nms(boxes, scores, mask_preds, threshold):
    """crowdsam/utils.py:422-459 (dead code in the reference) on the K-MIOU kernel: greedy by score,
    suppress when max(inter/area_i, inter/area_j) > threshold on 150x150 nearest-resized masks."""
    from crowdsam_b200 import ops

    if mask_preds.numel() == 0:
        return []
    inter, area = ops.mask_overlap(mask_preds.to(torch.bool))
    inter, area = inter.cpu().numpy().astype(np.float32), area.cpu().numpy().astype(np.float32)
    order = np.argsort(-np.asarray(scores)).tolist()
    keep: List[int] = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in order:
            if keep:
                cov = np.maximum(inter[i, keep] / area[i], inter[i, keep] / area[keep])
                if np.any(cov > threshold):
                    continue
            keep.append(i)
    return np.array(keep)


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(f"crowdsam.utils.{name} is a visualisation / evaluation helper outside the B200 hot path")

    f.__name__ = name
    return f


visualize_result = _out_of_scope("visualize_result")
evaluate_boxes = _out_of_scope("evaluate_boxes")
