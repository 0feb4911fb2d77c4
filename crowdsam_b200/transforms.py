"""Host-side geometry helpers of the predictor boundary (segment_anything_cs/utils/transforms.py)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


class ResizeLongestSide:
    """Scale so the longer image side equals `target_length` (transforms.py:16-102).  Image resizing
    stays on the host with PIL bilinear exactly as the reference does; coordinates are float64."""

    def __init__(self, target_length: int) -> None:
        self.target_length = int(target_length)

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        s = long_side_length * 1.0 / max(oldh, oldw)
        return int(oldh * s + 0.5), int(oldw * s + 0.5)

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        from PIL import Image

        nh, nw = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        if (nh, nw) == tuple(image.shape[:2]):
            return np.ascontiguousarray(image)
        return np.array(Image.fromarray(image).resize((nw, nh), Image.BILINEAR))

    def _scales(self, original_size):
        oh, ow = original_size
        nh, nw = self.get_preprocess_shape(oh, ow, self.target_length)
        return nw / ow, nh / oh

    def apply_coords(self, coords: np.ndarray, original_size) -> np.ndarray:
        sx, sy = self._scales(original_size)
        out = np.array(coords, dtype=float, copy=True)
        out[..., 0] *= sx
        out[..., 1] *= sy
        return out

    def apply_boxes(self, boxes: np.ndarray, original_size) -> np.ndarray:
        return self.apply_coords(np.asarray(boxes).reshape(-1, 2, 2), original_size).reshape(-1, 4)

    def apply_coords_torch(self, coords: torch.Tensor, original_size) -> torch.Tensor:
        sx, sy = self._scales(original_size)
        out = coords.clone().to(torch.float)
        out[..., 0] *= sx
        out[..., 1] *= sy
        return out

    def apply_boxes_torch(self, boxes: torch.Tensor, original_size) -> torch.Tensor:
        return self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size).reshape(-1, 4)

    def apply_image_torch(self, image: torch.Tensor) -> torch.Tensor:
        nh, nw = self.get_preprocess_shape(image.shape[2], image.shape[3], self.target_length)
        return torch.nn.functional.interpolate(image, (nh, nw), mode="bilinear", align_corners=False, antialias=True)
