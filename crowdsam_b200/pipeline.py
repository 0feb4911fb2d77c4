"""CrowdSAM pipeline (crowdsam/model.py) over the B200 kernels.

Same constructor (`CrowdSAM(config, logger)`), same `generate(image) -> MaskData` contract and the
same selection semantics, restructured so that per-prompt work stays on the device:
  decode (P prompts) -> PWD score + candidate select -> K-POST stats (stability, box) ->
  filters -> K-POST write (bool masks of survivors only) -> K-NMS -> RLE kernel.
The reference materialises [P,4,H,W] fp32 twice per batch (model.py:344-384); here nothing of that
size is ever written.  Host syncs per batch: one small D2H of the keep flags (+ occupancy flags).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import amg, ops
from .amg import MaskData


def resize_image(image: np.ndarray, max_size: int):
    """crowdsam/utils.py:141-156: scale so the longer side is max_size (may up-scale), cv2 bilinear."""
    import cv2

    h0, w0 = image.shape[:2]
    r = min(max_size / w0, max_size / h0)
    h, w = int(r * h0), int(r * w0)
    if (h, w) == (h0, w0):
        return image, r          # cv2.resize to the same size is the identity: skip the 3 MB copy
    return cv2.resize(image, (w, h)), r


def _near_crop_edge(boxes: torch.Tensor, crop_box, orig_box, downscale, atol: float = 20.0) -> torch.Tensor:
    """crowdsam/utils.py:213-223 on a tiny [n,4] tensor."""
    cb = torch.as_tensor(crop_box, dtype=torch.float, device=boxes.device)
    ob = torch.as_tensor(orig_box, dtype=torch.float, device=boxes.device)
    x0, y0 = crop_box[0], crop_box[1]
    b = (boxes / downscale + torch.tensor([[x0, y0, x0, y0]], device=boxes.device)).float()
    near_c = torch.isclose(b, cb[None, :], atol=atol, rtol=0)
    near_i = torch.isclose(b, ob[None, :], atol=atol, rtol=0)
    return torch.any(near_c & ~near_i, dim=1)


class CrowdSAM:
    vis_img_id = 0

    def __init__(self, config, logger=None, predictor=None):
        """config: the reference YAML dict (configs/crowdhuman.yaml).  `predictor` may be injected
        (tests / benchmarks with synthetic weights); otherwise checkpoints are loaded as the reference does
        (model.py:33-42,88-115)."""
        self.device = torch.device(config["environ"]["device"])
        self.train_free = False
        t = config["test"]
        if predictor is None:
            predictor = self.load_sam_model(config["model"]["sam_model"], config["model"]["sam_arch"],
                                            config["model"]["sam_checkpoint"],
                                            config["model"]["sam_adapter_checkpoint"],
                                            self._load_dino(config), config["model"]["n_class"])
        self.predictor = predictor
        for k in ("mask_selection", "apply_box_offsets", "max_prompts", "filter_thresh", "max_size", "grid_size",
                  "pred_iou_thresh", "fuse_simmap", "stability_score_thresh", "stability_score_offset",
                  "box_nms_thresh", "points_per_batch", "crop_n_layers", "crop_nms_thresh", "crop_overlap_ratio",
                  "min_mask_region_area", "pos_sim_thresh", "output_rles"):
            setattr(self, k, t[k])
        # extension: `test.nms_mode: "mask_iou"` replaces the per-crop box NMS by the reference's
        # mask-overlap NMS (crowdsam/utils.py:422-459, dead code upstream) on K-MIOU; default "box" = the reference path
        self.nms_mode = t.get("nms_mode", "box")
        if self.nms_mode not in ("box", "mask_iou"):
            raise NotImplementedError(f"unknown test.nms_mode {self.nms_mode!r}")
        if config["model"].get("trainfree"):
            raise NotImplementedError("trainfree mode is outside the B200 hot path")
        if self.fuse_simmap or self.apply_box_offsets:
            raise NotImplementedError("fuse_simmap / apply_box_offsets are outside the B200 hot path")

    def _load_dino(self, config):
        from .modules import DinoVisionTransformer
        from .spec import DINO_ARCHS

        D, depth, heads = DINO_ARCHS[config["model"]["dino_model"]]
        dino = DinoVisionTransformer(D, depth, heads)
        dino.load_state_dict(torch.load(config["model"]["dino_checkpoint"], map_location="cpu"))
        return dino.to(self.device)

    def load_sam_model(self, sam_model, sam_arch, sam_checkpoint, sam_adapter_checkpoint, dino_model, n_class):
        if sam_arch != "crowdsam":
            raise NotImplementedError(f"sam_arch {sam_arch!r} needs packages the reference does not ship")
        from .build import sam_model_registry
        from .predictor import SamPredictor

        sam = sam_model_registry[sam_model](checkpoint=sam_checkpoint, n_class=n_class)
        sam.mask_decoder.load_state_dict(torch.load(sam_adapter_checkpoint, map_location="cpu"), strict=False)
        sam = sam.to(self.device)
        return SamPredictor(sam, dino_model.to(self.device))

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def generate(self, image) -> MaskData:
        return self._generate_masks(image)

    # ------------------------------------------------------------------ internals
    def crop_image(self, image, crop_box):
        x0, y0, x1, y1 = crop_box
        if not isinstance(image, np.ndarray):
            image = np.array(image, dtype=np.uint8)
        self.orig_image = image
        self.image, self.downscale = resize_image(image[y0:y1, x0:x1, :], self.max_size)

    def _generate_masks(self, image) -> MaskData:
        img_size = np.asarray(image).shape[:2]          # (a view for ndarray input; the reference copies 3 MB here)
        crop_boxes, _ = amg.generate_crop_boxes(img_size, self.crop_n_layers, self.crop_overlap_ratio)
        data = MaskData()
        for crop_box in crop_boxes:
            crop_data = self._process_crop(image, crop_box)
            if crop_data is not None:
                if len(crop_boxes) > 1 and "rles_info" in crop_data._stats:
                    # The reference appends every crop's 2-element [crop_box, [H, W]] list (model.py:293) and then
                    # indexes the concatenation with the cross-crop NMS keep indices (amg.py:55): IndexError as soon
                    # as a kept index exceeds 2 x n_crops.  Here multi-crop runs keep one [crop_box, [H, W]] entry
                    # per detection, which filters like every other column; single-crop output is unchanged.
                    crop_data["rles_info"] = [list(crop_data["rles_info"]) for _ in range(len(crop_data["boxes"]))]
                data.cat(crop_data)
        if len(crop_boxes) > 1 and "crop_boxes" in data._stats and len(data["crop_boxes"]) > 0:
            cb = data["crop_boxes"].float()
            scores = (1.0 / ((cb[:, 2] - cb[:, 0]) * (cb[:, 3] - cb[:, 1]))).to(data["boxes"].device)
            keep = ops.box_nms(data["boxes"].float(), scores, self.crop_nms_thresh)
            data.filter(keep)
            del data["crop_boxes"]
        if len(data._stats.keys()) > 0:
            del data["iou_preds"]
        else:
            data["boxes"] = torch.zeros(0, 4)
            data["scores"] = torch.zeros(0, 4)
        data["rles"] = amg.coco_encode_rles(data["rles"]) if "rles" in data._stats else []
        data.to_numpy()
        return data

    def candidate_points(self, img_hw) -> np.ndarray:
        """Foreground prior -> prompt candidates (model.py:196-223,445-449)."""
        G = self.grid_size
        img_size = torch.tensor(img_hw)
        feat_size = (img_size * min(G / img_size)).int()
        sim = self.predictor.predict_fg_map(img_size)                       # [1,n_class,256,256]
        sim = ops.bilinear(sim[0], G, G, chlast=False)                       # model.py:202
        sim = sim.sigmoid().max(dim=0)[0][: int(feat_size[0]), : int(feat_size[1])]
        coords = (sim > self.pos_sim_thresh).nonzero()[:, [1, 0]].cpu()
        inv = torch.tensor([feat_size[1] / img_hw[1], feat_size[0] / img_hw[0]])
        return (coords / inv).numpy()

    def _process_crop(self, image, crop_box) -> Optional[MaskData]:
        self.crop_image(image, crop_box)
        self.predictor.set_image(self.image)
        return self._run_prompts(crop_box)

    @torch.no_grad()
    def run_resident(self, img_u8_chw: torch.Tensor, encode_rle: bool = False) -> Optional[MaskData]:
        """Device-resident variant of generate() for a full-frame crop whose uint8 CHW image already
        lives in HBM at the model resolution (long side 1024): no host resize, no H2D of pixels and, unless
        asked, no RLE/D2H of masks.  Used by bench.py for the `value` leg; same kernels as generate()."""
        h, w = int(img_u8_chw.shape[1]), int(img_u8_chw.shape[2])
        self.orig_image = np.empty((h, w, 0), dtype=np.uint8)
        self.image, self.downscale = self.orig_image, 1.0
        self.predictor.set_torch_image(img_u8_chw[None], (h, w))
        return self._run_prompts([0, 0, w, h], encode_rle=encode_rle)

    def _run_prompts(self, crop_box, encode_rle: bool = True) -> Optional[MaskData]:
        orig_h, orig_w = self.orig_image.shape[:2]
        self.last_counts = (0, 0)
        points = self.candidate_points(self.image.shape[:2])
        parts = []
        # EPS iterator (model.py:229-248): same RNG consumption.  Which prompts form the next batch depends on this
        # batch's masks, so one host decision per batch is inherent; everything else stays on the device: the keep
        # list is compacted there, the survivors' masks are written there, the occupancy of the remaining candidates is
        # tested there, and ONE device-to-host read per batch brings back the keep flags and the occupancy flags.
        points = points.astype("int")
        np.random.shuffle(points)
        count, bs = 0, self.points_per_batch
        while len(points) > 0 and count < self.max_prompts:
            bs = min(len(points), bs)
            sel, points = points[:bs], points[bs:]
            batch, occ = self._process_batch(sel, self.predictor.original_size, crop_box, rest=points)
            if occ is not None:
                points = points[~occ]
            parts.append(batch)
            count += bs
        data = MaskData.merged(parts)
        self.predictor.reset_image()
        if len(data.items()) == 0 or len(data["masks"]) == 0:
            return None
        n_into_nms = len(data["boxes"])
        if self.nms_mode == "mask_iou":
            from .dropin.crowdsam.utils import mask_iou_nms

            kept = mask_iou_nms(data["boxes"], data["iou_preds"].cpu().numpy(), data["masks"], self.box_nms_thresh)
            keep = torch.as_tensor(np.asarray(kept, dtype=np.int64), device=data["boxes"].device)
        else:
            keep = ops.box_nms(data["boxes"].float(), data["iou_preds"], self.box_nms_thresh)   # model.py:257-263
        data.filter(keep)
        if self.min_mask_region_area > 0:
            data = self.postprocess_small_regions(data, self.min_mask_region_area,
                                                  max(self.box_nms_thresh, self.crop_nms_thresh))
        data["scores"] = data["iou_preds"]
        self.last_counts = (n_into_nms, len(data["boxes"]))          # N masks into NMS, K detections kept
        if not encode_rle:
            return data
        data["rles"] = amg.mask_to_rle_arrays(data["masks"])
        data["rles_info"] = [crop_box, [orig_h, orig_w]]
        del data["masks"]
        x0, y0 = crop_box[0], crop_box[1]
        dev = data["boxes"].device
        data["boxes"] = data["boxes"] / self.downscale + torch.tensor([[x0, y0, x0, y0]], device=dev)
        data["points"] = data["points"] / self.downscale + torch.tensor([[x0, y0]])
        data["crop_boxes"] = torch.tensor([crop_box for _ in range(len(data["boxes"]))])
        data["fboxes"] = data["boxes"]
        return data

    def _process_batch(self, points: np.ndarray, im_size, crop_box, rest: Optional[np.ndarray] = None):
        """model.py:334-390 with the selection stages fused on the device.  `rest` = the candidate points still waiting
        (EPS): when given, also returns which of them this batch's confident masks occupy (model.py:240,246).
        -> (MaskData of the batch, occupancy bool [len(rest)] or None)."""
        pr = self.predictor
        thr = float(pr.model.mask_threshold)
        coords = torch.as_tensor(pr.transform.apply_coords(points, im_size))[:, None, :]
        labels = torch.ones(coords.shape[0], dtype=torch.int)[:, None]
        low, iou, cls = pr.decode_low_res(coords, labels)
        if cls.shape[-1] != 1:
            raise NotImplementedError("n_class > 1 breaks the reference itself (model.py:351 squeeze)")
        if self.mask_selection == "max_iou":
            score, sel, cat = ops.select_candidates(iou, cls)
        elif self.mask_selection in ("max_area", "min_area"):
            cnt4, _ = ops.mask_post_stats(low.reshape(-1, 256, 256), None, pr.input_size, pr.original_size, thr, 0.0)
            area = cnt4[:, 2].view(-1, 4)
            sel = (area.max(dim=-1)[1] if self.mask_selection == "max_area" else area.min(dim=-1)[1]).to(torch.int32)
            s4 = torch.clamp(iou, 0.0) * cls.squeeze(2).sigmoid()
            score = s4.gather(1, sel.long()[:, None])[:, 0]
            cat = torch.zeros_like(sel)
        else:
            raise NotImplementedError
        counts, boxes = ops.mask_post_stats(low, sel, pr.input_size, pr.original_size, thr, float(self.stability_score_offset))
        stab = counts[:, 0] / counts[:, 1]                     # int32 / int32 -> fp32 (amg.py:176)
        keep = torch.ones_like(score, dtype=torch.bool)
        if self.pred_iou_thresh > 0.0:
            keep &= score > self.pred_iou_thresh
        if self.stability_score_thresh > 0.0:
            keep &= stab >= self.stability_score_thresh
        orig_h, orig_w = self.orig_image.shape[:2]
        keep &= ~_near_crop_edge(boxes, crop_box, [0, 0, orig_w, orig_h], self.downscale)
        pts_t = torch.as_tensor(points)
        if rest is None or len(rest) == 0:
            idx = keep.nonzero()[:, 0]                          # the one host sync of the batch
            masks, _ = ops.mask_post_write(low, sel, idx.to(torch.int32), pr.input_size, pr.original_size, thr)
            idx_c, occ = idx.cpu(), None
        else:
            # more candidates wait: compact the keep list on the device (kept prompts first, in prompt order, -1 for the
            # rest), write the survivors' masks, test the remaining candidates against the masks whose score exceeds
            # filter_thresh (model.py:240,246), then read keep flags + occupancy flags back together
            bs = keep.shape[0]
            order = torch.argsort((~keep).to(torch.uint8), stable=True)
            kept_sorted = keep[order]
            klist = torch.where(kept_sorted, order, torch.full_like(order, -1)).to(torch.int32)
            masks_all, _ = ops.mask_post_write(low, sel, klist, pr.input_size, pr.original_size, thr)
            flag = (kept_sorted & (score[order] > self.filter_thresh)).to(torch.uint8)
            pts_d = torch.as_tensor(rest, dtype=torch.int32, device=self.device)
            occ_d = ops.points_occupied(masks_all, flag, pts_d)
            both = torch.cat([keep.to(torch.uint8), occ_d]).cpu().numpy().astype(bool)     # the one host sync
            n_keep = int(both[:bs].sum())
            idx_c = torch.as_tensor(np.flatnonzero(both[:bs]))
            idx, masks, occ = order[:n_keep], masks_all[:n_keep], both[bs:]
        data = MaskData(masks=masks, iou_preds=score[idx], points=pts_t[idx_c], categories=cat[idx].long(),
                        stability_score=stab[idx], boxes=boxes[idx].long())
        return data, occ

    @staticmethod
    def postprocess_small_regions(mask_data: MaskData, min_area: int, nms_thresh: float) -> MaskData:
        """model.py:395-443 with the connected-component passes on the device (csam_remove_small_regions, integer
        exact against the cv2.connectedComponentsWithStats path of the reference): no per-mask D2H copy, no host loop."""
        if len(mask_data["masks"]) == 0:
            return mask_data
        old = mask_data["masks"]
        masks = old.to(torch.uint8).contiguous().clone()
        # the reference compares `size < area_thresh` with the config value as is (float or int)
        thr = int(np.ceil(min_area))
        c1 = ops.remove_small_regions(masks, thr, "holes")
        c2 = ops.remove_small_regions(masks, thr, "islands")
        scores = ((c1 | c2) == 0).float()                      # unchanged masks win the NMS (model.py:425-434)
        new_masks = masks.to(old.dtype)
        boxes = amg.batched_mask_to_box(new_masks.bool())
        keep = ops.box_nms(boxes.float(), scores, nms_thresh)
        upd = keep[scores[keep] == 0.0]
        if upd.numel() > 0:
            mask_data["boxes"][upd] = boxes[upd].to(mask_data["boxes"].dtype)
            mask_data["masks"][upd] = new_masks[upd]
        mask_data.filter(keep)
        return mask_data
