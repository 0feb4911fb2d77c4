"""Generate tests/golden/*.npz by running the REAL reference from /root/reference on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed
here as small (sub-sampled) fixtures; tests/test_oracle_golden.py pins oracle/restate.py
against them, and the GPU parity tests then use the oracle as the travelling checker.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import fixtures, ref_import, weights  # noqa: E402
from oracle.restate import DEFAULT_TEST_CFG  # noqa: E402

torch.set_grad_enabled(False)


def sub(t, *steps):
    sl = tuple(slice(None, None, s) for s in steps)
    return np.ascontiguousarray(t[sl].numpy() if isinstance(t, torch.Tensor) else t[sl])


def model_case(name, sam_arch, dino_arch):
    """Encoder / DINOv2 / decoder / predict_torch outputs of the real modules."""
    sam_sd = weights.make_sam_state(sam_arch)
    dino_sd = weights.make_dino_state(dino_arch)
    sam = ref_import.build_sam(sam_sd, sam_arch)
    dino = ref_import.build_dino(dino_sd, dino_arch)
    sacs, _, _ = ref_import.load()
    pred = sacs.SamPredictor(sam, dino)
    out = {}
    # ---- square image: set_image + decoder on 6 prompts
    img = weights.synthetic_image(0)
    pred.set_image(img)
    out["features"] = sub(pred.features, 1, 8, 2, 2)
    out["dino_feats"] = sub(pred.dino_feats, 1, 6, 6, 8)
    out["fg_map"] = sub(pred.predict_fg_map(), 1, 1, 4, 4)
    pts = np.array([[10, 20], [512, 512], [1000, 30], [333, 777], [64, 960], [800, 801]])
    coords = torch.as_tensor(pred.transform.apply_coords(pts, pred.original_size))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    masks, iou, cls, low = pred.predict_torch(coords, labels, multimask_output=True, return_logits=True)
    out["points"] = pts
    out["low_res"] = sub(low, 1, 1, 8, 8)
    out["masks"] = sub(masks, 1, 1, 32, 32)
    out["iou_pred"] = iou.numpy()
    out["cls"] = cls.numpy()
    out["dense_pe"] = sub(sam.prompt_encoder.get_dense_pe(), 1, 4, 4, 4)
    # ---- non-square image (exercises pad + both postprocess resizes)
    img2 = weights.synthetic_image(1, 600, 900)
    pred.set_image(img2)
    pts2 = np.array([[5, 5], [450, 300], [880, 590]])
    coords2 = torch.as_tensor(pred.transform.apply_coords(pts2, pred.original_size))[:, None, :]
    masks2, iou2, cls2, low2 = pred.predict_torch(coords2, labels[:3], multimask_output=True, return_logits=True)
    out["ns_points"] = pts2
    out["ns_features"] = sub(pred.features, 1, 8, 2, 2)
    out["ns_low_res"] = sub(low2, 1, 1, 8, 8)
    out["ns_masks"] = sub(masks2, 1, 1, 24, 36)
    out["ns_iou_pred"] = iou2.numpy()
    out["ns_cls"] = cls2.numpy()
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})
    return sam, dino


def pipeline_case(name, sam, dino, overrides, image_index=0, hw=(1024, 1024)):
    """CrowdSAM.generate end to end through the real crowdsam/model.py."""
    cfg = dict(DEFAULT_TEST_CFG)
    cfg.update(overrides)
    m = ref_import.build_crowdsam(sam, dino, cfg)
    img = weights.synthetic_image(image_index, *hw)
    np.random.seed(42)
    res = m.generate(img)
    out = {"cfg_keys": np.array(sorted(overrides.keys())),
           "cfg_vals": np.array([str(overrides[k]) for k in sorted(overrides.keys())]),
           "image_index": np.array(image_index), "hw": np.array(hw)}
    for k, v in res.items():
        if k == "rles":
            out["rle_counts"] = np.array([r["counts"] for r in v])
            out["rle_sizes"] = np.array([r["size"] for r in v]).reshape(-1, 2)
        elif k == "rles_info":
            out["rles_info"] = np.array([list(v[0]), list(v[1]) + [0, 0]])
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, f"pipeline_{name}.npz"), **out)
    print("wrote pipeline", name, {k: (v.shape, v.dtype) for k, v in out.items()})


def stage_case():
    """Post-processing + NMS + RLE through the reference's own utility functions on seeded
    blob logits (the stage-level fixture of SURVEY.md §8d)."""
    ref_import.load()
    from segment_anything_cs.utils import amg
    from segment_anything_cs.modeling.sam import Sam
    from torchvision.ops import nms as tv_nms
    from torchvision.ops.boxes import batched_nms

    out = {}
    P = 48
    low, iou, cls = fixtures.blob_logits(P, seed=0)

    class _Enc:
        img_size = 1024

    class _S:
        image_encoder = _Enc()

    for tag, (inp, orig) in {"sq": ((1024, 1024), (1024, 1024)), "ns": ((683, 1024), (600, 900))}.items():
        full = Sam.postprocess_masks(_S(), low, inp, orig)
        score = torch.clamp(iou, 0.0) * cls.squeeze(2).sigmoid()
        sel = score.max(dim=-1)[1]
        m = full[torch.arange(P), sel]
        stab = amg.calculate_stability_score(m, 0.0, 1.0)
        binm = m > 0.0
        boxes = amg.batched_mask_to_box(binm)
        out[f"{tag}_score"] = score[torch.arange(P), sel].numpy()
        out[f"{tag}_sel"] = sel.numpy()
        out[f"{tag}_stability"] = stab.numpy()
        out[f"{tag}_boxes"] = boxes.numpy()
        out[f"{tag}_area"] = binm.flatten(1).sum(1).numpy()
        keep = batched_nms(boxes.float(), score[torch.arange(P), sel], torch.zeros_like(boxes[:, 0]), 0.65)
        out[f"{tag}_nms_keep"] = keep.numpy()
        if tag == "sq":
            rles = amg.mask_to_rle_pytorch(binm[:4])
            out["sq_rle_counts0"] = np.array(rles[0]["counts"])
            out["sq_rle_counts3"] = np.array(rles[3]["counts"])
            out["sq_rle_str0"] = np.array(amg.coco_encode_rle(rles[0])["counts"])
    # plain NMS goldens (torchvision.ops.nms, CPU = stable order)
    for n, seed, binary in ((257, 0, False), (3000, 1, False), (3000, 2, True), (1, 3, False)):
        b, s = fixtures.random_boxes(n, seed, binary_scores=binary)
        for thr in (0.65, 0.7):
            keep = tv_nms(torch.as_tensor(b), torch.as_tensor(s), thr)
            out[f"nms_{n}_{seed}_{thr}"] = keep.numpy()
    np.savez_compressed(os.path.join(HERE, "stage_post_nms.npz"), **out)
    print("wrote stage", {k: v.shape for k, v in out.items()})


def config0_case():
    """BASELINE.json configs[0] at FULL depth: SAM ViT-B (12 blocks) + DINOv2 ViT-L/14 (24 blocks) through the real
    reference on CPU, 1024x1024 synthetic image, 8x8 prompt grid -> model_vit_b.npz, pipeline_vit_b_grid8.npz."""
    sam, dino = model_case("vit_b", "vit_b", "dinov2_vitl14")
    pipeline_case("vit_b_grid8", sam, dino,
                  dict(grid_size=8, pos_sim_thresh=-1, max_prompts=64, points_per_batch=32,
                       filter_thresh=2.0, min_mask_region_area=0))


def config1_case():
    """BASELINE.json configs[1] (the headline workload) at full depth: SAM ViT-L + DINOv2 ViT-L/14 through the real
    reference on CPU, 1024x1024 synthetic image, 32x32 prompt grid = 1024 prompts, PWD-Net scoring + filters + NMS
    -> model_vit_l.npz, pipeline_vit_l_grid32.npz (about 5 minutes on 8 cores)."""
    sam, dino = model_case("vit_l", "vit_l", "dinov2_vitl14")
    pipeline_case("vit_l_grid32", sam, dino,
                  dict(grid_size=32, pos_sim_thresh=-1, max_prompts=1024, points_per_batch=64,
                       filter_thresh=2.0, min_mask_region_area=0))
    # configs[2]: 64x64 dense grid = 4096 prompts per image (about 5 more minutes)
    pipeline_case("vit_l_grid64", sam, dino,
                  dict(grid_size=64, pos_sim_thresh=-1, max_prompts=4096, points_per_batch=64,
                       filter_thresh=2.0, min_mask_region_area=0), image_index=1)


def config3_case():
    """BASELINE.json configs[3] (encoder-bound): SAM ViT-H (32 blocks, head dim 80) + DINOv2 ViT-L/14 through the real
    reference on CPU, 32x32 grid -> model_vit_h.npz, pipeline_vit_h_grid32.npz."""
    sam, dino = model_case("vit_h", "vit_h", "dinov2_vitl14")
    cfg = dict(grid_size=32, pos_sim_thresh=-1, max_prompts=1024, points_per_batch=64,
               filter_thresh=2.0, min_mask_region_area=0)
    pipeline_case("vit_h_grid32", sam, dino, cfg)                           # image 0: the reference finds nothing
    pipeline_case("vit_h_grid32_img3", sam, dino, cfg, image_index=3)       # image 3: one detection


# --------------------------------------------------------------------------------------------------
# Multi-detection end-to-end goldens: the REAL CrowdSAM pipeline on injected decoder outputs
# --------------------------------------------------------------------------------------------------
def _inject(pred, seed, log):
    """Replace the real predictor's predict_torch by the injected decoder (fixtures.injected_decoder_outputs)
    followed by the reference's own Sam.postprocess_masks, i.e. predictor.py:285-292 with the decoder swapped."""

    def predict_torch(point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                      return_logits=False, **kw):
        assert pred.is_image_set
        xy = point_coords[:, 0, :].cpu().numpy()
        log.append(xy.copy())
        low, iou, cls = fixtures.injected_decoder_outputs(xy, seed)
        masks = pred.model.postprocess_masks(low, pred.input_size, pred.original_size)
        if not return_logits:
            masks = masks > pred.model.mask_threshold
        return masks, iou, cls, low

    pred.predict_torch = predict_torch


def injected_pipeline_case(name, sam, dino, overrides, image_index=0, hw=(1024, 1024), seed=0, coordinate_trick=False):
    """CrowdSAM.generate of the real crowdsam/model.py (EPS iterator, select_mask, filters, boxes, crop-edge filter,
    per-crop NMS, postprocess_small_regions + second NMS, RLE, uncrop, cross-crop NMS) with the decoder outputs
    injected per prompt point -> many distinct instances instead of the single full-image box random weights give.
    coordinate_trick: route batched_nms through torchvision's coordinate-trick branch, the one the reference takes
    on CUDA for N <= 25000 boxes (on CPU it would switch to the vanilla branch above 1000 boxes, whose final
    non-stable sort orders tied scores differently)."""
    cfg = dict(DEFAULT_TEST_CFG)
    cfg.update(overrides)
    m = ref_import.build_crowdsam(sam, dino, cfg)
    log = []
    _inject(m.predictor, seed, log)
    _, cmodel, _ = ref_import.load()
    saved = cmodel.batched_nms
    if coordinate_trick:
        from torchvision.ops.boxes import _batched_nms_coordinate_trick

        cmodel.batched_nms = lambda b, s, i, iou_threshold: _batched_nms_coordinate_trick(b, s, i, iou_threshold)
    try:
        img = weights.synthetic_image(image_index, *hw)
        np.random.seed(42)
        res = m.generate(img)
    finally:
        cmodel.batched_nms = saved
    out = {"cfg_keys": np.array(sorted(overrides.keys())),
           "cfg_vals": np.array([str(overrides[k]) for k in sorted(overrides.keys())]),
           "image_index": np.array(image_index), "hw": np.array(hw), "inject_seed": np.array(seed),
           "call_sizes": np.array([len(x) for x in log]), "call_points": np.concatenate(log, 0)}
    for k, v in res.items():
        if k == "rles":
            out["rle_counts"] = np.array([r["counts"] for r in v])
            out["rle_sizes"] = np.array([r["size"] for r in v]).reshape(-1, 2)
        elif k == "rles_info":
            out["rles_info"] = np.array([list(v[0]), list(v[1]) + [0, 0]])
        else:
            out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, f"pipeline_inj_{name}.npz"), **out)
    print("wrote injected pipeline", name, "calls", len(log), "prompts", int(sum(len(x) for x in log)),
          "detections", len(res["boxes"]))


def injected_crops_case(sam, dino, overrides, image_index, hw, seed):
    """crop_n_layers = 1 (5 crops).  The reference's own generate() cannot finish this case: MaskData.cat appends
    the 2-element `rles_info` list of every crop (model.py:293) and MaskData.filter then indexes that list with the
    NMS keep indices (amg.py:55) -> IndexError as soon as a kept index exceeds 2 x n_crops [observed here].  So the
    golden pins what does run: the real `_process_crop` of every crop box of the real `generate_crop_boxes`
    (crop + resize to max_size, EPS, filters incl. the crop-edge filter, per-crop NMS, small regions, uncrop), and
    the reference's cross-crop NMS statement (model.py:167-176) applied to the concatenated crop outputs."""
    from torchvision.ops.boxes import batched_nms, box_area

    cfg = dict(DEFAULT_TEST_CFG)
    cfg.update(overrides)
    m = ref_import.build_crowdsam(sam, dino, cfg)
    log = []
    _inject(m.predictor, seed, log)
    sacs, _, _ = ref_import.load()
    from segment_anything_cs.utils.amg import generate_crop_boxes

    img = weights.synthetic_image(image_index, *hw)
    crop_boxes, _ = generate_crop_boxes(img.shape[:2], cfg["crop_n_layers"], cfg["crop_overlap_ratio"])
    np.random.seed(42)
    out = {"cfg_keys": np.array(sorted(overrides.keys())),
           "cfg_vals": np.array([str(overrides[k]) for k in sorted(overrides.keys())]),
           "image_index": np.array(image_index), "hw": np.array(hw), "inject_seed": np.array(seed),
           "crop_boxes_all": np.array(crop_boxes)}
    boxes, cbs, scores = [], [], []
    for ci, cb in enumerate(crop_boxes):
        with torch.no_grad():
            d = m._process_crop(img, cb)
        n = 0 if d is None else len(d["boxes"])
        out[f"crop{ci}_n"] = np.array(n)
        if d is None:
            continue
        for k in ("boxes", "points", "scores", "stability_score", "categories", "crop_boxes"):
            out[f"crop{ci}_{k}"] = d[k].cpu().numpy()
        out[f"crop{ci}_rle_counts"] = np.array([ref_import.coco_string(r) for r in d["rles"]])
        out[f"crop{ci}_rle_size"] = np.array(d["rles"][0]["size"]) if n else np.zeros(2, int)
        boxes.append(d["boxes"]); cbs.append(d["crop_boxes"]); scores.append(d["scores"])
    allb, allc = torch.cat(boxes), torch.cat(cbs)
    sc = 1 / box_area(allc)                                                       # model.py:169
    keep = batched_nms(allb.float(), sc, torch.zeros_like(allb[:, 0]), iou_threshold=cfg["crop_nms_thresh"])
    out["cross_keep"] = keep.numpy()
    out["cross_boxes"] = allb[keep].numpy()
    out["cross_scores"] = torch.cat(scores)[keep].numpy()
    out["call_sizes"] = np.array([len(x) for x in log])
    np.savez_compressed(os.path.join(HERE, "pipeline_inj_crops.npz"), **out)
    print("wrote injected crops: per-crop detections", [int(out[f"crop{i}_n"]) for i in range(len(crop_boxes))],
          "kept after cross-crop NMS", len(keep), "of", len(allb))


def extra_stage_case():
    """stage_extra.npz: (1) `mask_iou_nms` of the reference (crowdsam/utils.py:422-459, dead code there but callable)
    on overlapping instance masks; (2) torchvision nms with NaN / signed-zero scores; (3) the K-POST stage functions of
    the reference (postprocess_masks, calculate_stability_score, batched_mask_to_box) on injected logits at P = 64."""
    _, _, cutils = ref_import.load()
    from segment_anything_cs.utils import amg
    from segment_anything_cs.modeling.sam import Sam
    from torchvision.ops import nms as tv_nms

    out = {}
    # (1) 48 masks at 256x256 from the injected decoder around 12 cluster centres -> heavy overlap
    pts = fixtures.cluster_points(48, seed=7)
    low, iou, _ = fixtures.injected_decoder_outputs(pts, seed=7)
    masks = low[:, 2] > 0                                      # third candidate: the larger ones
    scores = iou[:, 2].numpy().copy()
    out["miou_points"] = pts
    for thr in (0.3, 0.5, 0.8):
        keep = cutils.mask_iou_nms(np.zeros((len(pts), 4)), scores, masks, thr)
        out[f"miou_keep_{thr}"] = np.asarray(keep)
    out["miou_empty"] = np.asarray(cutils.mask_iou_nms(np.zeros((0, 4)), np.zeros(0), masks[:0], 0.5))
    # (2) NaN / -0.0 scores (torch.sort: NaN first, -0.0 == +0.0, stable)
    b, sc = fixtures.random_boxes(300, 9)
    sc = sc.copy()
    sc[::17] = np.nan
    sc[5::23] = -0.0
    sc[6::23] = 0.0
    out["nan_scores"] = sc
    out["nan_keep"] = tv_nms(torch.as_tensor(b), torch.as_tensor(sc), 0.65).numpy()
    # (3) K-POST stage on injected logits, all four planes, P = 64, square and non-square geometry
    P = 64
    pts = fixtures.grid_points(8).astype(np.float64)
    low, iou, cls = fixtures.injected_decoder_outputs(pts, seed=11)

    class _Enc:
        img_size = 1024

    class _S:
        image_encoder = _Enc()

    for tag, (inp, orig) in {"sq": ((1024, 1024), (1024, 1024)), "ns": ((683, 1024), (600, 900))}.items():
        full = Sam.postprocess_masks(_S(), low, inp, orig).flatten(0, 1)
        out[f"p64_{tag}_stability"] = amg.calculate_stability_score(full, 0.0, 1.0).numpy()
        binm = full > 0.0
        out[f"p64_{tag}_boxes"] = amg.batched_mask_to_box(binm).numpy()
        out[f"p64_{tag}_area"] = binm.flatten(1).sum(1).numpy()
    np.savez_compressed(os.path.join(HERE, "stage_extra.npz"), **out)
    print("wrote stage_extra", {k: v.shape for k, v in out.items()})


def amg_case():
    """amg_inj.npz: the reference's SamAutomaticMaskGenerator, which cannot be constructed as shipped
    (`SamPredictor(model)` lacks dino_model, automatic_mask_generator.py:123) nor unpack predict_torch's 4 returns
    (:279).  Patched at RUN TIME only (no source copy): the module's `SamPredictor` name is bound to a subclass that
    supplies dino_model and returns the first three values; the decoder outputs are injected as for the CrowdSAM cases."""
    sacs, _, _ = ref_import.load()
    import segment_anything_cs.automatic_mask_generator as ramg

    sam_sd, dino_sd = weights.make_sam_state("tiny"), weights.make_dino_state("tiny")
    sam, dino = ref_import.build_sam(sam_sd, "tiny"), ref_import.build_dino(dino_sd, "tiny")
    seed = 21

    class _Pred(sacs.SamPredictor):
        def __init__(self, model):
            super().__init__(model, dino)

        def predict_torch(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                          return_logits=False, **kw):
            low, iou, cls = fixtures.injected_decoder_outputs(point_coords[:, 0, :].cpu().numpy(), seed)
            masks = self.model.postprocess_masks(low, self.input_size, self.original_size)
            if not return_logits:
                masks = masks > self.model.mask_threshold
            return masks, iou, cls

    saved = ramg.SamPredictor
    ramg.SamPredictor = _Pred
    out = {"inject_seed": np.array(seed)}
    try:
        for tag, hw, kw in (("sq", (1024, 1024), dict(min_mask_region_area=0)),
                            ("ns", (600, 900), dict(min_mask_region_area=100)),
                            ("crops", (600, 900), dict(min_mask_region_area=100, crop_n_layers=1,
                                                       crop_n_points_downscale_factor=2))):
            gen = ramg.SamAutomaticMaskGenerator(sam, points_per_side=12, points_per_batch=32, pred_iou_thresh=0.5,
                                                 stability_score_thresh=0.85, box_nms_thresh=0.7,
                                                 output_mode="coco_rle", **kw)
            img = weights.synthetic_image(6, *hw)
            recs = gen.generate(img)
            out[f"{tag}_hw"] = np.array(hw)
            out[f"{tag}_bbox"] = np.array([r["bbox"] for r in recs])
            out[f"{tag}_area"] = np.array([r["area"] for r in recs])
            out[f"{tag}_iou"] = np.array([r["predicted_iou"] for r in recs], dtype=np.float32)
            out[f"{tag}_stab"] = np.array([r["stability_score"] for r in recs], dtype=np.float32)
            out[f"{tag}_point"] = np.array([r["point_coords"][0] for r in recs])
            out[f"{tag}_crop_box"] = np.array([r["crop_box"] for r in recs])
            out[f"{tag}_rle"] = np.array([r["segmentation"]["counts"] for r in recs])
            print("amg", tag, len(recs), "records")
    finally:
        ramg.SamPredictor = saved
    np.savez_compressed(os.path.join(HERE, "amg_inj.npz"), **out)


def injected_cases():
    sam_sd, dino_sd = weights.make_sam_state("tiny"), weights.make_dino_state("tiny")
    sam, dino = ref_import.build_sam(sam_sd, "tiny"), ref_import.build_dino(dino_sd, "tiny")
    eps = dict(pos_sim_thresh=-1, points_per_batch=32, filter_thresh=0.3, min_mask_region_area=100)
    injected_pipeline_case("p64", sam, dino, dict(grid_size=8, max_prompts=64, **eps), seed=1)
    injected_pipeline_case("p1024", sam, dino, dict(grid_size=32, max_prompts=1024, **eps), seed=2)
    injected_pipeline_case("p4096", sam, dino, dict(grid_size=64, max_prompts=4096, pos_sim_thresh=-1, points_per_batch=64,
                                                    filter_thresh=2.0, min_mask_region_area=0), seed=3,
                           coordinate_trick=True)
    injected_crops_case(sam, dino, dict(grid_size=16, max_prompts=256, crop_n_layers=1, **eps), image_index=4,
                        hw=(600, 900), seed=4)
    for sel in ("max_area", "min_area"):
        injected_pipeline_case(sel, sam, dino, dict(grid_size=16, max_prompts=256, mask_selection=sel, **eps), seed=5)


if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    torch.manual_seed(0)
    if "--amg" in sys.argv:
        amg_case()
        sys.exit(0)
    if "--extra" in sys.argv:
        extra_stage_case()
        amg_case()
        sys.exit(0)
    if "--injected" in sys.argv:
        injected_cases()
        sys.exit(0)
    if "--config3" in sys.argv:
        config3_case()
        sys.exit(0)
    if "--config1" in sys.argv:
        config1_case()
        sys.exit(0)
    if "--config0" in sys.argv:          # only the (slow, ~3 min) full-depth case; the other files stay as they are
        config0_case()
        sys.exit(0)
    stage_case()
    sam, dino = model_case("tiny", "tiny", "tiny")
    pipeline_case("tiny_grid8", sam, dino,
                  dict(grid_size=8, pos_sim_thresh=-1, max_prompts=64, points_per_batch=16,
                       filter_thresh=2.0, min_mask_region_area=0))
    pipeline_case("tiny_eps", sam, dino,
                  dict(grid_size=16, pos_sim_thresh=0.5, max_prompts=48, points_per_batch=8,
                       filter_thresh=0.3, min_mask_region_area=100, pred_iou_thresh=0.05,
                       stability_score_thresh=0.5), image_index=2, hw=(768, 1024))
    model_case("tiny_l", "tiny_l", "tiny")
    injected_cases()
    extra_stage_case()
    amg_case()
    config0_case()
    config1_case()
    config3_case()
