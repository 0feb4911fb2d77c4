"""GPU parity of the full hot path against the CPU oracle (oracle/restate.py, pinned to the reference by
tests/golden) and against the committed golden vectors of the real reference.

Tolerance (north_star): 1e-3.  Stated here as max|a-b| <= 1e-3 * max|b| per tensor in the default
'x3' precision mode (fp16 hi/lo split, fp32 accumulate); measured errors are ~1e-5.  Boxes / NMS indices
are bit-exact; bool masks may differ only on pixels whose logit is within 1e-4 of the threshold.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate, weights  # noqa: E402

DEV = "cuda"
TOL = 1e-3


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


_cache = {}


def make_predictor(arch, dino_arch="tiny"):
    key = (arch, dino_arch, os.environ.get("CSAM_PRECISION", "x3"))
    if key in _cache:
        return _cache[key]
    from crowdsam_b200.build import _build_sam
    from crowdsam_b200.modules import DinoVisionTransformer
    from crowdsam_b200.predictor import SamPredictor

    D, depth, heads, glob = weights.SAM_ARCHS[arch]
    sam_sd, dino_sd = weights.make_sam_state(arch), weights.make_dino_state(dino_arch)
    sam = _build_sam(D, depth, heads, 1, glob)
    sam.load_state_dict(sam_sd, strict=True)
    dD, ddepth, dheads = weights.DINO_ARCHS[dino_arch]
    dino = DinoVisionTransformer(dD, ddepth, dheads)
    dino.load_state_dict(dino_sd, strict=True)
    pred = SamPredictor(sam.to(DEV), dino.to(DEV))
    _cache[key] = (pred, sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads))
    return _cache[key]


@pytest.mark.parametrize("arch", ["tiny", "tiny_l"])
def test_set_image_vs_oracle_and_golden(arch, golden_dir):
    pred, sam_sd, dino_sd, scfg, dcfg = make_predictor(arch)
    img = weights.synthetic_image(0)
    pred.set_image(img)
    t = torch.as_tensor(img).permute(2, 0, 1)[None]
    feats, dino = restate.set_image(sam_sd, dino_sd, t, scfg, dcfg)
    assert pred.features.shape == (1, 256, 64, 64) and pred.dino_feats.shape == (1, 73, 73, 1024)
    assert _rel(pred.features, feats) < TOL, _rel(pred.features, feats)
    assert _rel(pred.dino_feats, dino) < TOL, _rel(pred.dino_feats, dino)
    _check_model_golden(pred, np.load(os.path.join(golden_dir, f"model_{arch}.npz")))


def _check_model_golden(pred, g):
    """pred has the square golden image set: compare with the real reference's outputs stored by make_golden.py."""
    assert _rel(pred.features[:, ::8, ::2, ::2], g["features"]) < TOL
    assert _rel(pred.dino_feats[:, ::6, ::6, ::8], g["dino_feats"]) < TOL
    assert _rel(pred.predict_fg_map()[:, :, ::4, ::4], g["fg_map"]) < TOL
    assert _rel(pred.model.prompt_encoder.get_dense_pe()[:, ::4, ::4, ::4], g["dense_pe"]) < 1e-5
    # decoder on the golden prompts
    pts = g["points"]
    coords = torch.as_tensor(pred.transform.apply_coords(pts, pred.original_size))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    masks, iou, cls, low = pred.predict_torch(coords, labels, multimask_output=True, return_logits=True)
    assert _rel(low[:, :, ::8, ::8], g["low_res"]) < TOL, _rel(low[:, :, ::8, ::8], g["low_res"])
    assert _rel(iou, g["iou_pred"]) < TOL and _rel(cls, g["cls"]) < TOL
    assert _rel(masks[:, :, ::32, ::32], g["masks"]) < TOL
    bm = pred.predict_torch(coords, labels, multimask_output=True, return_logits=False)[0]
    assert bm.dtype == torch.bool and ((bm != (masks > 0)) & (masks.abs() > 1e-4)).sum() == 0
    one = pred.predict_torch(coords, labels, multimask_output=False, return_logits=True)
    assert one[0].shape[1] == 1 and one[1].shape == (len(pts), 1)


def test_decoder_teacher_forced():
    """Decoder alone, fed the ORACLE's embeddings through the assignable predictor state (the path
    tools/train.py uses), so decoder error is measured without encoder error."""
    pred, sam_sd, dino_sd, scfg, dcfg = make_predictor("tiny")
    img = weights.synthetic_image(3)
    t = torch.as_tensor(img).permute(2, 0, 1)[None]
    feats, dino = restate.set_image(sam_sd, dino_sd, t, scfg, dcfg)
    pred.reset_image()
    pred.features, pred.dino_feats = feats.to(DEV), dino.to(DEV)
    pred.original_size, pred.input_size, pred.is_image_set = (1024, 1024), (1024, 1024), True
    pts = np.array([[100, 200], [700, 40], [512, 900], [5, 1000], [1023, 0], [333, 333], [64, 64]])
    coords = torch.as_tensor(pred.transform.apply_coords(pts, (1024, 1024)))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    low, iou, cls = pred.decode_low_res(coords, labels)
    sparse = restate.embed_points(sam_sd, coords, labels)
    rl, ri, rc = restate.mask_decoder(sam_sd, feats, restate.dense_pe(sam_sd), sparse, dino)
    assert _rel(low, rl) < TOL, _rel(low, rl)
    assert _rel(iou, ri) < TOL and _rel(cls, rc) < TOL, (_rel(iou, ri), _rel(cls, rc))
    assert _rel(pred.predict_fg_map(), restate.fg_map(sam_sd, dino)) < TOL


def test_non_square_image_vs_golden(golden_dir):
    pred, *_ = make_predictor("tiny")
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    img = weights.synthetic_image(1, 600, 900)
    pred.set_image(img)
    assert pred.input_size == (683, 1024) and pred.original_size == (600, 900)
    assert _rel(pred.features[:, ::8, ::2, ::2], g["ns_features"]) < TOL
    pts = g["ns_points"]
    coords = torch.as_tensor(pred.transform.apply_coords(pts, pred.original_size))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    masks, iou, cls, low = pred.predict_torch(coords, labels, multimask_output=True, return_logits=True)
    assert masks.shape == (3, 4, 600, 900)
    assert _rel(low[:, :, ::8, ::8], g["ns_low_res"]) < TOL
    assert _rel(masks[:, :, ::24, ::36], g["ns_masks"]) < TOL
    assert _rel(iou, g["ns_iou_pred"]) < TOL


def _decode_coco(s, size):
    return restate.coco_rle_decode(s, size)


@pytest.mark.parametrize("name", ["tiny_grid8", "tiny_eps"])
def test_crowdsam_generate_vs_golden_and_oracle(name, golden_dir):
    from crowdsam_b200.pipeline import CrowdSAM

    pred, sam_sd, dino_sd, scfg, dcfg = make_predictor("tiny")
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    over = {}
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        try:
            over[str(k)] = int(str(v))
        except ValueError:
            over[str(k)] = float(str(v))
    test_cfg = dict(restate.DEFAULT_TEST_CFG)
    test_cfg.update(over)
    test_cfg.update(apply_box_offsets=False, fuse_simmap=False, output_rles=True)
    _check_pipeline_golden(pred, g, test_cfg)


def _check_pipeline_golden(pred, g, test_cfg):
    from crowdsam_b200.pipeline import CrowdSAM

    cfg = {"environ": {"device": DEV}, "model": {"trainfree": False}, "test": test_cfg}
    model = CrowdSAM(cfg, None, predictor=pred)
    hw = tuple(int(x) for x in g["hw"])
    img = weights.synthetic_image(int(g["image_index"]), *hw)
    np.random.seed(42)
    res = model.generate(img)
    if "points" not in g.files:
        # the reference found nothing (model.py:182-183: boxes / scores of shape [0,4], no other keys): same here
        assert set(dict(res.items()).keys()) == {"boxes", "scores", "rles"} and len(res["rles"]) == 0
        assert np.asarray(res["boxes"]).shape == g["boxes"].shape and np.asarray(res["scores"]).shape == g["scores"].shape
        return
    assert set(["points", "categories", "stability_score", "boxes", "scores", "rles", "rles_info", "crop_boxes", "fboxes"]) <= set(dict(res.items()).keys())
    np.testing.assert_array_equal(res["boxes"], g["boxes"])
    np.testing.assert_array_equal(res["points"], g["points"])
    np.testing.assert_array_equal(res["categories"], g["categories"])
    np.testing.assert_allclose(res["scores"], g["scores"], rtol=TOL, atol=1e-5)
    np.testing.assert_allclose(res["stability_score"], g["stability_score"], rtol=TOL, atol=1e-5)
    for r, ref in zip(res["rles"], g["rle_counts"]):
        a, b = _decode_coco(r["counts"], r["size"]), _decode_coco(str(ref), r["size"])
        assert (a != b).mean() < 1e-4


def test_config0_vit_b_full_depth_vs_reference(golden_dir):
    """BASELINE.json configs[0]: SAM ViT-B at full depth (12 blocks) + DINOv2 ViT-L/14 (24 blocks), 1024x1024 synthetic
    image, 8x8 prompt grid, against the outputs of the REAL reference run on CPU (tests/golden/make_golden.py
    --config0): encoder features, DINOv2 tokens, fg map, decoder logits / IoU / class scores, and CrowdSAM.generate."""
    pred, *_ = make_predictor("vit_b", "dinov2_vitl14")
    pred.set_image(weights.synthetic_image(0))
    _check_model_golden(pred, np.load(os.path.join(golden_dir, "model_vit_b.npz")))
    g = np.load(os.path.join(golden_dir, "pipeline_vit_b_grid8.npz"))
    test_cfg = dict(restate.DEFAULT_TEST_CFG)
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        test_cfg[str(k)] = int(str(v)) if str(v).lstrip("-").isdigit() else float(str(v))
    test_cfg.update(apply_box_offsets=False, fuse_simmap=False, output_rles=True)
    _check_pipeline_golden(pred, g, test_cfg)


def test_config1_vit_l_headline_vs_reference(golden_dir):
    """BASELINE.json configs[1], the bench workload: SAM ViT-L (24 blocks) + DINOv2 ViT-L/14, 1024x1024 synthetic image,
    32x32 grid = 1024 prompts, against the outputs of the REAL reference run on CPU (make_golden.py --config1)."""
    pred, *_ = make_predictor("vit_l", "dinov2_vitl14")
    pred.set_image(weights.synthetic_image(0))
    _check_model_golden(pred, np.load(os.path.join(golden_dir, "model_vit_l.npz")))
    g = np.load(os.path.join(golden_dir, "pipeline_vit_l_grid32.npz"))
    test_cfg = dict(restate.DEFAULT_TEST_CFG)
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        test_cfg[str(k)] = int(str(v)) if str(v).lstrip("-").isdigit() else float(str(v))
    test_cfg.update(apply_box_offsets=False, fuse_simmap=False, output_rles=True)
    _check_pipeline_golden(pred, g, test_cfg)
    # the same image with all 1024 prompts in ONE decoder batch (what bench.py runs) gives the same detections
    test_cfg["points_per_batch"] = 1024
    _check_pipeline_golden(pred, g, test_cfg)


def test_config2_vit_l_grid64_vs_reference(golden_dir):
    """BASELINE.json configs[2]: SAM ViT-L, 64x64 dense grid = 4096 prompts decoded in ONE batch, against the real
    reference's CrowdSAM.generate (which took them 64 at a time)."""
    pred, *_ = make_predictor("vit_l", "dinov2_vitl14")
    g = np.load(os.path.join(golden_dir, "pipeline_vit_l_grid64.npz"))
    test_cfg = dict(restate.DEFAULT_TEST_CFG)
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        test_cfg[str(k)] = int(str(v)) if str(v).lstrip("-").isdigit() else float(str(v))
    test_cfg.update(apply_box_offsets=False, fuse_simmap=False, output_rles=True, points_per_batch=4096)
    _check_pipeline_golden(pred, g, test_cfg)


def test_config3_vit_h_vs_reference(golden_dir):
    """BASELINE.json configs[3]: SAM ViT-H (32 blocks, head dim 80 on the tcgen05 attention) + DINOv2 ViT-L/14, 32x32 grid,
    against the outputs of the REAL reference run on CPU (make_golden.py --config3)."""
    pred, *_ = make_predictor("vit_h", "dinov2_vitl14")
    pred.set_image(weights.synthetic_image(0))
    _check_model_golden(pred, np.load(os.path.join(golden_dir, "model_vit_h.npz")))
    # image 0: nothing survives the filters in the reference (its empty-result branch); image 3: one detection
    for name in ("pipeline_vit_h_grid32.npz", "pipeline_vit_h_grid32_img3.npz"):
        g = np.load(os.path.join(golden_dir, name))
        test_cfg = dict(restate.DEFAULT_TEST_CFG)
        for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
            test_cfg[str(k)] = int(str(v)) if str(v).lstrip("-").isdigit() else float(str(v))
        test_cfg.update(apply_box_offsets=False, fuse_simmap=False, output_rles=True, points_per_batch=1024)
        _check_pipeline_golden(pred, g, test_cfg)


def test_automatic_mask_generator_vs_oracle():
    """SamAutomaticMaskGenerator (upstream-SAM grid semantics: every grid point, all 4 masks per point, IoU / stability
    filters, box NMS by predicted IoU) against the same steps composed from the CPU oracle."""
    from crowdsam_b200 import amg
    from crowdsam_b200.automask import SamAutomaticMaskGenerator

    pred, sam_sd, dino_sd, scfg, dcfg = make_predictor("tiny")
    img = weights.synthetic_image(5)
    gen = SamAutomaticMaskGenerator(pred.model, pred.dino_model, points_per_side=4, points_per_batch=8,
                                    pred_iou_thresh=0.0, stability_score_thresh=0.0, box_nms_thresh=0.7,
                                    output_mode="binary_mask")
    res = gen.generate(img)
    # oracle: same grid, same filters
    t = torch.as_tensor(img).permute(2, 0, 1)[None]
    feats, dino = restate.set_image(sam_sd, dino_sd, t, scfg, dcfg)
    pts = amg.build_point_grid(4) * np.array([[1024, 1024]])
    coords = torch.as_tensor(restate.apply_coords(pts, (1024, 1024)))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    low, iou, _ = restate.mask_decoder(sam_sd, feats, restate.dense_pe(sam_sd), restate.embed_points(sam_sd, coords, labels), dino)
    full = restate.postprocess_masks(low, (1024, 1024), (1024, 1024)).reshape(-1, 1024, 1024)
    iou_f = iou.reshape(-1)
    stab = restate.stability_score(full, 0.0, 1.0)
    ok = torch.ones_like(iou_f, dtype=torch.bool)          # thresholds of 0 disable both filters (reference :294,302)
    boxes = restate.mask_to_box(full[ok] > 0.0)
    keep = restate.nms_reference(boxes.float().numpy(), iou_f[ok].numpy(), 0.7)
    assert len(res) == len(keep) > 0
    ref_iou = iou_f[ok][keep]
    ref_boxes = boxes[keep]
    for i, r in enumerate(res):
        assert set(r.keys()) == {"segmentation", "area", "bbox", "predicted_iou", "point_coords", "stability_score", "crop_box"}
        assert abs(r["predicted_iou"] - float(ref_iou[i])) <= TOL * max(1.0, float(ref_iou.abs().max()))
        b = ref_boxes[i].tolist()
        assert r["bbox"] == [b[0], b[1], b[2] - b[0], b[3] - b[1]]
        assert r["segmentation"].shape == (1024, 1024) and r["segmentation"].dtype == bool
    # small-region cleanup path runs on the device and keeps the output well formed
    gen2 = SamAutomaticMaskGenerator(pred.model, pred.dino_model, points_per_side=4, points_per_batch=16,
                                     pred_iou_thresh=0.0, stability_score_thresh=0.0, min_mask_region_area=50,
                                     output_mode="coco_rle")
    res2 = gen2.generate(img)
    assert len(res2) > 0 and all(isinstance(r["segmentation"]["counts"], str) for r in res2)
    # dino_model is optional: masks / IoU do not depend on the PWD-Net features
    gen3 = SamAutomaticMaskGenerator(pred.model, None, points_per_side=4, points_per_batch=8, pred_iou_thresh=0.0,
                                     stability_score_thresh=0.0, box_nms_thresh=0.7, output_mode="binary_mask")
    res3 = gen3.generate(img)
    assert [r["bbox"] for r in res3] == [r["bbox"] for r in res] and [r["predicted_iou"] for r in res3] == [r["predicted_iou"] for r in res]
    with pytest.raises(RuntimeError):
        gen3.predictor.set_image(img)
        gen3.predictor.predict_fg_map()


def test_errors_and_state():
    pred, *_ = make_predictor("tiny")
    pred.reset_image()
    with pytest.raises(RuntimeError):
        pred.predict_torch(torch.zeros(1, 1, 2), torch.ones(1, 1))
    with pytest.raises(RuntimeError):
        pred.get_image_embedding()
    with pytest.raises(AssertionError):
        pred.set_image(np.zeros((10, 10, 3), np.uint8), image_format="XYZ")
    with pytest.raises(AssertionError):
        pred.set_torch_image(torch.zeros(1, 3, 100, 100, dtype=torch.uint8), (100, 100))
    # predict(): the reference's four return values (predictor.py:196-212), box / mask prompts refused up front
    pred.set_image(weights.synthetic_image(0))
    m, iou, cls, low = pred.predict(np.array([[500.0, 400.0]]), np.array([1]))
    assert m.shape == (4, 1024, 1024) and m.dtype == bool and iou.shape == (4,) and cls.shape == (4, 1)
    assert isinstance(low, torch.Tensor) and low.shape == (4, 256, 256)
    with pytest.raises(NotImplementedError):
        pred.predict(np.array([[5.0, 4.0]]), np.array([1]), box=np.array([0, 0, 10, 10]))


def test_full_scale_vit_l_grid32_against_oracle_run():
    """BASELINE.json configs[1] at full size (ViT-L + DINOv2-L, 1024 prompts, recipe-v2 weights, image 3).
    The expected values come from one run of the CPU oracle pipeline (oracle/restate.py OracleCrowdSAM,
    95 s on 8 cores, same config as bench.test_cfg): 881 masks survive the IoU / stability filters, box NMS
    keeps one full-image box with PWD score 1.5351906."""
    import bench
    from crowdsam_b200.pipeline import CrowdSAM

    pred, *_ = make_predictor("vit_l", "dinov2_vitl14")
    cfg = {"environ": {"device": DEV}, "model": {"trainfree": False}, "test": bench.test_cfg(256)}
    model = CrowdSAM(cfg, None, predictor=pred)
    np.random.seed(42)
    res = model.generate(weights.synthetic_image(3))
    n_into_nms, kept = model.last_counts
    # a prompt sitting exactly on a filter threshold may flip: allow 1% of the count
    assert abs(n_into_nms - 881) <= 9, n_into_nms
    assert kept == 1
    np.testing.assert_array_equal(res["boxes"], np.array([[0.0, 0.0, 1023.0, 1023.0]], dtype=np.float32))
    np.testing.assert_allclose(res["scores"], [1.5351906], rtol=TOL)
