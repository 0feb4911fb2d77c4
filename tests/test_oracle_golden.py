"""Pin oracle/restate.py to the REAL reference via the committed golden vectors
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference).

Tolerances: both sides are fp32 torch CPU; the restatement reorders nothing material, so
outputs agree to ~1e-5 relative.  Integer outputs (boxes, NMS indices, RLE) are bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, restate, weights

torch.set_grad_enabled(False)


def _close(a, b, rtol=2e-4, atol=None, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-6)
    err = np.abs(a - b).max()
    tol = rtol * scale if atol is None else atol
    assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e} (scale {scale:.3e})"


@pytest.fixture(scope="module", params=["tiny", "tiny_l"])
def model_case(request, golden_dir):
    name = request.param
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    sam_sd = weights.make_sam_state(name)
    dino_sd = weights.make_dino_state("tiny")
    D, depth, heads, glob = weights.SAM_ARCHS[name]
    _, ddepth, dheads = weights.DINO_ARCHS["tiny"]
    return g, sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads)


def test_encoder_dino_decoder_square(model_case):
    g, sam_sd, dino_sd, scfg, dcfg = model_case
    img = torch.as_tensor(weights.synthetic_image(0)).permute(2, 0, 1)[None]
    feats, dino = restate.set_image(sam_sd, dino_sd, img, scfg, dcfg)
    _close(feats[:, ::8, ::2, ::2], g["features"], what="features")
    _close(dino[:, ::6, ::6, ::8], g["dino_feats"], what="dino_feats")
    _close(restate.fg_map(sam_sd, dino)[:, :, ::4, ::4], g["fg_map"], what="fg_map")
    _close(restate.dense_pe(sam_sd)[:, ::4, ::4, ::4], g["dense_pe"], what="dense_pe")
    pts = g["points"]
    coords = torch.as_tensor(restate.apply_coords(pts, (1024, 1024)))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    sparse = restate.embed_points(sam_sd, coords, labels)
    low, iou, cls = restate.mask_decoder(sam_sd, feats, restate.dense_pe(sam_sd), sparse, dino)
    _close(low[:, :, ::8, ::8], g["low_res"], what="low_res")
    _close(iou, g["iou_pred"], what="iou")
    _close(cls, g["cls"], what="cls")
    full = restate.postprocess_masks(low, (1024, 1024), (1024, 1024))
    _close(full[:, :, ::32, ::32], g["masks"], what="masks")


def test_config0_vit_b_full_depth(golden_dir):
    """BASELINE.json configs[0] at full depth (SAM ViT-B 12 blocks + DINOv2 ViT-L/14 24 blocks): the oracle against
    the real reference's outputs (about a minute of CPU)."""
    g = np.load(os.path.join(golden_dir, "model_vit_b.npz"))
    sam_sd, dino_sd = weights.make_sam_state("vit_b"), weights.make_dino_state("dinov2_vitl14")
    _, depth, heads, glob = weights.SAM_ARCHS["vit_b"]
    _, ddepth, dheads = weights.DINO_ARCHS["dinov2_vitl14"]
    test_encoder_dino_decoder_square((g, sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads)))


def test_config1_vit_l_full_depth(golden_dir):
    """BASELINE.json configs[1] at full depth (SAM ViT-L + DINOv2 ViT-L/14): the oracle against the real reference."""
    g = np.load(os.path.join(golden_dir, "model_vit_l.npz"))
    sam_sd, dino_sd = weights.make_sam_state("vit_l"), weights.make_dino_state("dinov2_vitl14")
    _, depth, heads, glob = weights.SAM_ARCHS["vit_l"]
    _, ddepth, dheads = weights.DINO_ARCHS["dinov2_vitl14"]
    test_encoder_dino_decoder_square((g, sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads)))


@pytest.mark.skipif(os.environ.get("CSAM_SLOW_TESTS", "0") != "1", reason="~1.5 min of CPU: set CSAM_SLOW_TESTS=1")
def test_config3_vit_h_full_depth(golden_dir):
    """BASELINE.json configs[3] at full depth (SAM ViT-H + DINOv2 ViT-L/14): the oracle against the real reference."""
    g = np.load(os.path.join(golden_dir, "model_vit_h.npz"))
    sam_sd, dino_sd = weights.make_sam_state("vit_h"), weights.make_dino_state("dinov2_vitl14")
    _, depth, heads, glob = weights.SAM_ARCHS["vit_h"]
    _, ddepth, dheads = weights.DINO_ARCHS["dinov2_vitl14"]
    test_encoder_dino_decoder_square((g, sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads)))


def test_non_square_image(model_case):
    g, sam_sd, dino_sd, scfg, dcfg = model_case
    from PIL import Image

    img = weights.synthetic_image(1, 600, 900)
    ih, iw = restate.preprocess_shape(600, 900)
    assert (ih, iw) == (683, 1024)
    inp = np.array(Image.fromarray(img).resize((iw, ih), Image.BILINEAR))
    t = torch.as_tensor(inp).permute(2, 0, 1)[None]
    feats, dino = restate.set_image(sam_sd, dino_sd, t, scfg, dcfg)
    _close(feats[:, ::8, ::2, ::2], g["ns_features"], what="ns_features")
    pts = g["ns_points"]
    coords = torch.as_tensor(restate.apply_coords(pts, (600, 900)))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    sparse = restate.embed_points(sam_sd, coords, labels)
    low, iou, cls = restate.mask_decoder(sam_sd, feats, restate.dense_pe(sam_sd), sparse, dino)
    _close(low[:, :, ::8, ::8], g["ns_low_res"], what="ns_low_res")
    _close(iou, g["ns_iou_pred"], what="ns_iou")
    full = restate.postprocess_masks(low, (ih, iw), (600, 900))
    _close(full[:, :, ::24, ::36], g["ns_masks"], what="ns_masks")


@pytest.mark.parametrize("name", ["tiny_grid8", "tiny_eps"])
def test_pipeline_generate(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    over = {}
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        v = str(v)
        try:
            over[str(k)] = int(v)
        except ValueError:
            over[str(k)] = float(v)
    sam_sd = weights.make_sam_state("tiny")
    dino_sd = weights.make_dino_state("tiny")
    _, depth, heads, glob = weights.SAM_ARCHS["tiny"]
    m = restate.OracleCrowdSAM(sam_sd, dino_sd, (depth, heads, glob), weights.DINO_ARCHS["tiny"][1:], over)
    hw = tuple(int(x) for x in g["hw"])
    img = weights.synthetic_image(int(g["image_index"]), *hw)
    np.random.seed(42)
    res = m.generate(img)
    assert len(res["boxes"]) == len(g["boxes"])
    np.testing.assert_array_equal(res["boxes"], g["boxes"])
    np.testing.assert_array_equal(res["points"], g["points"])
    np.testing.assert_array_equal(res["categories"], g["categories"])
    np.testing.assert_array_equal(res["crop_boxes"], g["crop_boxes"])
    _close(res["scores"], g["scores"], rtol=1e-4, what="scores")
    _close(res["stability_score"], g["stability_score"], rtol=1e-4, what="stability")
    # near-zero logits may flip a handful of pixels between two fp32 evaluation orders:
    # compare decoded masks instead of strings, allow <= 1e-4 of the pixels
    from oracle.restate import rle_to_mask  # noqa: F401

    assert [r["size"] for r in res["rles"]] == g["rle_sizes"].tolist()
    for r, ref_counts in zip(res["rles"], g["rle_counts"]):
        if r["counts"] == str(ref_counts):
            continue
        a = _decode_coco(r["counts"], r["size"])
        b = _decode_coco(str(ref_counts), r["size"])
        assert (a != b).mean() < 1e-4


def _decode_coco(s, size):
    return restate.coco_rle_decode(s, size)


def test_stage_post_nms_rle(golden_dir):
    g = np.load(os.path.join(golden_dir, "stage_post_nms.npz"))
    P = 48
    low, iou, cls = fixtures.blob_logits(P, seed=0)
    for tag, (inp, orig) in {"sq": ((1024, 1024), (1024, 1024)), "ns": ((683, 1024), (600, 900))}.items():
        full = restate.postprocess_masks(low, inp, orig)
        score = torch.clamp(iou, 0.0) * cls.squeeze(2).sigmoid()
        sel = score.max(dim=-1)[1]
        np.testing.assert_array_equal(sel.numpy(), g[f"{tag}_sel"])
        m = full[torch.arange(P), sel]
        stab = restate.stability_score(m, 0.0, 1.0)
        np.testing.assert_array_equal(stab.numpy(), g[f"{tag}_stability"])
        boxes = restate.mask_to_box(m > 0.0)
        np.testing.assert_array_equal(boxes.numpy(), g[f"{tag}_boxes"])
        keep = restate.nms_reference(boxes.float().numpy(), score[torch.arange(P), sel].numpy(), 0.65)
        np.testing.assert_array_equal(keep, g[f"{tag}_nms_keep"])
        if tag == "sq":
            r0 = restate.mask_to_rle((m[0] > 0).numpy())
            np.testing.assert_array_equal(np.array(r0["counts"]), g["sq_rle_counts0"])
            r3 = restate.mask_to_rle((m[3] > 0).numpy())
            np.testing.assert_array_equal(np.array(r3["counts"]), g["sq_rle_counts3"])
            assert restate.coco_rle_string(r0["counts"]) == str(g["sq_rle_str0"])
            np.testing.assert_array_equal(restate.rle_to_mask(r0), (m[0] > 0).numpy())
            np.testing.assert_array_equal(_decode_coco(str(g["sq_rle_str0"]), [1024, 1024]), (m[0] > 0).numpy())


@pytest.mark.parametrize("n,seed,binary", [(257, 0, False), (3000, 1, False), (3000, 2, True), (1, 3, False)])
def test_nms_matches_torchvision_golden(n, seed, binary, golden_dir):
    g = np.load(os.path.join(golden_dir, "stage_post_nms.npz"))
    b, s = fixtures.random_boxes(n, seed, binary_scores=binary)
    for thr in (0.65, 0.7):
        keep = restate.nms_reference(b, s, thr)
        np.testing.assert_array_equal(keep, g[f"nms_{n}_{seed}_{thr}"])


def test_nms_empty():
    assert restate.nms_reference(np.zeros((0, 4)), np.zeros((0,)), 0.5).shape == (0,)


# ---- multi-detection end-to-end cases: the real reference pipeline on injected decoder outputs -------------------
@pytest.mark.parametrize("name", ["p64", "p1024", "max_area", "min_area"])
def test_injected_pipeline(name, golden_dir):
    """CrowdSAM.generate of the real reference (EPS pruning with many occupying masks, select_mask, filters, boxes,
    NMS with real suppression, small-region cleanup + its tied-score NMS, RLE order) vs the oracle, on decoder
    outputs injected per prompt point: 46 / 196 / 76 / 96 detections, everything bit-exact."""
    import inject_util as iu

    g = np.load(os.path.join(golden_dir, f"pipeline_inj_{name}.npz"))
    log = []
    m = iu.oracle_model(g, log)
    np.random.seed(42)
    res = m.generate(iu.golden_image(g))
    assert [len(x) for x in log] == g["call_sizes"].tolist()            # same EPS batches ...
    np.testing.assert_array_equal(np.concatenate(log, 0), g["call_points"])   # ... of the same prompts
    assert len(res["boxes"]) >= 20
    iu.compare_result(res, g)


def test_injected_crops(golden_dir):
    """crop_n_layers = 1: every crop through the real `_process_crop` (resize to max_size, crop-edge filter active,
    uncrop) and the reference's cross-crop NMS statement (model.py:167-176), vs the oracle."""
    import inject_util as iu

    g = np.load(os.path.join(golden_dir, "pipeline_inj_crops.npz"))
    m = iu.oracle_model(g)
    img = iu.golden_image(g)
    boxes = restate.crop_boxes_for(img.shape[:2], m.cfg["crop_n_layers"], m.cfg["crop_overlap_ratio"])
    np.testing.assert_array_equal(np.array(boxes), g["crop_boxes_all"])
    np.random.seed(42)
    allb, allc = [], []
    for ci, cb in enumerate(boxes):
        d = m.process_crop(img, cb)
        assert (0 if d is None else len(d["boxes"])) == int(g[f"crop{ci}_n"])
        d = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
        d["rles"] = [restate.coco_encode_rle(r) for r in d["rles"]]
        iu.compare_result(d, g, prefix=f"crop{ci}_")
        np.testing.assert_array_equal(d["crop_boxes"], g[f"crop{ci}_crop_boxes"])
        allb.append(d["boxes"]); allc.append(d["crop_boxes"])
    allb, allc = np.concatenate(allb), np.concatenate(allc).astype(np.float32)
    sc = 1.0 / ((allc[:, 2] - allc[:, 0]) * (allc[:, 3] - allc[:, 1]))
    keep = restate.nms_reference(allb.astype(np.float32), sc.astype(np.float32), m.cfg["crop_nms_thresh"])
    np.testing.assert_array_equal(keep, g["cross_keep"])
    assert len(keep) < len(allb)                                         # the cross-crop NMS really suppresses


def test_stage_extra(golden_dir):
    """mask_iou_nms of the reference (crowdsam/utils.py:422-459), torchvision nms with NaN / signed-zero scores, and the
    post-processing stage functions on injected logits at P = 64 (all four planes), vs the oracle."""
    g = np.load(os.path.join(golden_dir, "stage_extra.npz"))
    low, iou, _ = fixtures.injected_decoder_outputs(g["miou_points"], seed=7)
    masks, scores = low[:, 2] > 0, iou[:, 2].numpy()
    for thr in (0.3, 0.5, 0.8):
        np.testing.assert_array_equal(restate.mask_iou_nms(scores, masks, thr), g[f"miou_keep_{thr}"])
    assert len(g["miou_keep_0.3"]) < len(g["miou_keep_0.8"]) < 48
    b, _ = fixtures.random_boxes(300, 9)
    np.testing.assert_array_equal(restate.nms_reference(b, g["nan_scores"], 0.65), g["nan_keep"])
    low, _, _ = fixtures.injected_decoder_outputs(fixtures.grid_points(8).astype(np.float64), seed=11)
    for tag, (inp, orig) in {"sq": ((1024, 1024), (1024, 1024)), "ns": ((683, 1024), (600, 900))}.items():
        full = restate.postprocess_masks(low, inp, orig).flatten(0, 1)
        np.testing.assert_array_equal(restate.stability_score(full, 0.0, 1.0).numpy(), g[f"p64_{tag}_stability"])
        np.testing.assert_array_equal(restate.mask_to_box(full > 0).numpy(), g[f"p64_{tag}_boxes"])
