"""Shared helpers of the injected-decoder pipeline tests (tests/golden/pipeline_inj_*.npz).

The goldens hold the outputs of the REAL reference's CrowdSAM pipeline run with its predictor's decoder replaced by
`oracle.fixtures.injected_decoder_outputs` (make_golden.py `_inject`).  The CPU tests give the oracle the same
replacement, the GPU tests give it to this repo's predictor at the same boundary (`SamPredictor.decode_low_res`).
"""
import numpy as np
import torch

from oracle import fixtures, restate, weights


def cfg_from_golden(g) -> dict:
    cfg = dict(restate.DEFAULT_TEST_CFG)
    for k, v in zip(g["cfg_keys"], g["cfg_vals"]):
        v = str(v)
        try:
            cfg[str(k)] = int(v)
        except ValueError:
            try:
                cfg[str(k)] = float(v)
            except ValueError:
                cfg[str(k)] = v
    return cfg


def golden_image(g) -> np.ndarray:
    return weights.synthetic_image(int(g["image_index"]), *(int(x) for x in g["hw"]))


def oracle_model(g, log=None):
    """OracleCrowdSAM (tiny SAM + tiny DINOv2 for set_image / the fg map) with the decoder injected."""
    sam_sd, dino_sd = weights.make_sam_state("tiny"), weights.make_dino_state("tiny")
    _, depth, heads, glob = weights.SAM_ARCHS["tiny"]
    m = restate.OracleCrowdSAM(sam_sd, dino_sd, (depth, heads, glob), weights.DINO_ARCHS["tiny"][1:], cfg_from_golden(g))
    seed = int(g["inject_seed"])

    def predict(feats, dino, coords, labels, input_size, original_size):
        xy = coords[:, 0, :].numpy()
        if log is not None:
            log.append(xy.copy())
        low, iou, cls = fixtures.injected_decoder_outputs(xy, seed)
        return restate.postprocess_masks(low, input_size, original_size), iou, cls, low

    m.predict = predict
    return m


def inject_predictor(pred, seed: int, log=None):
    """Give this repo's SamPredictor the injected decoder at the low-res-logit boundary (decode_low_res returns
    device tensors exactly as the CUDA decoder would)."""
    dev = pred.device

    def decode_low_res(point_coords, point_labels, boxes=None, mask_input=None):
        xy = torch.as_tensor(point_coords)[:, 0, :].cpu().numpy()
        if log is not None:
            log.append(xy.copy())
        low, iou, cls = fixtures.injected_decoder_outputs(xy, seed)
        return low.to(dev), iou.to(dev), cls.to(dev)

    pred.decode_low_res = decode_low_res
    return pred


def compare_result(res, g, exact_rle=True, prefix="", float_rtol=0.0):
    """res: mapping with boxes / points / categories / scores / stability_score / rles (COCO dicts)."""
    gb = g[prefix + "boxes"]
    assert len(res["boxes"]) == len(gb), (len(res["boxes"]), len(gb))
    np.testing.assert_array_equal(np.asarray(res["boxes"]), gb)
    np.testing.assert_array_equal(np.asarray(res["points"]), g[prefix + "points"])
    np.testing.assert_array_equal(np.asarray(res["categories"]), g[prefix + "categories"])
    # floating point: the PWD score clamp(iou) * sigmoid(cls) goes through the device's expf (1 ulp from ATen's CPU
    # sigmoid); the stability ratio is a quotient of integer counts.  north_star bar is 1e-3; observed <= 2.2e-7.
    np.testing.assert_allclose(np.asarray(res["scores"]), g[prefix + "scores"], rtol=float_rtol, atol=0)
    np.testing.assert_allclose(np.asarray(res["stability_score"]), g[prefix + "stability_score"], rtol=float_rtol, atol=0)
    ref = [str(s) for s in g[prefix + "rle_counts"]]
    got = [r["counts"] for r in res["rles"]]
    if exact_rle:
        assert got == ref
    return sum(a != b for a, b in zip(got, ref))
