"""CUDA-graph replay of set_image / decode (crowdsam_b200/graphs.py) gives bit-identical results to the eager
launches, keeps earlier results valid, and two predictors sharing one model do not see each other's image."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from crowdsam_b200 import synthetic as weights  # noqa: E402

DEV = "cuda"


@pytest.fixture(autouse=True)
def _restore_graph_switch():
    """Graph replay is opt-in (CSAM_GRAPHS=1); these tests switch it on and off explicitly."""
    from crowdsam_b200 import graphs

    saved = graphs.ENABLED
    yield
    graphs.ENABLED = saved


def _fresh_predictor():
    from crowdsam_b200.predictor import SamPredictor
    from test_gpu_model import make_predictor

    base, *_ = make_predictor("tiny")
    base.model.mask_decoder.engine().graphs.clear()
    return SamPredictor(base.model, base.dino_model)


def _prompts(pred, n=8):
    pts = np.array([[37 + 113 * i, 900 - 97 * i] for i in range(n)])
    coords = torch.as_tensor(pred.transform.apply_coords(pts, pred.original_size))[:, None, :]
    return coords, torch.ones(n, dtype=torch.int)[:, None]


def test_graph_replay_equals_eager():
    from crowdsam_b200 import graphs

    pred = _fresh_predictor()
    imgs = [weights.synthetic_image(20 + i) for i in range(4)]
    # eager references
    graphs.ENABLED = False
    try:
        ref = []
        for im in imgs:
            pred.set_image(im)
            c, l = _prompts(pred)
            low, iou, cls = pred.decode_low_res(c, l)
            ref.append((pred.features.clone(), pred.dino_feats.clone(), low.clone(), iou.clone(), cls.clone()))
    finally:
        graphs.ENABLED = True
    pred.model.mask_decoder.engine().graphs.clear()
    c0, r0 = graphs.captures, graphs.replayed_launches
    kept = []
    for i, im in enumerate(imgs):
        pred.set_image(im)                       # 1st eager, 2nd captures + replays, 3rd / 4th replay
        c, l = _prompts(pred)
        masks, iou, cls, low = pred.predict_torch(c, l, return_logits=False)
        kept.append((pred.features, pred.dino_feats, low, iou, cls))
        assert masks.dtype == torch.bool
    assert graphs.captures - c0 >= 2 and graphs.replayed_launches - r0 > 150     # set_image + decode graphs ran
    for got, want in zip(kept, ref):             # every image's results are still intact after later replays
        for a, b in zip(got, want):
            assert torch.equal(a, b)


def test_pipeline_with_graphs_equals_eager():
    from crowdsam_b200 import graphs
    from crowdsam_b200.pipeline import CrowdSAM

    pred = _fresh_predictor()
    cfg = dict(weights.DEFAULT_TEST_CFG)
    cfg.update(grid_size=8, pos_sim_thresh=-1, max_prompts=64, points_per_batch=16, filter_thresh=2.0,
               min_mask_region_area=0, apply_box_offsets=False, fuse_simmap=False, output_rles=True)
    model = CrowdSAM({"environ": {"device": DEV}, "model": {"trainfree": False}, "test": cfg}, None, predictor=pred)
    outs = []
    for enabled in (False, True):
        graphs.ENABLED = enabled
        try:
            res = []
            for i in range(3):
                np.random.seed(42)
                r = dict(model.generate(weights.synthetic_image(30 + i)).items())
                res.append((np.asarray(r["boxes"]).copy(), np.asarray(r["scores"]).copy(), [x["counts"] for x in r["rles"]]))
            outs.append(res)
        finally:
            graphs.ENABLED = True
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        assert a[2] == b[2]


def test_two_predictors_sharing_a_model():
    """The decoder engine holds one image state; a second predictor on the same model must not leak its image into
    the first one's predictions (the reference keeps features per predictor, predictor.py:62-69)."""
    from crowdsam_b200.predictor import SamPredictor

    a = _fresh_predictor()
    b = SamPredictor(a.model, a.dino_model)
    from crowdsam_b200 import graphs

    graphs.ENABLED = True
    im_a, im_b = weights.synthetic_image(40), weights.synthetic_image(41)
    a.set_image(im_a)
    c, l = _prompts(a, 4)
    want = [t.clone() for t in a.decode_low_res(c, l)]
    for _ in range(3):
        b.set_image(im_b)                         # also walks b through eager -> capture -> replay
        got = a.decode_low_res(c, l)
        for x, y in zip(got, want):
            assert (x - y).abs().max() <= 1e-4 * y.abs().max()
        a.set_image(im_a)


def test_two_stream_encoders_equal_one_stream(monkeypatch):
    """SAM encoder and DINOv2 on two streams (engine.interleave_two_streams, the default) give bit-identical features,
    DINOv2 tokens and decoder outputs to the single-stream launch order, image after image (the side stream's buffers
    are recycled between images)."""
    from crowdsam_b200 import engine, graphs

    graphs.ENABLED = False
    pred = _fresh_predictor()
    imgs = [weights.synthetic_image(40 + i) for i in range(3)]
    runs = {}
    for mode in ("0", "1", "0"):
        monkeypatch.setenv("CSAM_TWO_STREAMS", mode)
        assert engine.two_streams_enabled() == (mode == "1")
        out = []
        for im in imgs:
            pred.set_image(im)
            c, l = _prompts(pred)
            low, iou, cls = pred.decode_low_res(c, l)
            out.append((pred.features.clone(), pred.dino_feats.clone(), low.clone(), iou.clone(), cls.clone()))
        runs.setdefault(mode, []).append(out)
    torch.cuda.synchronize()
    for other in (runs["1"][0], runs["0"][1]):
        for got, want in zip(other, runs["0"][0]):
            for a, b in zip(got, want):
                assert torch.equal(a, b)
