"""CrowdSAM.generate on two engine instances with identical weights, twice each, must agree bit for bit
(run under compute-sanitizer memcheck to stretch timing).  usage: debug_determinism_pipeline.py [arch]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from crowdsam_b200 import synthetic as weights
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor

arch = sys.argv[1] if len(sys.argv) > 1 else "vit_l"
dino_arch = "tiny" if arch.startswith("tiny") else "dinov2_vitl14"
dev = torch.device("cuda", 0)
cfg = dict(weights.DEFAULT_TEST_CFG)
cfg.update(grid_size=8, pos_sim_thresh=-1, max_prompts=64, filter_thresh=2.0, apply_box_offsets=False, fuse_simmap=False,
           output_rles=True)


def make():
    D, depth, heads, glob = weights.SAM_ARCHS[arch]
    sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
    dD, dd, dh = weights.DINO_ARCHS[dino_arch]
    dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(dino_arch), strict=True)
    return CrowdSAM({"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": cfg}, None,
                    predictor=SamPredictor(sam.to(dev), dino.to(dev)))


a, b = make(), make()
imgs = [weights.synthetic_image(3), weights.synthetic_image(4, 600, 900), weights.synthetic_image(5)]
runs = []
for m in (a, b, a, b):
    np.random.seed(42)
    out = []
    for im in imgs:
        r = dict(m.generate(im).items())
        out.append((np.asarray(r["boxes"]).tolist(), np.asarray(r["scores"]).tolist(), np.asarray(r["stability_score"]).tolist(),
                    [x["counts"] for x in r["rles"]], m.last_counts))
    runs.append(out)
for i, name in ((1, "b"), (2, "a again"), (3, "b again")):
    for k in range(len(imgs)):
        x, y = runs[0][k], runs[i][k]
        same = x == y
        print(f"image {k} a vs {name}:", "identical" if same else f"DIFFERENT boxes {x[0] == y[0]} scores {x[1]} {y[1]} stab {x[2] == y[2]} rle {x[3] == y[3]} counts {x[4]} {y[4]}")
