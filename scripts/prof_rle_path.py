"""Breakdown of the result path of CrowdSAM.generate after NMS (RLE of the kept masks, COCO strings, result to numpy) on
the bench workload, with the GPU idle at the start of each piece."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from crowdsam_b200 import amg, lib, ops
from crowdsam_b200 import synthetic as weights
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor

dev = torch.device("cuda", 0)
lib.load()
D, depth, heads, glob = weights.SAM_ARCHS["vit_l"]
sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state("vit_l"), strict=True)
dD, dd, dh = weights.DINO_ARCHS[bench.DINO]
dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(bench.DINO), strict=True)
pred = SamPredictor(sam.to(dev), dino.to(dev))
cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": bench.test_cfg(1024)}
model = CrowdSAM(cfg, None, predictor=pred)
res = [torch.as_tensor(weights.synthetic_image(i)).permute(2, 0, 1).contiguous().to(dev) for i in range(4)]
acc = {}


def lap(name, t0):
    torch.cuda.synchronize()
    acc.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
    return time.perf_counter()


for i in range(4):
    np.random.seed(42)
    data = model.run_resident(res[i])
    torch.cuda.synchronize()
    masks = data["masks"]
    for nm, m in (("1 mask", masks), ("64 masks", masks.expand(64, -1, -1).contiguous() if masks.shape[0] == 1 else masks[:64])):
        t = time.perf_counter()
        rl = ops.rle_encode(m.to(torch.bool)); t = lap(f"rle_encode [{nm}]", t)
        rles = [{"size": [1024, 1024], "counts": r} for r in rl]
        enc = amg.coco_encode_rles(rles); t = lap(f"coco strings [{nm}]", t)
    t = time.perf_counter()
    d2 = model.run_resident(res[i], encode_rle=True); t = lap("run_resident(encode_rle=True)", t)
    del d2["iou_preds"]; t = time.perf_counter()
    d2["rles"] = amg.coco_encode_rles(d2["rles"]); t = lap("coco strings (pipeline)", t)
    d2.to_numpy(); t = lap("to_numpy", t)
    t = time.perf_counter()
    model.run_resident(res[i]); t = lap("run_resident(encode_rle=False)", t)
for k, v in acc.items():
    print(f"{k:40s} {min(v[1:]):8.3f} ms (min of {len(v) - 1})")
