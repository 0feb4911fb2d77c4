"""One ViT-L step of the hot path for ncu (launch list / full captures) and a filter-statistics dump.
usage: profile_step.py [n_steps] [--stats]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from crowdsam_b200 import lib, ops
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor
from crowdsam_b200 import synthetic as weights
import bench

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
arch = os.environ.get("ARCH", "vit_l")
dev = torch.device("cuda", 0)
D, depth, heads, glob = weights.SAM_ARCHS[arch]
sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
dD, dd, dh = weights.DINO_ARCHS[bench.DINO]
dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(bench.DINO), strict=True)
pred = SamPredictor(sam.to(dev), dino.to(dev))
cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": bench.test_cfg(int(os.environ.get("PPB", "1024")))}
model = CrowdSAM(cfg, None, predictor=pred)
imgs = [torch.as_tensor(weights.synthetic_image(i)).permute(2, 0, 1).contiguous().to(dev) for i in range(n_steps)]
if "--stats" in sys.argv:
    pred.set_torch_image(imgs[0][None], (1024, 1024))
    pts = weights.grid_points(32)[::4]
    coords = torch.as_tensor(pred.transform.apply_coords(pts, (1024, 1024)))[:, None, :]
    labels = torch.ones(len(pts), dtype=torch.int)[:, None]
    low, iou, cls = pred.decode_low_res(coords, labels)
    score, sel, cat = ops.select_candidates(iou, cls)
    counts, boxes = ops.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
    stab = (counts[:, 0] / counts[:, 1]).cpu()
    q = lambda t: [round(float(x), 4) for x in torch.quantile(t.float().cpu(), torch.tensor([0., .1, .5, .9, 1.]))]
    print("low std", float(low.std()), "features std", float(pred.features.std()), "dino std", float(pred.dino_feats.std()))
    print("stability q", q(stab), "pass>=0.8", float((stab >= 0.8).float().mean()))
    print("score q", q(score), "pass>0.1", float((score > 0.1).float().mean()))
    print("iou q", q(iou.flatten()), "cls q", q(cls.flatten()))
    sys.exit(0)
for i in range(n_steps):
    np.random.seed(42)
    l0 = lib.launch_count()
    if i == n_steps - 1:
        torch.cuda.nvtx.range_push("step")
    model.run_resident(imgs[i])
    torch.cuda.synchronize()
    if i == n_steps - 1:
        torch.cuda.nvtx.range_pop()
    print("step", i, "launches", lib.launch_count() - l0, "counts", model.last_counts)
