"""Two engine instances with identical weights must agree bit for bit at every stage (run it under
`compute-sanitizer --tool memcheck` to stretch the kernels' timing).  usage: debug_determinism.py [arch] [reps]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from crowdsam_b200 import synthetic as weights
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.predictor import SamPredictor

arch = sys.argv[1] if len(sys.argv) > 1 else "tiny_l"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dino_arch = "tiny" if arch.startswith("tiny") else "dinov2_vitl14"
dev = torch.device("cuda", 0)


def make():
    D, depth, heads, glob = weights.SAM_ARCHS[arch]
    sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
    dD, dd, dh = weights.DINO_ARCHS[dino_arch]
    dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(dino_arch), strict=True)
    return SamPredictor(sam.to(dev), dino.to(dev))


a, b = make(), make()
pts = weights.grid_points(16)
for rep in range(reps):
    img = weights.synthetic_image(50 + rep)
    out = []
    for p in (a, b, a):
        p.set_image(img)
        coords = torch.as_tensor(p.transform.apply_coords(pts, p.original_size))[:, None, :]
        labels = torch.ones(len(pts), dtype=torch.int)[:, None]
        low, iou, cls = p.decode_low_res(coords, labels)
        torch.cuda.synchronize()
        out.append(dict(features=p.features.clone(), dino=p.dino_feats.clone(), fg=p.predict_fg_map().clone(), low=low.clone(),
                        iou=iou.clone(), cls=cls.clone()))
    for name, (x, y) in (("a-vs-b", (out[0], out[1])), ("a-vs-a", (out[0], out[2]))):
        bad = {k: float((x[k] - y[k]).abs().max()) for k in x if not torch.equal(x[k], y[k])}
        print(f"rep {rep} {name}:", "identical" if not bad else f"DIFFERENT {bad}")
