"""Where does CrowdSAM.generate(ndarray) lose time against the device-resident leg of bench.py?  Variants of the same
image loop, 10 images each (wall clock, synchronised at both ends)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from crowdsam_b200 import amg, lib
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
N = 10
imgs = [weights.synthetic_image(i) for i in range(N)]
pinned = [torch.as_tensor(im).pin_memory() for im in imgs]
res = [torch.as_tensor(im).permute(2, 0, 1).contiguous().to(dev) for im in imgs]
for i in range(3):
    np.random.seed(42); model.generate(pinned[i].numpy())


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        np.random.seed(42)
        fn(i)
    torch.cuda.synchronize()
    print(f"{name:58s} {1e3 * (time.perf_counter() - t0) / N:7.2f} ms/image", flush=True)


def resident_sync(i):
    model.run_resident(res[i]); torch.cuda.synchronize()


def resident_rle(i):
    d = model.run_resident(res[i], encode_rle=True)
    if d is not None:
        del d["iou_preds"]
        d["rles"] = amg.coco_encode_rles(d["rles"])
        d.to_numpy()


def set_image_only(i):
    pred.set_image(pinned[i].numpy()); torch.cuda.synchronize()


def set_resident_only(i):
    pred.set_torch_image(res[i][None], (1024, 1024)); torch.cuda.synchronize()


for rep in range(2):
    timed("A resident, images back to back (bench value leg)", lambda i: model.run_resident(res[i]))
    timed("B resident + synchronize per image", resident_sync)
    timed("C resident + RLE + result to numpy", resident_rle)
    timed("D generate(pinned ndarray) (bench e2e leg)", lambda i: model.generate(pinned[i].numpy()))
    timed("E generate(pageable ndarray)", lambda i: model.generate(imgs[i]))
    timed("F set_image(pinned ndarray) + sync", set_image_only)
    timed("G set_torch_image(resident) + sync", set_resident_only)
