"""Host-side profile of CrowdSAM.generate(ndarray) on the bench workload: where does the end-to-end leg lose time
against the device-resident leg?  (cProfile; the entries that block on the GPU show up as .cpu()/synchronize time.)"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from crowdsam_b200 import lib
from crowdsam_b200 import synthetic as weights
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor

dev = torch.device("cuda", 0)
lib.load()
arch = "vit_l"
D, depth, heads, glob = weights.SAM_ARCHS[arch]
sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
dD, dd, dh = weights.DINO_ARCHS[bench.DINO]
dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(bench.DINO), strict=True)
pred = SamPredictor(sam.to(dev), dino.to(dev))
cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": bench.test_cfg(1024)}
model = CrowdSAM(cfg, None, predictor=pred)
imgs = [weights.synthetic_image(i) for i in range(8)]
pinned = [torch.as_tensor(im).pin_memory() for im in imgs]
res = [torch.as_tensor(im).permute(2, 0, 1).contiguous().to(dev) for im in imgs]
for i in range(3):
    np.random.seed(42); model.generate(pinned[i].numpy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3, 8):
    np.random.seed(42); model.run_resident(res[i])
torch.cuda.synchronize()
t1 = time.perf_counter()
for i in range(3, 8):
    np.random.seed(42); model.generate(pinned[i].numpy())
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"resident {1e3*(t1-t0)/5:.2f} ms/image, generate {1e3*(t2-t1)/5:.2f} ms/image")
pr = cProfile.Profile()
pr.enable()
for i in range(3, 8):
    np.random.seed(42); model.generate(pinned[i].numpy())
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
