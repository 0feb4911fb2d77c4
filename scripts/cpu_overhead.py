"""How much of a step is host-side launch overhead?  Times enqueue-only vs synchronised for the encoder part
(set_image) and the prompt part of one image."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from crowdsam_b200 import lib, ops
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor
from oracle import weights

dev = torch.device("cuda", 0)
lib.load()
arch = bench.ARCH
D, depth, heads, glob = weights.SAM_ARCHS[arch]
sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
dD, ddepth, dheads = weights.DINO_ARCHS[bench.DINO]
dino = DinoVisionTransformer(dD, ddepth, dheads); dino.load_state_dict(weights.make_dino_state(bench.DINO), strict=True)
pred = SamPredictor(sam.to(dev), dino.to(dev))
cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": bench.test_cfg(1024)}
model = CrowdSAM(cfg, None, predictor=pred)
imgs = [torch.as_tensor(weights.synthetic_image(i)).permute(2, 0, 1).contiguous().to(dev) for i in range(6)]
for i in range(3):
    np.random.seed(42); model.run_resident(imgs[i])
torch.cuda.synchronize()
for i in range(3, 6):
    np.random.seed(42)
    l0 = lib.launch_count()
    t0 = time.perf_counter()
    pred.set_torch_image(imgs[i][None], (1024, 1024))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    l1 = lib.launch_count()
    model.orig_image = np.empty((1024, 1024, 0), dtype=np.uint8); model.image, model.downscale = model.orig_image, 1.0
    d = model._run_prompts([0, 0, 1024, 1024], encode_rle=False)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    l2 = lib.launch_count()
    print(f"set_image: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms, {l1-l0} launches | prompts: total {1e3*(t3-t2):.1f} ms, {l2-l1} launches")
# pure ctypes call overhead
L = lib.load()
t0 = time.perf_counter()
for _ in range(20000): L.csam_abi_version()
print(f"ctypes call: {1e6*(time.perf_counter()-t0)/20000:.2f} us")
x = torch.randn(4096, 1024, device=dev)
t0 = time.perf_counter()
for _ in range(2000): ops.layernorm(x, normalize=False, want_h16=True, split=True)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"layernorm op enqueue: {1e6*(t1-t0)/2000:.1f} us")
