"""Run one BASELINE.json configuration end to end on the GPU and print its timing (functional check of the
non-headline configs: grid 64 = 4096 prompts, ViT-B, ViT-H).  usage: run_config.py ARCH GRID [points_per_batch]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from crowdsam_b200 import lib
from crowdsam_b200.build import _build_sam
from crowdsam_b200.modules import DinoVisionTransformer
from crowdsam_b200.pipeline import CrowdSAM
from crowdsam_b200.predictor import SamPredictor
from oracle import weights

arch, grid = sys.argv[1], int(sys.argv[2])
ppb = int(sys.argv[3]) if len(sys.argv) > 3 else grid * grid
bench.GRID = grid
dev = torch.device("cuda", 0)
lib.load()
D, depth, heads, glob = weights.SAM_ARCHS[arch]
sam = _build_sam(D, depth, heads, 1, glob); sam.load_state_dict(weights.make_sam_state(arch), strict=True)
dD, dd, dh = weights.DINO_ARCHS[bench.DINO]
dino = DinoVisionTransformer(dD, dd, dh); dino.load_state_dict(weights.make_dino_state(bench.DINO), strict=True)
pred = SamPredictor(sam.to(dev), dino.to(dev))
cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": bench.test_cfg(ppb)}
model = CrowdSAM(cfg, None, predictor=pred)
imgs = [torch.as_tensor(weights.synthetic_image(i)).permute(2, 0, 1).contiguous().to(dev) for i in range(4)]
for i in range(4):
    np.random.seed(42)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.run_resident(imgs[i])
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    print(f"{arch} grid {grid} ppb {ppb}: step {i} {ms:.1f} ms, masks into NMS / kept {model.last_counts}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
