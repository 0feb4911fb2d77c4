"""K-POST write at P = 1024 (all-survive), for the ncu traffic capture."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from crowdsam_b200 import ops
g = torch.Generator().manual_seed(0)
low = (torch.randn(256, 4, 64, 64, generator=g) * 8).cuda()
low = torch.nn.functional.interpolate(low, (256, 256), mode="nearest").repeat(4, 1, 1, 1).contiguous()
sel = torch.randint(0, 4, (1024,), generator=g).to(torch.int32).cuda()
for _ in range(3):
    m, _ = ops.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
torch.cuda.synchronize()
print("ok", m.shape)
