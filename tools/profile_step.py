"""Minimal driver for ncu: a few lift steps (prepare + fwd + bwd) of one workload.
    python tools/profile_step.py [workload] [steps] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from veon_b200 import synthetic as S  # noqa: E402
from veon_b200.view_transformer import LSSViewTransformer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = S.CONFIGS[name]
B = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.batch
C = cfg.channels
dev = torch.device("cuda", 0)
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C,
                          collapse_z=False, sync_free=True)
coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B)).to(dev)
_, N, D, H, W, _ = coor.shape
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B, N, D, H, W, device=dev, generator=g) * 4, dim=2)
feat = torch.randn(B, N, C, H, W, device=dev, generator=g)
og = torch.randn(B, C, 16, 200, 200, device=dev, generator=g)
for _ in range(steps):
    d = depth.detach().requires_grad_()
    f = feat.detach().requires_grad_()
    bev = neck.voxel_pooling_v2(coor, d, f)
    bev.backward(og)
torch.cuda.synchronize()
print("ok", float(bev.detach().abs().mean()), float(d.grad.abs().mean()), float(f.grad.abs().mean()))
