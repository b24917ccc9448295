"""One forward call per iteration on the C2 workload (for ncu): python tools/fwd_profile.py [cfg B C iters]"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from fwd_check import setup  # noqa: E402
from veon_b200 import bev_pool as BP  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prep, depth, feat, shape = setup(cfg, B, C)
for _ in range(iters):
    out = BP._fwd_planar(depth, feat, prep.ranks_depth, prep.ranks_feat, prep.ranks_bev, prep.plan,
                         B, C, 640000, shape)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
