"""How fast can the forward's store pattern go?  Times k_pool_fwd on (a) the C2 workload,
(b) the same shapes with NO point inside the grid (pure zero-fill through the same stores),
and a plain torch fill of the same volume for reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from veon_b200 import bev_pool as BP, synthetic as S

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = S.CONFIGS[name]; B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch; C = cfg.channels
dev = torch.device("cuda", 0)
lower, interval, size = S.grid_vectors(cfg.grid_config)
coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B)).to(dev)
_, N, D, H, W, _ = coor.shape
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B, N, D, H, W, device=dev, generator=g) * 4, dim=2)
feat = torch.randn(B, N, H, W, C, device=dev, generator=g)
shape = (B, 16, 200, 200, C)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

vol_gb = 4 * 640000 * C * B / 1e9
for label, c in (("real", coor), ("empty", coor + 1000.0)):
    prep = BP.prepare_ranks(c, lower, interval, size)
    torch.cuda.synchronize()
    ms = timeit(lambda: BP.pool_prepared(depth, feat, prep, shape))
    print(f"{label:6s} fwd {ms*1e3:8.1f} us  -> {vol_gb/ms*1e3:7.1f} GB/s of volume  (kept={prep.plan.n_points})")
buf = torch.empty(B, C, 16, 200, 200, device=dev)
ms = timeit(lambda: buf.fill_(0.0))
print(f"torch fill_ {ms*1e3:8.1f} us -> {vol_gb/ms*1e3:7.1f} GB/s")
src = torch.empty_like(buf)
ms = timeit(lambda: buf.copy_(src))
print(f"torch copy_ {ms*1e3:8.1f} us -> {2*vol_gb/ms*1e3:7.1f} GB/s (r+w)")
