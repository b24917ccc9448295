"""Raw-neck training step (forward + backward through pool and the 2x2x2 max) at C2:
one autograd node (PoolMaxDown) vs pool node + MaxDown2x2x2 node."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S, bev_pool as BP
from veon_b200.view_transformer import LSSViewTransformerRaw
cfg = S.CONFIGS["C2"]; B = cfg.batch; C = cfg.channels; dev = torch.device("cuda", 0)
neck = LSSViewTransformerRaw(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
cal = S.calibration(cfg, batch=B); metas = [torch.from_numpy(cal[k]).to(dev) for k in KEYS]
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B, N, D, H, W, device=dev, generator=g) * 4, dim=2)
feat = torch.randn(B, N, C, H, W, device=dev, generator=g)
go = torch.randn(B, C, 8, 100, 100, device=dev, generator=g)
def step():
    f = feat.detach().requires_grad_(); d = depth.detach().requires_grad_()
    neck([f] + metas, d).backward(go)
    return f.grad, d.grad
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
a = step(); ms_one = timeit(step)
neck._forward_pool_maxdown = lambda *a, **k: None      # plain route: two autograd nodes
b = step(); ms_two = timeit(step)
print(f"Raw-neck training step C2 (geometry + prepare + fwd + bwd): one node {ms_one*1e3:.0f} us, "
      f"two nodes {ms_two*1e3:.0f} us; same grads: {torch.equal(a[0], b[0])} {torch.equal(a[1], b[1])}")
