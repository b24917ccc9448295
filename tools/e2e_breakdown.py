import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S
from veon_b200.view_transformer import LSSViewTransformer
cfg = S.CONFIGS["C2"]; B = 8; C = 64; dev = torch.device("cuda", 0)
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, 16, 8, C, collapse_z=False)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
cal = S.calibration(cfg, batch=B); metas_h = [torch.from_numpy(cal[k]).pin_memory() for k in KEYS]
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
g = torch.Generator().manual_seed(0)
hd = torch.softmax(torch.randn(B*N, D, H, W, generator=g)*4, 1).pin_memory(); hf = torch.randn(B*N, C, H, W, generator=g).pin_memory()
og = torch.randn(B, C, 16, 200, 200, device=dev)
img = torch.zeros(B, N, 1, H, W, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
metas = [m.to(dev) for m in metas_h]; dd = hd.to(dev); ff = hf.to(dev)
print("h2d depth+feat+metas  ms", t(lambda: ([m.to(dev, non_blocking=True) for m in metas_h], hd.to(dev, non_blocking=True), hf.to(dev, non_blocking=True))))
print("get_lidar_coor        ms", t(lambda: neck.get_lidar_coor(*metas)))
coor = neck.get_lidar_coor(*metas)
def lift():
    d = dd.detach().requires_grad_(); f = ff.detach().requires_grad_()
    bev = neck.voxel_pooling_v2(coor, d.view(B, N, D, H, W), f.view(B, N, C, H, W)); bev.backward(og); return d, f
print("voxel_pooling_v2+bwd  ms", t(lift))
def vt():
    d = dd.detach().requires_grad_(); f = ff.detach().requires_grad_()
    bev, _ = neck.view_transform([img] + metas, d, f); bev.backward(og); return d, f
print("view_transform+bwd    ms", t(vt))
d, f = vt()
dgh = torch.empty(d.grad.shape).pin_memory(); fgh = torch.empty(f.grad.shape).pin_memory()
print("feat.grad contiguous?", f.grad.is_contiguous(), d.grad.is_contiguous())
print("d2h grads             ms", t(lambda: (dgh.copy_(d.grad, non_blocking=True), fgh.copy_(f.grad, non_blocking=True))))
