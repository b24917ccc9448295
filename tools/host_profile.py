"""cProfile of the host side of one lift step (device-resident inputs)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S
from veon_b200.view_transformer import LSSViewTransformer
cfg = S.CONFIGS["C2"]; B = 8; C = 64; dev = torch.device("cuda", 0)
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, 16, 8, C, collapse_z=False)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
cal = S.calibration(cfg, batch=B); metas = [torch.from_numpy(cal[k]).to(dev) for k in KEYS]
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
g = torch.Generator(device=dev).manual_seed(0)
dd = torch.softmax(torch.randn(B*N, D, H, W, device=dev, generator=g)*4, 1); ff = torch.randn(B*N, C, H, W, device=dev, generator=g)
og = torch.randn(B, C, 16, 200, 200, device=dev, generator=g); img = torch.zeros(B, N, 1, H, W, device=dev)
def step():
    d = dd.detach().requires_grad_(); f = ff.detach().requires_grad_()
    bev, _ = neck.view_transform([img] + metas, d, f); bev.backward(og)
for _ in range(20): step()
torch.cuda.synchronize()
# pure host time: how long does it take to ISSUE n steps (GPU far behind is fine)
n = 200; t0 = time.perf_counter()
for _ in range(n): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue time {1e3*(t1-t0)/n:.3f} ms/step; wall incl. drain {1e3*(t2-t0)/n:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
