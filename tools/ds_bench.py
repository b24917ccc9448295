"""Raw neck forward (no grad): fused pool + 2x2x2 max vs pool followed by torch amax."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S
from veon_b200.view_transformer import LSSViewTransformerRaw
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = S.CONFIGS[name]; B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch; C = cfg.channels
dev = torch.device("cuda", 0)
neck = LSSViewTransformerRaw(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, sync_free=True, fuse_ds=True)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
cal = S.calibration(cfg, batch=B); metas = [torch.from_numpy(cal[k]).to(dev) for k in KEYS]
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B, N, D, H, W, device=dev, generator=g) * 4, dim=2)
feat = torch.randn(B, N, C, H, W, device=dev, generator=g)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    a = neck([feat] + metas, depth)
    ms_f = timeit(lambda: neck([feat] + metas, depth))
    neck.fuse_ds = False           # pool, then the streaming 2x2x2 kernel
    b_ = neck([feat] + metas, depth)
    ms_p = timeit(lambda: neck([feat] + metas, depth))
    from veon_b200 import bev_pool as BP
    BP.MaxDown2x2x2.supports = staticmethod(lambda x: False)   # the reference's ATen expression
    c_ = neck([feat] + metas, depth)
    ms_a = timeit(lambda: neck([feat] + metas, depth), n=5)
print(f"{name} B={B} C={C} Raw-neck forward incl. get_lidar_coor + prepare: fused kernel {ms_f*1e3:.1f} us, "
      f"pool + k_maxdown2 {ms_p*1e3:.1f} us, pool + ATen amax {ms_a*1e3:.1f} us; "
      f"equal: {torch.equal(a, b_)} {torch.equal(a, c_)}")
