"""Stage times of the logit-space lift + classify pipeline (veon_b200.pipeline.lift_classify) on
one GPU, C3 geometry (6 cams 32x88, D=88, C=512), Q=18.  Also the program to put under
`ncu --metrics gpu__time_duration.sum` for the launch list.

    python tools/pipeline_profile.py [samples] [iters] [channel_pad]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import bev_pool as BP
from veon_b200 import synthetic as S
from veon_b200 import tail as T
from veon_b200.pipeline import lift_classify, _heads
from veon_b200.view_transformer import LSSViewTransformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pad = int(sys.argv[3]) if len(sys.argv) > 3 else 4
Q_ARG = int(sys.argv[4]) if len(sys.argv) > 4 else 18
dev = torch.device("cuda", 0)
cfg = S.CONFIGS["C3"]; C = cfg.channels; Q = Q_ARG
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, collapse_z=False)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
cal = S.calibration(cfg, batch=B)
metas = [torch.from_numpy(cal[k]).to(dev) for k in KEYS]
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B * N, D, H, W, device=dev, generator=g) * 4, 1)
feat = torch.randn(B * N, C, H, W, device=dev, generator=g) * 0.05
w = torch.randn(Q, C, device=dev, generator=g); w = 100 * w / w.norm(dim=1, keepdim=True)
gate_w = torch.randn(2, C, device=dev, generator=g)
SIZES = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
cls = T.class_of_prompt(list(range(Q - 1)) if Q != 67 else [k for k, n in enumerate(SIZES) for _ in range(n)]).to(dev)
img = torch.zeros(B, N, 1, H, W, device=dev)


def ev_ms(fn, n):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


whole = ev_ms(lambda: lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w, channel_pad=pad), iters)
rows = _heads(w, gate_w, pad)
t_px = ev_ms(lambda: T.semantic_inference_3d(rows, feat.view(B * N, C, 1, H, W)), iters)
px = T.semantic_inference_3d(rows, feat.view(B * N, C, 1, H, W)).view(B, N, rows.shape[0], H, W)
d5 = depth.view(B, N, D, H, W)
BP.enable_kernel_timing(True)
with torch.no_grad():
    for _ in range(iters):
        vol = neck._voxel_pooling_calib(metas, d5, px)
torch.cuda.synchronize()
kt = {k: round(sum(v[2:]) / max(len(v) - 2, 1), 4) for k, v in BP.kernel_timings_ms().items()}
BP.enable_kernel_timing(False)
with torch.no_grad():
    t_lift = ev_ms(lambda: neck._voxel_pooling_calib(metas, d5, px), iters)
t_cls = ev_ms(lambda: T.classify_logits(vol[:, :Q], vol[:, Q:Q + 2], cls), iters)
print(f"B={B} C={C} Q={Q} channels pooled={rows.shape[0]}: whole {whole*1e3:.1f} us "
      f"({B/whole*1e3:.0f} samples/s) = pixel logits {t_px*1e3:.1f} + lift {t_lift*1e3:.1f} "
      f"(kernel timings ms: {kt}) + classify {t_cls*1e3:.1f} "
      f"({4.0*B*(Q+2)*640000/t_cls/1e6:.0f} GB/s)", flush=True)
