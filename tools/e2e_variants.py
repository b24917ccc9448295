"""Which leg of the end-to-end loop fails to overlap?  Same loop as bench.py's e2e, with the
H2D and/or D2H legs switched off."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S
from veon_b200.view_transformer import LSSViewTransformer
cfg = S.CONFIGS["C2"]; B = 8; C = 64; dev = torch.device("cuda", 0)
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, 16, 8, C, collapse_z=False, sync_free=bool(int(os.environ.get("SF", "0"))))
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
cal = S.calibration(cfg, batch=B)
parts = [torch.from_numpy(cal[k]).reshape(-1) for k in KEYS]; shapes = [tuple(cal[k].shape) for k in KEYS]
packed_h = torch.cat(parts).pin_memory()
g = torch.Generator().manual_seed(0)
hd = torch.softmax(torch.randn(B*N, D, H, W, generator=g)*4, 1).pin_memory(); hf = torch.randn(B*N, C, H, W, generator=g).pin_memory()
og = torch.randn(B, C, 16, 200, 200, device=dev); img = torch.zeros(B, N, 1, H, W, device=dev)
dev_in = [(torch.empty_like(packed_h, device=dev), torch.empty_like(hd, device=dev), torch.empty_like(hf, device=dev)) for _ in range(2)]
for d_ in dev_in:
    for dst, src in zip(d_, (packed_h, hd, hf)): dst.copy_(src)
dgh = [torch.empty(hd.shape).pin_memory() for _ in range(2)]; fgh = [torch.empty(hf.shape).pin_memory() for _ in range(2)]
cs, ds = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
dg_dev = [torch.empty(hd.shape, device=dev) for _ in range(2)]; fg_dev = [torch.empty(hf.shape, device=dev) for _ in range(2)]
grads_done = [torch.cuda.Event(), torch.cuda.Event()]; STAGE = int(os.environ.get('STAGE', '0'))
ready = [torch.cuda.Event(), torch.cuda.Event()]; consumed = [torch.cuda.Event(), torch.cuda.Event()]
def unpack(pk):
    out, o = [], 0
    for shp in shapes:
        n = 1
        for v in shp: n *= v
        out.append(pk[o:o+n].view(shp)); o += n
    return out
def run(steps, h2d, d2h):
    main = torch.cuda.current_stream(dev)
    for ev in consumed: ev.record(main)
    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(cs):
            cs.wait_event(consumed[slot])
            if h2d:
                for dst, src in zip(dev_in[slot], (packed_h, hd, hf)): dst.copy_(src, non_blocking=True)
            ready[slot].record(cs)
    prefetch(0)
    for i in range(steps):
        slot = i % 2
        if i + 1 < steps: prefetch(i + 1)
        main.wait_event(ready[slot])
        pk, depth, feat = dev_in[slot]
        depth = depth.detach().requires_grad_(); feat = feat.detach().requires_grad_()
        bev, _ = neck.view_transform([img] + unpack(pk), depth, feat); bev.backward(og)
        consumed[slot].record(main)
        if i == 0 and d2h == 3: print('dg contiguous', depth.grad.is_contiguous(), depth.grad.shape, 'fg contiguous', feat.grad.is_contiguous(), feat.grad.stride())
        if d2h and STAGE:
            dg_dev[slot].copy_(depth.grad); fg_dev[slot].copy_(feat.grad); grads_done[slot].record(main)
            with torch.cuda.stream(ds):
                ds.wait_event(grads_done[slot])
                if d2h & 1: dgh[slot].copy_(dg_dev[slot], non_blocking=True)
                if d2h & 2: fgh[slot].copy_(fg_dev[slot], non_blocking=True)
        elif d2h:
            dg, fg = depth.grad, feat.grad
            with torch.cuda.stream(ds):
                ds.wait_event(consumed[slot]); dg.record_stream(ds); fg.record_stream(ds)
                if d2h & 1: dgh[slot].copy_(dg, non_blocking=True)
                if d2h & 2: fgh[slot].copy_(fg, non_blocking=True)
    main.wait_stream(ds)
print('grad layouts', None)
for h2d, d2h in ((0, 0), (0, 3), (1, 3)):
    run(5, h2d, d2h); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(40, h2d, d2h); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"h2d={h2d} d2h={d2h}: {(t1-t0)/40*1e3:.3f} ms/step")
