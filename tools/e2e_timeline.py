"""GPU timeline of the end-to-end loop bench.py times (pinned host -> view_transform -> backward ->
pinned host), from CUDA events on the three streams: where does a step wait?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import synthetic as S
from veon_b200.view_transformer import LSSViewTransformer
cfg = S.CONFIGS["C2"]; B = 8; C = 64; dev = torch.device("cuda", 0)
sync_free = "--sync-free" in sys.argv
neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, 16, 8, C, collapse_z=False, sync_free=sync_free)
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
N, D = cfg.n_cams, cfg.D; H, W = cfg.feat_hw
g = torch.Generator().manual_seed(0)
cal = S.calibration(cfg, batch=B)
parts = [torch.from_numpy(cal[k]).reshape(-1) for k in KEYS]; shapes = [tuple(cal[k].shape) for k in KEYS]
packed = torch.cat(parts).pin_memory()
hd = torch.softmax(torch.randn(B*N, D, H, W, generator=g)*4, 1).pin_memory(); hf = torch.randn(B*N, C, H, W, generator=g).pin_memory()
host = (packed, hd, hf)
og = torch.randn(B, C, 16, 200, 200, device=dev); img = torch.zeros(B, N, 1, H, W, device=dev)
def unpack(p):
    out, o = [], 0
    for shp in shapes:
        n = 1
        for v in shp: n *= v
        out.append(p[o:o+n].view(shp)); o += n
    return out
copy_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ready = [torch.cuda.Event(), torch.cuda.Event()]; consumed = [torch.cuda.Event(), torch.cuda.Event()]
dgh = [torch.empty((B*N, D, H, W)).pin_memory() for _ in range(2)]; fgh = [torch.empty((B*N, C, H, W)).pin_memory() for _ in range(2)]
dev_in = [tuple(x.to(dev) for x in host) for _ in range(2)]
T = lambda: torch.cuda.Event(enable_timing=True)
log = []
def prefetch(i, rec):
    slot = i % 2
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(consumed[slot])
        a = T(); a.record(copy_s)
        if "--no-h2d" not in sys.argv:
            for dst, src in zip(dev_in[slot], host): dst.copy_(src, non_blocking=True)
        b = T(); b.record(copy_s)
        ready[slot].record(copy_s)
    if rec: log.append(("h2d", i, a, b))
gated = "--gated" in sys.argv
def run(steps, rec=False):
    if gated:
        return run_gated(steps, rec)
    main = torch.cuda.current_stream(dev)
    for ev in consumed: ev.record(main)
    t0 = T(); t0.record(main)
    cpu = []
    prefetch(0, rec)
    for i in range(steps):
        c0 = time.perf_counter()
        slot = i % 2
        if i + 1 < steps: prefetch(i + 1, rec)
        main.wait_event(ready[slot])
        p, depth, feat = dev_in[slot]
        a = T(); a.record(main)
        depth = depth.detach().requires_grad_(); feat = feat.detach().requires_grad_()
        bev, _ = neck.view_transform([img] + unpack(p), depth, feat)
        m = T(); m.record(main)
        bev.backward(og)
        b = T(); b.record(main)
        consumed[slot].record(main)
        dg, fg = depth.grad, feat.grad
        with torch.cuda.stream(d2h_s):
            d2h_s.wait_event(consumed[slot])
            dg.record_stream(d2h_s); fg.record_stream(d2h_s)
            x = T(); x.record(d2h_s)
            if "--no-d2h" not in sys.argv:
                dgh[slot].copy_(dg, non_blocking=True); fgh[slot].copy_(fg, non_blocking=True)
            y = T(); y.record(d2h_s)
        cpu.append(time.perf_counter() - c0)
        if rec: log.append(("fwd", i, a, m)); log.append(("bwd", i, m, b)); log.append(("d2h", i, x, y))
    main.wait_stream(d2h_s)
    t1 = T(); t1.record(main)
    torch.cuda.synchronize()
    return t0, t0.elapsed_time(t1) / steps, sum(cpu) / len(cpu) * 1e3
def run_gated(steps, rec=False):
    """copies are released from the neck's prepared_hook: H2D of step i+1 and D2H of step
    i-1 start when prepare(i) is done and overlap the pooling kernels only"""
    main = torch.cuda.current_stream(dev)
    for ev in consumed: ev.record(main)
    t0 = T(); t0.record(main)
    state = {"i": 0, "pending": None}
    gate = torch.cuda.Event()
    def issue_d2h(pending):
        slot, dg, fg, i = pending
        with torch.cuda.stream(d2h_s):
            d2h_s.wait_event(consumed[slot]); d2h_s.wait_event(gate)
            dg.record_stream(d2h_s); fg.record_stream(d2h_s)
            x = T(); x.record(d2h_s)
            dgh[slot].copy_(dg, non_blocking=True); fgh[slot].copy_(fg, non_blocking=True)
            y = T(); y.record(d2h_s)
        if rec: log.append(("d2h", i, x, y))
    def hook():
        gate.record(main)
        i = state["i"]
        if i + 1 < steps:
            slot = (i + 1) % 2
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(consumed[slot]); copy_s.wait_event(gate)
                a = T(); a.record(copy_s)
                for dst, src in zip(dev_in[slot], host): dst.copy_(src, non_blocking=True)
                b = T(); b.record(copy_s)
                ready[slot].record(copy_s)
            if rec: log.append(("h2d", i + 1, a, b))
        if state["pending"] is not None:
            issue_d2h(state["pending"]); state["pending"] = None
    neck.prepared_hook = hook
    cpu = []
    prefetch(0, rec)
    for i in range(steps):
        c0 = time.perf_counter()
        slot = i % 2; state["i"] = i
        main.wait_event(ready[slot])
        p, depth, feat = dev_in[slot]
        a = T(); a.record(main)
        depth = depth.detach().requires_grad_(); feat = feat.detach().requires_grad_()
        bev, _ = neck.view_transform([img] + unpack(p), depth, feat)
        m = T(); m.record(main)
        bev.backward(og)
        b = T(); b.record(main)
        consumed[slot].record(main)
        state["pending"] = (slot, depth.grad, feat.grad, i)
        cpu.append(time.perf_counter() - c0)
        if rec: log.append(("fwd", i, a, m)); log.append(("bwd", i, m, b))
    neck.prepared_hook = None
    gate.record(main)
    issue_d2h(state["pending"])
    main.wait_stream(d2h_s)
    t1 = T(); t1.record(main)
    torch.cuda.synchronize()
    return t0, t0.elapsed_time(t1) / steps, sum(cpu) / len(cpu) * 1e3

run(5)
_, ms, cpu = run(30); print(f"e2e {ms:.3f} ms/step, host loop {cpu:.3f} ms/step (sync_free={sync_free})")
t0, ms, cpu = run(8, rec=True)
rows = sorted(((t0.elapsed_time(a), t0.elapsed_time(b), k, i) for k, i, a, b in log))
for s, e, k, i in rows: print(f"{k:4s} step {i}  {s:7.3f} -> {e:7.3f}  ({e-s:.3f})")
# per-call device times inside the e2e loop (copies running) vs a device-only loop
from veon_b200 import bev_pool as BP
import statistics
BP.enable_kernel_timing(True); run(12); tm = BP.kernel_timings_ms(); BP.enable_kernel_timing(False)
print("with copies :", {k: round(statistics.median(v[2:]), 4) for k, v in tm.items()})
p, depth, feat = dev_in[0]
def dev_step():
    d = depth.detach().requires_grad_(); f = feat.detach().requires_grad_()
    bev, _ = neck.view_transform([img] + unpack(p), d, f); bev.backward(og)
for _ in range(3): dev_step()
BP.enable_kernel_timing(True)
for _ in range(12): dev_step()
tm = BP.kernel_timings_ms(); BP.enable_kernel_timing(False)
print("device only :", {k: round(statistics.median(v[2:]), 4) for k, v in tm.items()})
if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); run(100); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(32)
if "--trace" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    run(5)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        run(6)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", "e2e_trace.json"))
    print("trace written")
if "--cpu" in sys.argv:
    blocked = [0.0]
    orig = torch.cuda.Event.synchronize
    def timed_sync(self):
        t = time.perf_counter(); orig(self); blocked[0] += time.perf_counter() - t
    torch.cuda.Event.synchronize = timed_sync
    run(5); blocked[0] = 0.0
    t = time.perf_counter(); run(50); torch.cuda.synchronize(); wall = time.perf_counter() - t
    print(f"wall {wall/50*1e3:.3f} ms/step, blocked in Event.synchronize {blocked[0]/50*1e3:.3f} ms/step, "
          f"host busy {(wall-blocked[0])/50*1e3:.3f} ms/step")
