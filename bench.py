#!/usr/bin/env python
"""bench.py -- 6-camera samples/s of the lift (voxel_pooling_prepare_v2 +
bev_pool_v2 forward + backward) on N B200s, with the roofline of the dominant
kernel and the CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic
nuScenes-shaped input: BASELINE.json configs[1] (6 cams 16x44 feats, D=88,
C=64, batch 8 per GPU; weak scaling: every rank lifts its own 8 samples, no
data-path collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "6-cam samples/sec for lift (bev_pool_v2 fwd+bwd)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", help="veon_b200.synthetic.CONFIGS key")
    ap.add_argument("--batch", type=int, default=0, help="override samples per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0,
                    help="budget of the cpu_baseline leg (rank 0, N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tail", action="store_true", help="skip the secondary tail measurement")
    ap.add_argument("--sync-free", type=int, default=0,
                    help="1: skip the host read-back of the counts (see LSSViewTransformer)")
    return ap.parse_args()


def workload_dict(cfg, batch, extra=None):
    H, W = cfg.feat_hw
    d = {"workload": f"{cfg.name}: {cfg.n_cams} cams {cfg.input_size[0]}x{cfg.input_size[1]} -> "
                     f"{H}x{W} feats, D={cfg.D}, C={cfg.channels} -> 200x200x16, "
                     f"prepare_v2 + bev_pool_v2 fwd + bwd",
         "samples_per_gpu_per_step": batch, "parallelism": "sample-sharded, no collective"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi-equivalent (NVML) clock / throttle sampling during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self._stop = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------ CPU baseline
def cpu_lift_once(cfg, batch, seed=0):
    """inputs for the CPU port (oracle.lift_oracle.torch_cpu_lift)"""
    import torch
    from veon_b200 import synthetic as S
    coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=batch, sample_offset=seed))
    B, N, D, H, W, _ = coor.shape
    g = torch.Generator().manual_seed(seed)
    depth = torch.softmax(torch.randn(B, N, D, H, W, generator=g) * 4, dim=2)
    feat = torch.randn(B, N, cfg.channels, H, W, generator=g)
    og = torch.randn(B, cfg.channels, 16, 200, 200, generator=g)
    return coor, depth, feat, og


def time_cpu_port(cfg, budget_s, steps=None, warmup=1, batch=1):
    """Times the reference's pure-PyTorch CPU lift (BASELINE.json configs[0]) as restated in
    oracle/ (kind "port").  Returns (samples/s, cores, description, ms per step)."""
    import torch
    from oracle import lift_oracle as O
    from veon_b200 import synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor, depth, feat, og = cpu_lift_once(cfg, batch)
    for _ in range(warmup):
        O.torch_cpu_lift(coor, depth, feat, lower, interval, size, og)
    times = []
    t_end = time.perf_counter() + budget_s
    while True:
        t0 = time.perf_counter()
        O.torch_cpu_lift(coor, depth, feat, lower, interval, size, og)
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() > t_end and len(times) >= 3):
            break
    times.sort()
    med = times[len(times) // 2]
    desc = (f"{len(times)} timed passes (median) of prepare_v2 + scatter-add pool fwd + autograd bwd "
            f"on {batch} sample(s) of {cfg.name} (6 cams, D={cfg.D}, C={cfg.channels}), torch CPU fp32, "
            f"{cores} threads")
    return batch / med, cores, desc, med * 1e3, len(times)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the
    reference is Python + CUDA, its CPU path is the pure-PyTorch lift of BASELINE configs[0])."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from veon_b200 import synthetic as S
    cfg = S.CONFIGS[args.workload]
    # each step = ONE sample of the workload (bounded so K steps end in minutes)
    val, cores, desc, ms, n = time_cpu_port(cfg, 0.0, steps=max(args.steps, 1),
                                            warmup=max(args.warmup, 1), batch=1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_dict(cfg, 1, {"note": "CPU port; one sample per step"}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------- ours
def algorithmic_bytes(cfg, B, n_kept, n_int):
    """compulsory traffic per step (every input element read once, every
    output element written once)."""
    V, C = 640000, cfg.channels
    H, W = cfg.feat_hw
    NHW = cfg.n_cams * H * W
    P = NHW * cfg.D * B
    fwd = 4 * V * C * B + 4 * NHW * C * B + 4 * n_kept + 8 * n_kept + 12 * n_int
    n_bp = NHW * B
    bwd = 4 * n_int * C + 4 * NHW * C * B + 4 * n_kept + 4 * P + 4 * NHW * C * B + 12 * n_kept + 8 * n_bp
    prep = 12 * P + 12 * n_kept + 8 * n_int
    return {"prepare": prep, "pool_fwd": fwd, "pool_bwd": bwd}


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads (and so, by first touch, its pinned host buffers) to the
    CPUs of the NUMA node its GPU hangs off.  With 8 ranks each streaming 20 MB per
    direction per step, host buffers on the far socket cost more than anything on the GPU.
    Best effort: returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # nvml pads the PCI domain to 8 hex digits
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from veon_b200 import _lib, bev_pool as BP, synthetic as S
    from veon_b200.view_transformer import LSSViewTransformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    cfg = S.CONFIGS[args.workload]
    B = args.batch or cfg.batch
    C = cfg.channels
    neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, in_channels=8,
                              out_channels=C, collapse_z=False, sync_free=bool(args.sync_free))
    lower, interval, size = S.grid_vectors(cfg.grid_config)

    # rotating input sets so that a step's inputs are not L2-hot from the previous step
    n_sets = 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    host_sets, dev_sets = [], []
    for i in range(n_sets):
        coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B, sample_offset=rank * 1000 + i * B))
        Bc, N, D, H, W, _ = coor.shape
        depth = torch.softmax(torch.randn(B, N, D, H, W, device=dev, generator=g) * 4, dim=2)
        feat = torch.randn(B, N, C, H, W, device=dev, generator=g)
        host_sets.append((coor.pin_memory(), depth.cpu().pin_memory(), feat.cpu().pin_memory()))
        dev_sets.append((coor.to(dev), depth, feat))
    out_grad = torch.randn(B, C, 16, 200, 200, device=dev, generator=g)

    def step_device(i):
        coor, depth, feat = dev_sets[i % n_sets]
        depth = depth.detach().requires_grad_()
        feat = feat.detach().requires_grad_()
        bev = neck.voxel_pooling_v2(coor, depth, feat)
        bev.backward(out_grad)
        return depth.grad, feat.grad

    # ---- end to end: the user's call, view_transform(input, depth, tran_feat), on HOST data.
    # Per step: calibration + depth + feat from pinned host memory -> device (copy stream,
    # overlapped with the previous step's compute), get_lidar_coor + prepare + pool forward +
    # backward on device, depth_grad + feat_grad back to pinned host memory (D2H stream).
    KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")
    e2e_host = []
    meta_shapes = None
    for i in range(n_sets):
        cal = S.calibration(cfg, batch=B, sample_offset=rank * 1000 + i * B)
        # the six calibration tensors travel as ONE pinned buffer (one H2D copy), and are
        # handed to view_transform as views of it
        parts = [torch.from_numpy(cal[k]).reshape(-1) for k in KEYS]
        meta_shapes = [tuple(cal[k].shape) for k in KEYS]
        packed = torch.cat(parts).pin_memory()
        _, hd, hf = host_sets[i]
        e2e_host.append((packed, hd.view(B * N, D, H, W), hf.view(B * N, C, H, W)))

    def unpack_metas(packed_dev):
        out, o = [], 0
        for shp in meta_shapes:
            n = 1
            for v in shp:
                n *= v
            out.append(packed_dev[o:o + n].view(shp))
            o += n
        return out

    img_shape = torch.zeros(B, N, 1, H, W, device=dev)      # only its shape is read
    copy_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    dg_host = [torch.empty((B * N, D, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    fg_host = [torch.empty((B * N, C, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]

    # static device staging buffers (two slots): no allocator traffic on the side streams
    dev_in = [(torch.empty_like(e2e_host[0][0], device=dev),
               torch.empty((B * N, D, H, W), dtype=torch.float32, device=dev),
               torch.empty((B * N, C, H, W), dtype=torch.float32, device=dev)) for _ in range(2)]

    def e2e_h2d(i, gate=None):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the slot's previous user is done
            if gate is not None:
                copy_stream.wait_event(gate)
            for dst, src in zip(dev_in[slot], e2e_host[i % n_sets]):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_d2h(slot, dgrad, fgrad, gate):
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(consumed[slot])
            d2h_stream.wait_event(gate)
            dgrad.record_stream(d2h_stream)
            fgrad.record_stream(d2h_stream)
            dg_host[slot].copy_(dgrad, non_blocking=True)
            fg_host[slot].copy_(fgrad, non_blocking=True)

    def run_e2e(steps):
        """Every step copies its inputs in and its gradients out.  The copies are queued from
        the neck's `prepared_hook`: H2D of step i+1 and D2H of step i-1 are released when the
        index preparation of step i is done, so they overlap the two long pooling kernels
        instead of the dozen short launches before them (those are measurably slower while
        PCIe is saturated; tools/e2e_timeline.py)."""
        main = torch.cuda.current_stream(dev)
        for ev in consumed:
            ev.record(main)
        gate = torch.cuda.Event()
        state = {"i": 0, "pending": None}
        done = [torch.cuda.Event() for _ in range(4)]   # bounds the host's run-ahead to 2 steps

        def hook():
            gate.record(main)
            if state["i"] + 1 < steps:
                e2e_h2d(state["i"] + 1, gate)
            if state["pending"] is not None:
                e2e_d2h(*state["pending"], gate)
                state["pending"] = None

        neck.prepared_hook = hook
        try:
            e2e_h2d(0)
            for i in range(steps):
                slot = i % 2
                state["i"] = i
                if i >= 2:
                    done[(i - 2) % 4].synchronize()
                main.wait_event(ready[slot])
                packed_dev, depth, feat = dev_in[slot]
                metas = unpack_metas(packed_dev)
                depth = depth.detach().requires_grad_()
                feat = feat.detach().requires_grad_()
                bev, _ = neck.view_transform([img_shape] + metas, depth, feat)
                bev.backward(out_grad)
                consumed[slot].record(main)
                done[i % 4].record(main)
                state["pending"] = (slot, depth.grad, feat.grad)
        finally:
            neck.prepared_hook = None
        gate.record(main)
        e2e_d2h(*state["pending"], gate)
        main.wait_stream(d2h_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, whole_loop=False):
        if whole_loop:
            fn(warmup)
        else:
            for i in range(warmup):
                fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.veon_kernel_launch_count()
        e0.record()
        if whole_loop:
            fn(steps)
        else:
            for i in range(steps):
                fn(i)
        e1.record()
        barrier()
        launches = lib.veon_kernel_launch_count() - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches

    K, Wm = max(args.steps, 1), max(args.warmup, 3)
    with ClockSampler(local) as clocks:
        ms_total, launches = timed(step_device, K, Wm)
    ms_e2e, _ = timed(run_e2e, K, max(3, Wm // 2), whole_loop=True)

    # ---- secondary: the same device-resident job with the index preparation of step i+1
    # issued on a second stream, under the pooling kernels of step i.  The preparation only
    # depends on the calibration (not on features or weights), so a training / serving loop can
    # run it one step ahead through the public prepare_ranks() + pool_prepared() pair.  Reported
    # next to `value` (which keeps the strictly sequential step), never instead of it.
    prep_stream = torch.cuda.Stream(dev)
    grid_vecs = (neck.grid_lower_bound, neck.grid_interval, neck.grid_size)

    def run_pipelined(steps):
        main = torch.cuda.current_stream(dev)
        preps, evs = [None, None], [torch.cuda.Event(), torch.cuda.Event()]
        used = [torch.cuda.Event(), torch.cuda.Event()]   # main is done with a slot's plan
        for ev in used:
            ev.record(main)

        def prepare(i):
            with torch.cuda.stream(prep_stream):
                prep_stream.wait_event(used[i % 2])   # its buffers go back to this stream's pool
                preps[i % 2] = None
                preps[i % 2] = BP.prepare_ranks(dev_sets[i % n_sets][0], *grid_vecs)
                evs[i % 2].record(prep_stream)
        prepare(0)
        for i in range(steps):
            if i + 1 < steps:
                prepare(i + 1)
            _, depth, feat = dev_sets[i % n_sets]
            depth = depth.detach().requires_grad_()
            feat = feat.detach().requires_grad_()
            main.wait_event(evs[i % 2])
            prep = preps[i % 2]
            bev = BP.pool_prepared(depth, feat.permute(0, 1, 3, 4, 2), prep,
                                   neck._bev_shape(depth, feat.shape[2]))
            if prep.plan.n_intervals == 0:          # the faithful path's read-back
                raise RuntimeError("no point inside the grid")
            bev.backward(out_grad)
            used[i % 2].record(main)
        prep_stream.wait_stream(main)
    ms_pipe, _ = timed(run_pipelined, K, max(3, Wm // 2), whole_loop=True)

    # live per-call device times (CUDA events on the launching stream) over K more steps
    BP.enable_kernel_timing(True)
    for i in range(K):
        step_device(i)
    t = BP.kernel_timings_ms()
    BP.enable_kernel_timing(False)
    avg = {k: sum(v) / len(v) for k, v in t.items()}

    # counts for the algorithmic-byte denominators
    prep = BP.prepare_ranks(dev_sets[0][0], lower, interval, size)
    n_kept, n_int = prep.plan.n_points, prep.plan.n_intervals
    alg = algorithmic_bytes(cfg, B, n_kept, n_int)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get("pool_fwd_dram_bytes_per_launch")
    except Exception:
        pass
    fwd_ms = avg.get("pool_fwd")
    achieved = alg["pool_fwd"] / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None

    value = world * B * K / (ms_total * 1e-3)
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in e2e_host[0])
    d2h = dg_host[0].numel() * 4 + fg_host[0].numel() * 4

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_dict(cfg, B, {
            "l2": f"inputs rotate over {n_sets} sets; each step streams a "
                  f"{4 * 640000 * C * B / 1e9:.2f} GB volume (>> 126 MB L2)",
            "host_sync_per_step": 0 if args.sync_free else 1,
            "numa_node": numa_node,
            "n_kept": n_kept, "n_intervals": n_int}),
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": "pinned host calibration+depth+feat -> LSSViewTransformer.view_transform "
                        "(get_lidar_coor fused into prepare_v2, bev_pool_v2) -> backward -> depth_grad "
                        "+ feat_grad to pinned host; copies on side streams, double-buffered, "
                        "released from the neck's prepared_hook"},
        "roofline": {"bound": "hbm", "kernel": "k_pool_fwd + k_pool_fwd_heavy (one veon_bev_pool_v2_fwd_planar call; the heavy-tile grid runs in the tail of the main grid)",
                     "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": (achieved / peak_gbs) if achieved else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg["pool_fwd"],
                     "ms_per_launch": fwd_ms, "traffic": traffic},
        "value_prepare_overlapped": {
            "value": world * B * K / (ms_pipe * 1e-3), "unit": UNIT, "ms_per_step": ms_pipe / K,
            "what": "same work per step; prepare_ranks() of step i+1 runs on a second stream under "
                    "the pooling kernels of step i (it depends on the calibration only)"},
        "phases_ms": {k: round(v, 4) for k, v in avg.items()},
        "phases_gbs_algorithmic": {
            k: round(alg[a] / (avg[k] * 1e-3) / 1e9, 1)
            for k, a in (("prepare_v2", "prepare"), ("pool_fwd", "pool_fwd"), ("pool_bwd", "pool_bwd"))
            if k in avg},
        "clocks": clocks.summary(),
    }
    # ---- secondary: the open-vocabulary tail (tcgen05 3xTF32 logits + fused argmax), reported
    # as samples/s, algorithmic GB/s and fraction of the HBM roofline (it is HBM-bound: 9 flop/B)
    if not args.no_tail:
        from veon_b200.tail import class_of_prompt, voxel_text_argmax
        Bt, Ct, Qt = 2, 512, 18
        gt = torch.Generator(device=dev).manual_seed(7)
        feat_occ = torch.sigmoid(torch.randn(Bt, Ct, 16, 200, 200, device=dev, generator=gt)) - 0.5
        wt = torch.randn(Qt, Ct, device=dev, generator=gt)
        wt = 100.0 * wt / wt.norm(dim=1, keepdim=True)
        bin_occ = torch.randn(Bt, 2, 16, 200, 200, device=dev, generator=gt)
        cls_t = class_of_prompt(list(range(Qt - 1))).to(dev)
        for _ in range(3):
            voxel_text_argmax(feat_occ, wt, cls_t, bin_occ)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nt = 10
        t0.record()
        for _ in range(nt):
            voxel_text_argmax(feat_occ, wt, cls_t, bin_occ)
        t1.record()
        torch.cuda.synchronize()
        ms_t = t0.elapsed_time(t1) / nt
        bytes_t = Bt * (4 * 640000 * Ct + 8 * 640000 + 640000) + 4 * Qt * Ct
        line["tail"] = {"what": f"veon_voxel_text_argmax, C={Ct}, Q={Qt} prompt rows, {Bt} samples/call, "
                                "3xTF32 tcgen05 + fused class-max/argmax/gate -> uint8 [B,200,200,16]",
                        "samples_per_s_per_gpu": Bt / (ms_t * 1e-3), "ms_per_call": ms_t,
                        "achieved_gbs_algorithmic": bytes_t / (ms_t * 1e-3) / 1e9,
                        "frac_of_hbm_peak": bytes_t / (ms_t * 1e-3) / 1e9 / peak_gbs,
                        "tf32_tflops_issued": 3 * 2.0 * Bt * 640000 * Ct * 32 / (ms_t * 1e-3) / 1e12}
        if world == 1:   # the two 8f-4 legs are single-GPU figures (every rank would repeat them)
            # ---- secondary: the same tail from the decoder's resolution: the reference
            # up-samples feat_occ [B,C,8,100,100] to 16x200x200 and classifies there; ours classifies
            # the low-resolution volume (tcgen05) and interpolates the Q logit channels
            from veon_b200.tail import voxel_text_argmax_lowres
            Bl = 8
            feat_lr = torch.sigmoid(torch.randn(Bl, Ct, 8, 100, 100, device=dev, generator=gt)) - 0.5
            bin_lr = torch.randn(Bl, 2, 8, 100, 100, device=dev, generator=gt)
            ws_lr = torch.empty(Bl * Qt * 80000, dtype=torch.float32, device=dev)
            for _ in range(3):
                voxel_text_argmax_lowres(feat_lr, wt, cls_t, bin_lr, workspace=ws_lr)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(nt):
                voxel_text_argmax_lowres(feat_lr, wt, cls_t, bin_lr, workspace=ws_lr)
            t1.record()
            torch.cuda.synchronize()
            ms_l = t0.elapsed_time(t1) / nt
            bytes_l = Bl * (4 * 80000 * Ct + 8 * 80000 + 640000) + 4 * Qt * Ct
            line["tail_lowres"] = {
                "what": f"veon_voxel_text_argmax_lowres, C={Ct}, Q={Qt}, {Bl} samples/call, feat_occ "
                        "[B,C,8,100,100] -> logits (tcgen05 3xTF32) -> trilinear up-sampling of the Q "
                        "logits + class-max/argmax/gate -> uint8 [B,200,200,16]; equals the full-"
                        "resolution route on >= 99.99 % of voxels (tests/test_tail_gpu.py)",
                "samples_per_s_per_gpu": Bl / (ms_l * 1e-3), "ms_per_call": ms_l,
                "achieved_gbs_algorithmic": bytes_l / (ms_l * 1e-3) / 1e9,
                "frac_of_hbm_peak": bytes_l / (ms_l * 1e-3) / 1e9 / peak_gbs,
                "speedup_vs_full_resolution_tail": (Bl / ms_l) / (Bt / ms_t)}
            del feat_lr, bin_lr, ws_lr
            # ---- secondary: BASELINE configs[2] as one pipeline, lift + classify (C=512, Q=18) with
            # the classifier in front of the pooling (veon_b200.pipeline.lift_classify)
            from veon_b200.pipeline import lift_classify
            c3 = S.CONFIGS["C3"]
            Bp, Np, Dp = 8, c3.n_cams, c3.D
            Hp, Wp = c3.feat_hw
            neck3 = LSSViewTransformer(c3.grid_config, c3.input_size, c3.downsample, 8, Ct,
                                       collapse_z=False)
            cal3 = S.calibration(c3, batch=Bp)
            metas3 = [torch.from_numpy(cal3[k]).to(dev) for k in
                      ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")]
            depth3 = torch.softmax(torch.randn(Bp * Np, Dp, Hp, Wp, device=dev, generator=gt) * 4, 1)
            feat3 = torch.randn(Bp * Np, Ct, Hp, Wp, device=dev, generator=gt) * 0.05
            gate_w = torch.randn(2, Ct, device=dev, generator=gt)
            img3 = torch.zeros(Bp, Np, 1, Hp, Wp, device=dev)
            for _ in range(3):
                lift_classify(neck3, [img3] + metas3, depth3, feat3, wt, cls_t, gate_w)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(nt):
                lift_classify(neck3, [img3] + metas3, depth3, feat3, wt, cls_t, gate_w)
            t1.record()
            torch.cuda.synchronize()
            ms_p = t0.elapsed_time(t1) / nt
            line["lift_classify"] = {
                "what": f"C3 geometry (6 cams 32x88, D={Dp}), C={Ct} image features -> per-pixel logits "
                        f"(tcgen05) -> get_lidar_coor + prepare_v2 + bev_pool_v2 forward of Q+2={Qt + 2} "
                        "channels -> merge/argmax/gate -> uint8 [B,200,200,16]; equals pooling the C-channel "
                        "features and classifying the volume on >= 99.99 % of voxels (tests/test_tail_gpu.py)",
                "samples_per_call": Bp, "ms_per_call": ms_p,
                "samples_per_s_per_gpu": Bp / (ms_p * 1e-3)}
            del depth3, feat3, neck3
        del feat_occ, bin_occ
        # ---- secondary: the neck's 2x2x2 max-downsample of the pooled volume
        vol = out_grad.detach().clone().requires_grad_()
        go_ds = torch.randn(B, C, 8, 100, 100, device=dev, generator=gt)

        def ev_ms(fn, n=10):
            for _ in range(2):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        with torch.no_grad():
            ms_df = ev_ms(lambda: BP.MaxDown2x2x2.apply(vol))
        ms_dfb = ev_ms(lambda: torch.autograd.grad(BP.MaxDown2x2x2.apply(vol), vol, go_ds))
        vb = 4.0 * vol.numel()
        line["downsample"] = {
            "what": "2x2x2 max-downsample of the [B,C,16,200,200] volume "
                    "(view_transformer_raw.py:549-553) and ATen-exact gradient, own kernels",
            "fwd_ms": ms_df, "fwd_gbs": 1.125 * vb / (ms_df * 1e-3) / 1e9,
            "bwd_ms": ms_dfb - ms_df, "bwd_gbs": 2.25 * vb / ((ms_dfb - ms_df) * 1e-3) / 1e9,
            "frac_of_hbm_peak_fwd": 1.125 * vb / (ms_df * 1e-3) / 1e9 / peak_gbs}
        del vol, go_ds
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c1 = S.CONFIGS["C1"] if args.workload == "C2" else cfg
        v, cores, desc, _, _ = time_cpu_port(c1, args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": desc}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
