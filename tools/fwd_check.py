"""Forward pooling: streaming two-role kernel vs the general kernel (bit-exact) + timings.
    python tools/fwd_check.py [check] [time]
Run under `timeout`: the streaming kernel spins on flags, a bug there hangs instead of failing."""
import sys
import time

import torch

import os
sys.path.insert(0, ".")
from veon_b200 import _lib, bev_pool as BP, synthetic as S  # noqa: E402
if os.environ.get("VEON_LIB"):          # an experimental build of tools/build_variant.sh
    _lib.LIB_PATH = os.environ["VEON_LIB"]


def setup(cfg_name, B, C, seed=0):
    cfg = S.CONFIGS[cfg_name]
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B)).cuda()
    _, N, D, H, W, _ = coor.shape
    g = torch.Generator(device="cuda").manual_seed(seed)
    depth = torch.softmax(torch.randn(B, N, D, H, W, device="cuda", generator=g) * 4, dim=2)
    feat = torch.randn(B, N, H, W, C, device="cuda", generator=g)
    prep = BP.prepare_ranks(coor, lower, interval, size)
    return prep, depth, feat, (B, C, 16, 200, 200)


def general(prep, depth, feat, shape, out=None):
    lib = _lib.load()
    B, C = shape[0], shape[1]
    V = 640000
    if out is None:
        out = torch.full(shape, float("nan"), device="cuda")
    p = prep.plan
    rc = lib.veon_bev_pool_v2_fwd_planar(
        BP._ptr(depth), BP._ptr(feat), BP._ptr(prep.ranks_depth), BP._ptr(prep.ranks_feat),
        BP._ptr(prep.ranks_bev), BP._ptr(p.tile_start), None, None, BP._ptr(p.tile_heavy),
        p.tile_heavy.numel(), B, C, V, feat.numel() // C, BP._ptr(out), None, 0,
        BP._stream_ptr(depth.device))
    assert rc == 0, rc
    return out


def ev_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    what = sys.argv[1:] or ["check", "time"]
    import ctypes
    ctypes.CDLL(_lib.LIB_PATH).veon_internal_fwd_stream_force(1)   # wide rows too
    if "check" in what:
        for cfg, B, C in (("C1", 1, 64), ("C1", 2, 64), ("small", 3, 128), ("C1", 1, 256),
                          ("C3", 1, 512), ("C1", 1, 768), ("C2", 8, 64)):
            prep, depth, feat, shape = setup(cfg, B, C)
            t0 = time.time()
            a = BP._fwd_planar(depth, feat, prep.ranks_depth, prep.ranks_feat, prep.ranks_bev,
                               prep.plan, B, C, 640000, shape)
            torch.cuda.synchronize()
            b = general(prep, depth, feat, shape)
            torch.cuda.synchronize()
            same = torch.equal(a, b)
            print(f"check {cfg} B={B} C={C}: equal={same} ({time.time() - t0:.2f}s)", flush=True)
            if not same:
                d = (a != b) | (a.isnan() != b.isnan())
                idx = d.nonzero()
                print("  mismatches:", int(d.sum()), "first:", idx[:5].tolist())
            # twice more: the ring / control block are reused
            for _ in range(2):
                a2 = BP._fwd_planar(depth, feat, prep.ranks_depth, prep.ranks_feat,
                                    prep.ranks_bev, prep.plan, B, C, 640000, shape)
                assert torch.equal(a2, b) == same
            del a, b, prep, depth, feat
    if "time" in what:
        for cfg, B, C in (("C2", 8, 64), ("C3", 2, 512), ("C3", 4, 256), ("C3", 8, 128)):
            prep, depth, feat, shape = setup(cfg, B, C)
            scratch = torch.empty(shape, device="cuda")
            tg = ev_ms(lambda: general(prep, depth, feat, shape, scratch), n=10)
            del scratch
            line = f"time {cfg} B={B} C={C}: general {tg * 1e3:.1f} us"
            slot = 512 * min(C, 128) * 4
            for ns in (4, 8, 16) if C <= 512 else (8, 16):
                BP.FWD_RING_BYTES = 148 * ns * slot
                ts = ev_ms(lambda: BP._fwd_planar(depth, feat, prep.ranks_depth, prep.ranks_feat,
                                                  prep.ranks_bev, prep.plan, B, C, 640000, shape), n=10)
                line += f" | {ns} slots/SM: {ts * 1e3:.1f} us"
            BP.FWD_RING_BYTES = None
            gb = 4.0 * 640000 * C * B / 1e9
            line += f"   (volume {gb:.2f} GB)"
            print(line, flush=True)
            del prep, depth, feat


if __name__ == "__main__":
    main()
