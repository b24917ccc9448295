"""Time the index preparation (veon_prepare_v2) alone: python tools/prepare_bench.py [cfg B]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import hashlib
import torch
from veon_b200 import _lib
if os.environ.get("VEON_LIB"):
    _lib.LIB_PATH = os.environ["VEON_LIB"]
from veon_b200 import bev_pool as BP, synthetic as S
for cfg_name, B in ((sys.argv[1], int(sys.argv[2])),) if len(sys.argv) > 2 else (("C2", 8), ("C3", 8)):
    cfg = S.CONFIGS[cfg_name]
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B)).cuda()
    for _ in range(3):
        prep = BP.prepare_ranks(coor, lower, interval, size)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        prep = BP.prepare_ranks(coor, lower, interval, size)
    b.record()
    torch.cuda.synchronize()
    n = prep.plan.n_points
    h = hashlib.sha256()
    for t in (prep.ranks_bev[:n], prep.ranks_depth[:n], prep.ranks_feat[:n],
              prep.interval_starts[:prep.plan.n_intervals], prep.plan.tile_start, prep.plan.tile_occ):
        h.update(t.cpu().numpy().tobytes())
    print(f"{cfg_name} B={B}: prepare {a.elapsed_time(b) / 50 * 1e3:.1f} us  kept {n}  sha {h.hexdigest()[:16]}")
