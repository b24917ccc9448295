"""Where the roles of k_fwd_stream (CTA 0) spend their cycles on one forward; needs the traced build
of tools/fwd_trace.sh (VEON_LIB)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from veon_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.environ["VEON_LIB"]
from fwd_check import setup  # noqa: E402
from veon_b200 import bev_pool as BP  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C = int(sys.argv[3]) if len(sys.argv) > 3 else 64
lib = ctypes.CDLL(_lib.LIB_PATH)
prep, depth, feat, shape = setup(cfg, B, C)
def fwd():
    return BP._fwd_planar(depth, feat, prep.ranks_depth, prep.ranks_feat, prep.ranks_bev, prep.plan,
                          B, C, 640000, shape)
for _ in range(3):
    fwd()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 256)()
assert lib.veon_internal_fwd_trace(buf, 1) == 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fwd(); b.record()
torch.cuda.synchronize()
assert lib.veon_internal_fwd_trace(buf, 0) == 0
print(f"forward call {a.elapsed_time(b) * 1e3:.1f} us (traced build)")
mhz = 1965.0
names_a = ["prefetch wait", "header+issue", "slot wait", "rows", "-", "-", "-", "-"]
names_e = ["header+setup", "full wait", "copies", "group barrier", "write-out", "-", "-", "-"]
for role, lo, hi, names in (("A", 0, 16, names_a), ("E", 16, 32, names_e)):
    tot = [sum(buf[w * 8 + k] for w in range(lo, hi)) / (hi - lo) / mhz for k in range(8)]
    print(f"role {role} (mean over its warps, us): " + ", ".join(f"{n} {t:.1f}" for n, t in zip(names, tot) if n != "-"))
    per = [sum(buf[w * 8 + k] for k in range(8)) / mhz for w in range(lo, hi)]
    print(f"   per-warp total us: min {min(per):.1f} max {max(per):.1f}")
