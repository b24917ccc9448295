"""Does the tail slow down over time (power / thermal management)?  One configuration, timed
every 0.5 s for ~12 s, with the SM / memory clocks, power draw and throttle reasons beside it."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import _lib
if os.environ.get("VEON_LIB"): _lib.LIB_PATH = os.environ["VEON_LIB"]
from veon_b200.tail import class_of_prompt, voxel_text_argmax
dev = torch.device("cuda", 0)
C, B, V, Q = 512, 2, 640000, 18
feat = torch.rand(B, C, 16, 200, 200, device=dev) - 0.5
w = torch.randn(Q, C, device=dev); w = 100 * w / w.norm(dim=1, keepdim=True)
bin_occ = torch.randn(B, 2, 16, 200, 200, device=dev)
cls = class_of_prompt(list(range(17))).to(dev)
def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                          "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    return out
for _ in range(3): voxel_text_argmax(feat, w, cls, bin_occ)
torch.cuda.synchronize()
t_end = time.time() + float(os.environ.get("DRIFT_S", "12"))
while time.time() < t_end:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(400): voxel_text_argmax(feat, w, cls, bin_occ)
    e1.record()
    s = smi()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / 400 * 1e3:7.1f} us/call   {s}")
