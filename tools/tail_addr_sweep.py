"""Does the tail's speed depend on where the feature volume lies?  Same call, the volume placed
at different offsets inside one big allocation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import _lib
if os.environ.get("VEON_LIB"): _lib.LIB_PATH = os.environ["VEON_LIB"]
from veon_b200.tail import class_of_prompt, voxel_text_argmax
dev = torch.device("cuda", 0)
C, B, V = 512, 2, 640000
Q = 18
pool = torch.empty(B * C * V + (64 << 20), device=dev)
w = torch.randn(Q, C, device=dev); w = 100 * w / w.norm(dim=1, keepdim=True)
bin_occ = torch.randn(B, 2, 16, 200, 200, device=dev)
cls = class_of_prompt(list(range(17))).to(dev)
print("pool base %x" % pool.data_ptr())
for off_kb in (0, 4, 64, 256, 1024, 2048, 4096, 16384, 65536, 131072):
    off = off_kb * 256
    feat = pool[off:off + B * C * V].view(B, C, 16, 200, 200)
    feat.uniform_(-0.5, 0.5)
    for _ in range(3): voxel_text_argmax(feat, w, cls, bin_occ)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): voxel_text_argmax(feat, w, cls, bin_occ)
    e1.record(); torch.cuda.synchronize()
    print(f"offset {off_kb:7d} KB: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us/call")
