"""Throughput of the fused voxel-text tail (veon_voxel_text_argmax) on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200 import _lib
if os.environ.get("VEON_LIB"):          # an experimental build of tools/build_variant.sh
    _lib.LIB_PATH = os.environ["VEON_LIB"]
from veon_b200.tail import class_of_prompt, voxel_text_argmax
SIZES = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
dev = torch.device("cuda", 0)
for C, refl, B in ((512, list(range(17)), 2), (512, [k for k, n in enumerate(SIZES) for _ in range(n)], 2),
                   (768, list(range(17)), 2)):
    Q = len(refl) + 1
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.sigmoid(torch.randn(B, C, 16, 200, 200, device=dev, generator=g)) - 0.5
    w = torch.randn(Q, C, device=dev, generator=g); w = 100 * w / w.norm(dim=1, keepdim=True)
    bin_occ = torch.randn(B, 2, 16, 200, 200, device=dev, generator=g)
    cls = class_of_prompt(refl).to(dev)
    for _ in range(3): voxel_text_argmax(feat, w, cls, bin_occ)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50; e0.record()
    for _ in range(n): lab = voxel_text_argmax(feat, w, cls, bin_occ)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    V = 640000
    byts = B * (4 * V * C + 8 * V + V) + 4 * Q * C
    flops = 2.0 * B * V * C * Q
    print(f"C={C} Q={Q} B={B}: {ms*1e3:8.1f} us/call  {B/ms*1e3:8.1f} samples/s  {byts/ms/1e6:7.1f} GB/s  {flops/ms/1e9:7.1f} TFLOP/s(fp32 useful)")
