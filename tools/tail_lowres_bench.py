"""Decoder-resolution tail on one GPU: the two launches timed separately and the
reference's order of operations (F.interpolate x2 -> our full-resolution tail, and -> torch einsum
+ merge + label rule) beside them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from veon_b200 import _lib
if os.environ.get("VEON_LIB"): _lib.LIB_PATH = os.environ["VEON_LIB"]
from veon_b200.tail import (class_of_prompt, semantic_inference_3d, upsample_classify,
                            voxel_text_argmax, voxel_text_argmax_lowres)
SIZES = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
dev = torch.device("cuda", 0)
size = (16, 200, 200)


def ev_ms(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for C, refl, B in ((512, list(range(17)), 8), (512, [k for k, n in enumerate(SIZES) for _ in range(n)], 8),
                   (768, list(range(17)), 8), (512, list(range(17)), 1)):
    Q = len(refl) + 1
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.sigmoid(torch.randn(B, C, 8, 100, 100, device=dev, generator=g)) - 0.5
    w = torch.randn(Q, C, device=dev, generator=g); w = 100 * w / w.norm(dim=1, keepdim=True)
    gate = torch.randn(B, 2, 8, 100, 100, device=dev, generator=g)
    cls = class_of_prompt(refl).to(dev)
    ws = torch.empty(B * Q * 80000, dtype=torch.float32, device=dev)
    t_all = ev_ms(lambda: voxel_text_argmax_lowres(feat, w, cls, gate, size, workspace=ws))
    t_log = ev_ms(lambda: semantic_inference_3d(w, feat))
    sem = semantic_inference_3d(w, feat)
    t_up = ev_ms(lambda: upsample_classify(sem, gate, cls, size))
    line = (f"C={C} Q={Q} B={B}: lowres route {t_all*1e3:7.1f} us ({B/t_all*1e3:8.0f} samples/s; "
            f"logits {t_log*1e3:6.1f} us = {4.0*B*80000*C/t_log/1e6:6.0f} GB/s, "
            f"upsample+classify {t_up*1e3:6.1f} us)")
    if B <= 2 or C <= 512:
        Bs = min(B, 2)

        def ref_order():
            f = F.interpolate(feat[:Bs], size=size, mode="trilinear", align_corners=False)
            b = F.interpolate(gate[:Bs], size=size, mode="trilinear", align_corners=False)
            return voxel_text_argmax(f, w, cls, b)
        t_ref = ev_ms(ref_order, 5)
        line += f"; interpolate + full-resolution tail {t_ref/Bs*1e3:7.1f} us/sample"
    print(line, flush=True)
