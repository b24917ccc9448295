import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from veon_b200.tail import class_of_prompt, voxel_text_argmax
SIZES = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
Q67 = len(sys.argv) > 1 and sys.argv[1] == "q67"
dev = torch.device("cuda", 0); C = 512; B = 1
refl = [k for k, n in enumerate(SIZES) for _ in range(n)] if Q67 else list(range(17))
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.sigmoid(torch.randn(B, C, 16, 200, 200, device=dev, generator=g)) - 0.5
w = torch.randn(len(refl) + 1, C, device=dev, generator=g); w = 100 * w / w.norm(dim=1, keepdim=True)
bo = torch.randn(B, 2, 16, 200, 200, device=dev, generator=g); cls = class_of_prompt(refl).to(dev)
for _ in range(3): lab = voxel_text_argmax(feat, w, cls, bo)
torch.cuda.synchronize(); print("ok", int(lab.sum()))
