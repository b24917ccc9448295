"""GPU parity: fused voxel-text logits + class merge + argmax + gate (tail)
against the CPU oracle.  north_star: labels agree on >= 99.99 % of voxels."""
import numpy as np
import pytest
import torch

from oracle import lift_oracle as O

pytestmark = pytest.mark.gpu
SIZES = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]   # nuscenes_brief prompts per class


def synth(B, C, Q_refl, Z, Y, X, seed):
    g = torch.Generator().manual_seed(seed)
    feat = torch.sigmoid(torch.randn(B, C, Z, Y, X, generator=g)) - 0.5
    w = torch.randn(len(Q_refl) + 1, C, generator=g)
    w = 100.0 * w / w.norm(dim=1, keepdim=True)
    bin_occ = torch.randn(B, 2, Z, Y, X, generator=g)
    return feat, w, bin_occ


def run(feat, w, refl, bin_occ):
    from veon_b200.tail import class_of_prompt, voxel_text_argmax
    cls = class_of_prompt(refl)
    lab = voxel_text_argmax(feat.cuda(), w.cuda(), cls.cuda(), bin_occ.cuda())
    torch.cuda.synchronize()
    return lab.cpu().numpy(), cls.numpy()


@pytest.mark.parametrize("C,refl,vol", [
    (512, list(range(17)), (4, 40, 50)),                                   # Q=18 (one prompt per class + bg)
    (512, [k for k, n in enumerate(SIZES) for _ in range(n)], (4, 40, 50)),  # Q=67, real group sizes
    (768, [k for k, n in enumerate(SIZES) for _ in range(n)], (2, 20, 37)),  # ViT-L dim, ragged X
    (30, [0, 0, 1, 2, 2, 2], (3, 5, 7)),                                   # odd C
])
def test_labels_match_oracle(C, refl, vol):
    Z, Y, X = vol
    feat, w, bin_occ = synth(2, C, refl, Z, Y, X, seed=C)
    got, cls = run(feat, w, refl, bin_occ)
    assert got.shape == (2, X, Y, Z) and got.dtype == np.uint8
    want = O.voxel_text_labels(feat.numpy(), w.numpy(), cls, bin_occ.numpy())
    agree = float((got == want).mean())
    assert agree >= 0.9999, agree
    # disagreements (if any) must be float near-ties, never gate errors
    want64 = O.voxel_text_labels(feat.numpy(), w.numpy(), cls, bin_occ.numpy(), dtype=np.float64)
    assert float((got == want64).mean()) >= 0.9999
    free = (bin_occ[:, 0] <= bin_occ[:, 1]).permute(0, 3, 2, 1).numpy()
    assert np.all(got[free] == 17)


def test_full_volume_one_sample():
    """one full Occ3D sample (16x200x200), Q=18, C=128 (smaller C keeps the CPU oracle quick)"""
    refl = list(range(17))
    feat, w, bin_occ = synth(1, 128, refl, 16, 200, 200, seed=3)
    got, cls = run(feat, w, refl, bin_occ)
    want = O.voxel_text_labels(feat.numpy(), w.numpy(), cls, bin_occ.numpy())
    assert float((got == want).mean()) >= 0.9999
    assert got.shape == (1, 200, 200, 16)


def test_ties_take_first_class_and_nan_is_free():
    refl = [0, 1, 2]
    Z, Y, X = 1, 2, 4
    C = 8
    feat = torch.zeros(1, C, Z, Y, X)            # all logits equal (0) -> class 0
    w = torch.randn(4, C)
    bin_occ = torch.zeros(1, 2, Z, Y, X)
    bin_occ[:, 0] = 1.0                          # occupied everywhere
    got, _ = run(feat, w, refl, bin_occ)
    assert np.all(got == 0)
    feat[0, 0, 0, 0, 0] = float("nan")
    got, _ = run(feat, w, refl, bin_occ)
    assert got[0, 0, 0, 0] == 17 and np.all(got.reshape(-1)[1:] == 0)
