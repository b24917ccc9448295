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


# ---- semantic_inference_3d alone and the decoder-resolution route ---------
@pytest.mark.parametrize("C,Q,vol", [
    (512, 18, (4, 40, 50)),     # tcgen05 path, one W piece
    (256, 67, (8, 25, 20)),     # tcgen05 path, two W pieces
    (30, 6, (3, 5, 7)),         # FFMA path (C % 32 != 0, V % 4 != 0)
])
def test_semantic_inference_3d_matches_fp32_einsum(C, Q, vol):
    """logits of san_in_veon_temporal.py:257-259; tolerance: max-abs error <= 1e-5 of the
    largest logit (north_star allows 1e-3 relative for fp32 accumulation; 3xTF32 gives ~2^-21)."""
    from veon_b200.tail import semantic_inference_3d
    g = torch.Generator().manual_seed(C + Q)
    feat = torch.sigmoid(torch.randn(2, C, *vol, generator=g)) - 0.5
    w = torch.randn(Q, C, generator=g)
    w = 100.0 * w / w.norm(dim=1, keepdim=True)
    got = semantic_inference_3d(w.cuda(), feat.cuda())
    torch.cuda.synchronize()
    assert got.shape == (2, Q, *vol) and got.dtype == torch.float32
    want = torch.einsum("qc,bczhw->bqzhw", w.double(), feat.double())
    err = (got.cpu().double() - want).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), err


def lowres_case(B, C, refl, lr, seed):
    g = torch.Generator().manual_seed(seed)
    feat = torch.sigmoid(torch.randn(B, C, *lr, generator=g)) - 0.5
    w = torch.randn(len(refl) + 1, C, generator=g)
    w = 100.0 * w / w.norm(dim=1, keepdim=True)
    gate = torch.randn(B, 2, *lr, generator=g)
    return feat, w, gate


@pytest.mark.parametrize("C,refl,lr,size", [
    (512, list(range(17)), (8, 20, 25), (16, 40, 50)),                                   # column kernel, Q=18
    (256, [k for k, n in enumerate(SIZES) for _ in range(n)], (8, 12, 16), (16, 24, 32)),  # Q=67
    (64, list(range(17)), (8, 10, 10), (16, 37, 23)),         # column kernel, non-integer x/y scales
    (30, [0, 0, 1, 2, 2, 2], (3, 5, 7), (7, 9, 20)),          # generic kernel + FFMA logits
    (32, [0, 1, 1, 2], (4, 6, 8), (4, 6, 8)),                 # identity sizes
    (32, [0, 1, 1, 2], (6, 8, 12), (3, 4, 5)),                # down-sampling
])
def test_lowres_route_matches_oracle(C, refl, lr, size):
    """classify-then-interpolate (ours) against interpolate-then-classify (the reference's order,
    CPU oracle): labels agree on >= 99.99 % of voxels."""
    from veon_b200.tail import class_of_prompt, voxel_text_argmax_lowres
    feat, w, gate = lowres_case(2, C, refl, lr, seed=C + len(refl))
    cls = class_of_prompt(refl)
    got = voxel_text_argmax_lowres(feat.cuda(), w.cuda(), cls.cuda(), gate.cuda(), occ_size=size)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    Z, Y, X = size
    assert got.shape == (2, X, Y, Z) and got.dtype == np.uint8
    want = O.voxel_text_labels_lowres(feat.numpy(), w.numpy(), cls.numpy(), gate.numpy(), size)
    agree = float((got == want).mean())
    n = want.size
    assert agree >= 0.9999 or (n < 20000 and (got != want).sum() <= 1), agree
    gate_hr = O.trilinear_upsample(gate.numpy(), size)
    clear = np.abs(gate_hr[:, 0] - gate_hr[:, 1]) > 1e-4
    free = np.transpose((gate_hr[:, 0] <= gate_hr[:, 1]) & clear, (0, 3, 2, 1))
    assert np.all(got[free] == 17)
    assert 0.2 < (got != 17).mean() < 0.8


def test_lowres_route_full_volume_and_full_resolution_route_agree():
    """8x100x100 -> 16x200x200 (the model's sizes), C=128: against the oracle and against our own
    full-resolution kernel fed with torch's up-sampled volume."""
    import torch.nn.functional as F
    from veon_b200.tail import class_of_prompt, voxel_text_argmax, voxel_text_argmax_lowres
    refl = list(range(17))
    size = (16, 200, 200)
    feat, w, gate = lowres_case(1, 128, refl, (8, 100, 100), seed=11)
    cls = class_of_prompt(refl)
    ws = torch.empty(18 * 80000, dtype=torch.float32, device="cuda")
    got = voxel_text_argmax_lowres(feat.cuda(), w.cuda(), cls.cuda(), gate.cuda(), occ_size=size,
                                   workspace=ws)
    feat_hr = F.interpolate(feat.cuda(), size=size, mode="trilinear", align_corners=False)
    gate_hr = F.interpolate(gate.cuda(), size=size, mode="trilinear", align_corners=False)
    full = voxel_text_argmax(feat_hr, w.cuda(), cls.cuda(), gate_hr)
    torch.cuda.synchronize()
    assert got.shape == full.shape == (1, 200, 200, 16)
    assert float((got == full).float().mean()) >= 0.9999
    want = O.voxel_text_labels_lowres(feat.numpy(), w.numpy(), cls.numpy(), gate.numpy(), size)
    assert float((got.cpu().numpy() == want).mean()) >= 0.9999


def test_upsample_classify_on_given_logits_and_nan():
    """exactly representable weights (2x up-sampling: 0.25 / 0.75) on integer logits: the label
    map must equal the oracle's everywhere; a NaN logit frees every output it touches."""
    from veon_b200.tail import upsample_classify
    g = torch.Generator().manual_seed(2)
    refl = [0, 1, 1, 2]
    cls = torch.from_numpy(O.class_groups(refl))
    for lr, size in (((8, 6, 9), (16, 12, 18)), ((2, 3, 4), (4, 6, 8))):
        sem = torch.randint(-8, 9, (1, 5, *lr), generator=g).float() * 4.0 + torch.arange(5).view(1, 5, 1, 1, 1) * 0.125
        gate = torch.randint(-3, 4, (1, 2, *lr), generator=g).float() * 4.0 + torch.tensor([0.0, 1.0]).view(1, 2, 1, 1, 1)
        got = upsample_classify(sem.cuda(), gate.cuda(), cls.cuda(), size).cpu().numpy()
        sem_hr = O.trilinear_upsample(sem.numpy(), size)
        gate_hr = O.trilinear_upsample(gate.numpy(), size)
        merged = np.stack([sem_hr[:, np.where(cls.numpy() == k)[0]].max(axis=1) for k in range(4)], 1)
        want = np.where(gate_hr[:, 0] > gate_hr[:, 1], merged.argmax(1), 17)
        np.testing.assert_array_equal(got, np.transpose(want, (0, 3, 2, 1)).astype(np.uint8))
        sem[0, 2, 0, 0, 0] = float("nan")
        got = upsample_classify(sem.cuda(), gate.cuda(), cls.cuda(), size).cpu().numpy()
        touched = np.isnan(O.trilinear_upsample(sem.numpy(), size)).any(axis=1)
        assert touched.sum() >= 1
        assert np.all(got[np.transpose(touched, (0, 3, 2, 1))] == 17)


def test_classify_logits_on_channel_slices():
    """merge + label rule on ready-made logits, given as slices of one volume (no copy), against
    the oracle's rule; column kernel (Z=16) and the generic one"""
    from veon_b200.tail import classify_logits
    refl = [0, 0, 1, 2, 2, 2, 3]
    cls = torch.from_numpy(O.class_groups(refl))
    Q = len(refl) + 1
    for vol in ((16, 9, 13), (5, 4, 7)):
        g = torch.Generator().manual_seed(vol[0])
        both = torch.randn(3, Q + 4, *vol, generator=g).cuda()
        got = classify_logits(both[:, :Q], both[:, Q:Q + 2], cls.cuda()).cpu().numpy()
        sem, gate = both[:, :Q].cpu().numpy(), both[:, Q:Q + 2].cpu().numpy()
        merged = np.stack([sem[:, np.where(cls.numpy() == k)[0]].max(axis=1) for k in range(5)], 1)
        want = np.where(gate[:, 0] > gate[:, 1], merged.argmax(1), 17)
        np.testing.assert_array_equal(got, np.transpose(want, (0, 3, 2, 1)).astype(np.uint8))


def test_lift_classify_in_logit_space_matches_feature_space_and_oracle():
    """pooling the per-pixel logits (Q+2 channels) == classifying the pooled C-channel volume:
    against our own feature-space route and against the CPU oracle (lift + tail)."""
    from veon_b200 import synthetic as S
    from veon_b200.pipeline import lift_classify, lift_then_classify
    from veon_b200.tail import class_of_prompt
    from veon_b200.view_transformer import LSSViewTransformer
    cfg = S.CONFIGS["small"]
    B, C = 2, 64
    refl = list(range(17))
    H, W = cfg.feat_hw
    g = torch.Generator().manual_seed(4)
    cal = S.calibration(cfg, batch=B)
    metas = [torch.from_numpy(cal[k]).cuda() for k in
             ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")]
    depth = torch.softmax(torch.randn(B * cfg.n_cams, cfg.D, H, W, generator=g) * 4, dim=1).cuda()
    feat = (torch.randn(B * cfg.n_cams, C, H, W, generator=g) * 0.05).cuda()
    w = torch.randn(len(refl) + 1, C, generator=g)
    w = (100.0 * w / w.norm(dim=1, keepdim=True)).cuda()
    gate_w = torch.randn(2, C, generator=g).cuda()
    cls = class_of_prompt(refl).cuda()
    img = torch.zeros(B, cfg.n_cams, 8, H, W, device="cuda")
    neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, collapse_z=False)
    got = lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w)
    ref = lift_then_classify(neck, [img] + metas, depth, feat, w, cls, gate_w)
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (B, 200, 200, 16) and got.dtype == torch.uint8
    assert float((got == ref).float().mean()) >= 0.9999
    # CPU oracle: the reference's lift (index_add_ order), then the tail's oracle.  The pooled
    # fp32 features themselves depend on the summation order, so next to the agreement rate
    # every disagreement must be a near-tie of the float64 evaluation (top-2 class margin or
    # gate margin below 1e-4 of the voxel's largest logit / gate value, or an absolute gate
    # difference below 1e-6: `softmax(bin_occ)[:,0] > 0.5` in fp32 cannot resolve less than
    # ~1.2e-7 and what it returns there depends on the exp implementation), never a gross error.
    coor = neck.get_lidar_coor(*metas).cpu()
    d5, f5 = depth.view(B, cfg.n_cams, cfg.D, H, W).cpu(), feat.view(B, cfg.n_cams, C, H, W).cpu()
    grid = (neck.grid_lower_bound, neck.grid_interval, neck.grid_size)
    vol, _, _ = O.torch_cpu_lift(coor, d5, f5, *grid)
    gate = torch.einsum("kc,bczyx->bkzyx", gate_w.cpu(), vol)
    want = O.voxel_text_labels(vol.numpy(), w.cpu().numpy(), cls.cpu().numpy(), gate.numpy())
    got_np = got.cpu().numpy()
    assert float((got_np == want).mean()) >= 0.9995
    vol64, _, _ = O.torch_cpu_lift(coor, d5.double(), f5.double(), *grid)
    sem64 = torch.einsum("qc,bczyx->bqzyx", w.cpu().double(), vol64)     # one prompt per class
    gate64 = torch.einsum("kc,bczyx->bkzyx", gate_w.cpu().double(), vol64)
    top2 = sem64.topk(2, dim=1).values
    cls_margin = (top2[:, 0] - top2[:, 1]) / sem64.abs().amax(dim=1).clamp_min(1e-300)
    gate_margin = (gate64[:, 0] - gate64[:, 1]).abs() / gate64.abs().amax(dim=1).clamp_min(1e-300)
    gate_abs = (gate64[:, 0] - gate64[:, 1]).abs()
    near_tie = ((cls_margin < 1e-4) | (gate_margin < 1e-4) | (gate_abs < 1e-6)) \
        .permute(0, 3, 2, 1).numpy()
    want64 = O.voxel_text_labels(vol64.numpy(), w.cpu().numpy(), cls.cpu().numpy(), gate64.numpy(),
                                 dtype=np.float64)
    assert np.all(near_tie[got_np != want64]), int((~near_tie[got_np != want64]).sum())
    assert 0.001 < float((got != 17).float().mean()) < 0.9


@pytest.mark.parametrize("cfg_name,Q", [("small", 18), ("small", 67), ("small", 5), ("small", 30),
                                        ("small", 40), ("small", 62), ("small", 94), ("C3", 18),
                                        ("C3", 40), ("C3", 67)])
def test_fused_lift_classify_equals_pool_then_classify(cfg_name, Q):
    """veon_lift_classify_fwd (labels straight from the pooling kernel's registers, the logit
    volume never written) against pooling the volume and running classify_logits on it: the same
    labels voxel for voxel -- the sums are the same bits and the rule is the same.  C3 density
    exercises the CTA-per-tile kernel (tiles of 512+ points); Q = 30 / 40, 62 / 67, 94 the widest
    one-pass row and the two- and three-pass forms (rows of 32 / 48, 64 / 72, 96 channels)."""
    from veon_b200 import synthetic as S
    from veon_b200.pipeline import lift_classify
    from veon_b200.tail import class_of_prompt
    from veon_b200.view_transformer import LSSViewTransformer
    cfg = S.CONFIGS[cfg_name]
    B, C = 2, 64
    refl = ([k for k, n in enumerate(SIZES) for _ in range(n)] if Q == 67 else list(range(Q - 1)))
    H, W = cfg.feat_hw
    g = torch.Generator().manual_seed(Q)
    cal = S.calibration(cfg, batch=B)
    metas = [torch.from_numpy(cal[k]).cuda() for k in
             ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")]
    depth = torch.softmax(torch.randn(B * cfg.n_cams, cfg.D, H, W, generator=g) * 4, dim=1).cuda()
    feat = (torch.randn(B * cfg.n_cams, C, H, W, generator=g) * 0.05).cuda()
    w = torch.randn(Q, C, generator=g)
    w = (100.0 * w / w.norm(dim=1, keepdim=True)).cuda()
    gate_w = torch.randn(2, C, generator=g).cuda()
    cls = class_of_prompt(refl).cuda()
    img = torch.zeros(B, cfg.n_cams, 8, H, W, device="cuda")
    neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, collapse_z=False)
    from veon_b200 import bev_pool as BP
    BP.enable_kernel_timing(True)
    fused = lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w)
    torch.cuda.synchronize()
    assert "lift_classify_fwd" in BP.kernel_timings_ms()       # the fused kernel did run
    BP.enable_kernel_timing(False)
    plain = lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w, fused=False)
    torch.cuda.synchronize()
    assert fused.shape == plain.shape == (B, 200, 200, 16) and fused.dtype == torch.uint8
    assert torch.equal(fused, plain), int((fused != plain).sum())
    assert 0.001 < float((fused != 17).float().mean()) < 0.9


@pytest.mark.parametrize("voc", ["nuscenes_brief", "nuscenes_default"])
def test_tail_kernels_match_reference_golden(golden_dir, voc):
    """the reference's own `semantic_inference_3d` / `_merge_classes_prob` outputs and the label
    rule applied literally (tests/golden/make_golden_tail.py), real prompt groups"""
    import os
    from veon_b200.tail import (class_of_prompt, classify_logits, semantic_inference_3d,
                                voxel_text_argmax)
    t = np.load(os.path.join(golden_dir, "tail_reference.npz"))
    cls = class_of_prompt(t[f"{voc}.class_reflection"].tolist()).cuda()
    feat, w = torch.from_numpy(t[f"{voc}.feat"]).cuda(), torch.from_numpy(t[f"{voc}.w"]).cuda()
    bin_occ = torch.from_numpy(t[f"{voc}.bin_occ"]).cuda()
    want = t[f"{voc}.labels"]
    got = voxel_text_argmax(feat, w, cls, bin_occ).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    sem = semantic_inference_3d(w, feat)
    ref_sem = t[f"{voc}.sem_occ"]
    assert np.abs(sem.cpu().numpy() - ref_sem).max() <= 1e-5 * np.abs(ref_sem).max()
    # merge + label rule on the reference's own logits: nothing left to round
    got2 = classify_logits(torch.from_numpy(ref_sem).cuda(), bin_occ, cls).cpu().numpy()
    np.testing.assert_array_equal(got2, want)


# ---- adversarial near-ties and the full-size C=512 case -------------------
@pytest.mark.parametrize("eps", [1e-6, 1e-5, 1e-4, 1e-3])
@pytest.mark.parametrize("Q_per_class", [1, 3])
def test_near_tie_labels(eps, Q_per_class):
    """Classifier rows that differ from each other by a relative perturbation `eps`, so that the
    two best classes of EVERY voxel are a near-tie with a logit gap of about eps x |logit|
    (the 3xTF32 contraction has ~2^-21 relative error per product; plain TF32 would flip
    classes up to eps ~ 1e-3).  Rule checked: wherever the float64 gap between the best and the
    second-best merged class exceeds 2e-5 of the largest logit (twice the logit tolerance of
    test_semantic_inference_3d_matches_fp32_einsum) the label is the float64 arg-max; below that
    any class within the tolerance of the best one is acceptable."""
    C, Z, Y, X, B = 512, 4, 40, 50, 2
    n_cls = 6
    g = torch.Generator().manual_seed(int(-np.log10(eps)) * 10 + Q_per_class)
    base = torch.randn(1, C, generator=g)
    rows = base + eps * base.norm() * torch.nn.functional.normalize(
        torch.randn(n_cls * Q_per_class + 1, C, generator=g), dim=1)
    w = (100.0 * rows / rows.norm(dim=1, keepdim=True)).float()
    refl = [k for k in range(n_cls) for _ in range(Q_per_class)]
    feat = torch.sigmoid(torch.randn(B, C, Z, Y, X, generator=g)) - 0.5
    bin_occ = torch.zeros(B, 2, Z, Y, X)
    bin_occ[:, 0] = 1.0                                    # everything occupied: classes decide
    got, cls = run(feat, w, refl, bin_occ)
    sem = torch.einsum("qc,bczyx->bqzyx", w.double(), feat.double())
    merged = torch.stack([sem[:, torch.from_numpy(cls == k)].max(dim=1).values
                          for k in range(int(cls.max()) + 1)], 1)          # [B,K,Z,Y,X]
    top = merged.max(dim=1)
    scale = sem.abs().amax()
    got_t = torch.from_numpy(got.astype(np.int64)).permute(0, 3, 2, 1)     # [B,Z,Y,X]
    assert int(got_t.max()) < merged.shape[1]                              # never "free" here
    chosen = merged.gather(1, got_t.unsqueeze(1)).squeeze(1)
    # the chosen class is the float64 best one up to the logit tolerance -- which pins it
    # exactly wherever the runner-up is further away than that
    assert float(((top.values - chosen) / scale).max()) <= 2e-5
    second = merged.topk(2, dim=1).values[:, 1]
    clear = (top.values - second) / scale > 2e-5
    assert bool((got_t[clear] == top.indices[clear]).all())
    if eps >= 1e-4:
        assert float(clear.float().mean()) > 0.1        # the construction does decide many voxels


def test_full_volume_c512_q18_and_q67():
    """BASELINE configs[2] at full size: one Occ3D sample (16x200x200), C=512, the 18-row and the
    real 67-row vocabulary; labels against the fp32 CPU oracle (>= 99.99 %) and, where they
    differ, against float64 (near-ties only)."""
    for refl in (list(range(17)), [k for k, n in enumerate(SIZES) for _ in range(n)]):
        feat, w, bin_occ = synth(1, 512, refl, 16, 200, 200, seed=len(refl))
        got, cls = run(feat, w, refl, bin_occ)
        want = O.voxel_text_labels(feat.numpy(), w.numpy(), cls, bin_occ.numpy())
        assert got.shape == (1, 200, 200, 16)
        assert float((got == want).mean()) >= 0.9999
        want64 = O.voxel_text_labels(feat.numpy(), w.numpy(), cls, bin_occ.numpy(), dtype=np.float64)
        assert float((got == want64).mean()) >= 0.9999


# ---- training-time arg-max over point lists, camera-group sharding ----------
def _loss_merge(tensor, class_reflection):
    """loss/occ_loss_utils/occ3d_nuscenes.py:249-265 (_merge_classes_prob, dim=1), literally"""
    dim_length = tensor.shape[1]
    assert dim_length == len(class_reflection)
    merged, left = [], 0
    while left < dim_length:
        right = left
        while right < dim_length - 1 and class_reflection[left] == class_reflection[right + 1]:
            right += 1
        merged.append(tensor[:, left:right + 1].max(dim=1, keepdim=True).values)
        left = right + 1
    return torch.cat(merged, dim=1)


@pytest.mark.parametrize("N,C,refl", [
    (50000, 512, [k for k, n in enumerate(SIZES) for _ in range(n)]),   # real 66 prompts, ViT-B dim
    (12345, 768, list(range(17))),                                        # N % 4 != 0, ViT-L dim
    (777, 30, [0, 0, 1, 2, 2, 2]),                                        # FFMA logits path
    (0, 64, [0, 1]),
])
def test_point_text_argmax_matches_the_loss_expressions(N, C, refl):
    """occ3d_nuscenes.py:472-482: einsum('nc,dc->nd', feat, W[:-1]) -> max(dim=1).indices and the
    merged-class arg-max.  Checked against the same torch expressions in float64: equal wherever
    the decision margin exceeds the logit tolerance, >= 99.99 % overall."""
    from veon_b200.tail import point_text_argmax
    g = torch.Generator().manual_seed(N + C)
    feat = torch.sigmoid(torch.randn(N, C, generator=g)) - 0.5
    w = torch.randn(len(refl) + 1, C, generator=g)
    w = 100.0 * w / w.norm(dim=1, keepdim=True)
    p_idx, c_idx = point_text_argmax(feat.cuda(), w.cuda(), refl)
    torch.cuda.synchronize()
    assert p_idx.shape == c_idx.shape == (N,) and p_idx.dtype == c_idx.dtype == torch.int64
    if N == 0:
        return
    probs = torch.einsum("nc,dc->nd", feat.double(), w[:-1].double())
    want_p = probs.max(dim=1).indices
    merged = _loss_merge(probs, refl)
    want_c = merged.max(dim=1).indices
    p_idx, c_idx = p_idx.cpu(), c_idx.cpu()
    assert float((p_idx == want_p).float().mean()) >= 0.9999
    assert float((c_idx == want_c).float().mean()) >= 0.9999
    scale = probs.abs().max()
    chosen_p = probs.gather(1, p_idx[:, None]).squeeze(1)
    chosen_c = merged.gather(1, c_idx[:, None]).squeeze(1)
    assert float(((probs.max(1).values - chosen_p) / scale).max()) <= 2e-5
    assert float(((merged.max(1).values - chosen_c) / scale).max()) <= 2e-5


def test_camera_group_sharding_sums_to_the_unsharded_labels():
    """SURVEY 8e, B < G: the logit volumes of disjoint camera groups add up to the full one
    (pooling is linear), so all-reducing them and classifying gives the unsharded labels up to
    float re-association: >= 99.99 % of voxels.  (Two 'ranks' emulated in one process; the
    all-reduce itself runs in tests/test_host_cpu.py over gloo and in bench.py over NCCL.)"""
    from veon_b200 import synthetic as S
    from veon_b200.dist import shard_cameras
    from veon_b200.pipeline import lift_classify, lift_logits
    from veon_b200.tail import class_of_prompt, classify_logits
    from veon_b200.view_transformer import LSSViewTransformer
    cfg = S.CONFIGS["small"]
    B, C = 1, 64
    refl = list(range(17))
    Q = len(refl) + 1
    H, W = cfg.feat_hw
    g = torch.Generator().manual_seed(14)
    cal = S.calibration(cfg, batch=B)
    metas = [torch.from_numpy(cal[k]).cuda() for k in
             ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")]
    depth = torch.softmax(torch.randn(B * cfg.n_cams, cfg.D, H, W, generator=g) * 4, dim=1).cuda()
    feat = (torch.randn(B * cfg.n_cams, C, H, W, generator=g) * 0.05).cuda()
    w = torch.randn(Q, C, generator=g)
    w = (100.0 * w / w.norm(dim=1, keepdim=True)).cuda()
    gate_w = torch.randn(2, C, generator=g).cuda()
    cls = class_of_prompt(refl).cuda()
    img = torch.zeros(B, cfg.n_cams, 8, H, W, device="cuda")
    neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C, collapse_z=False)
    want = lift_classify(neck, [img] + metas, depth, feat, w, cls, gate_w)
    for world in (2, 3):
        vol = None
        for r in range(world):
            part = lift_logits(neck, [img] + metas, depth, feat, w, gate_w,
                               cameras=shard_cameras(cfg.n_cams, world, r))
            vol = part if vol is None else vol + part
        got = classify_logits(vol[:, :Q], vol[:, Q:Q + 2], cls, 17)
        torch.cuda.synchronize()
        assert got.shape == want.shape
        assert float((got == want).float().mean()) >= 0.9999


def test_prepared_classifier_image_and_call_time_image_agree():
    """`w_image` of the C ABI: a classifier image prepared once with veon_text_classifier_image and
    NULL (the library builds it for the call in a stream-ordered allocation) give the same labels
    and the same logits, also after the stream has been synchronised in between (the allocation
    pool may have been trimmed)."""
    import ctypes
    from veon_b200 import _lib
    from veon_b200.tail import class_of_prompt
    lib = _lib.load()
    refl = [k for k, n in enumerate(SIZES) for _ in range(n)]
    B, C, Z, Y, X = 2, 256, 4, 30, 52
    feat, w, bin_occ = synth(B, C, refl, Z, Y, X, seed=11)
    feat, w, bin_occ = feat.cuda(), w.cuda(), bin_occ.cuda()
    cls = class_of_prompt(refl).cuda()
    Q = w.shape[0]
    P = ctypes.c_void_p
    st = P(torch.cuda.current_stream().cuda_stream)
    need = lib.veon_text_classifier_image_bytes(Q, C)
    assert need == (C // 32) * 2 * 80 * 32 * 4
    image = torch.empty(need, dtype=torch.uint8, device="cuda")
    assert lib.veon_text_classifier_image(P(w.data_ptr()), Q, C, P(image.data_ptr()), need - 16, st) == -2
    assert lib.veon_text_classifier_image(P(w.data_ptr()), Q, C, P(image.data_ptr()), need, st) == 0
    labels = [torch.empty((B, X, Y, Z), dtype=torch.uint8, device="cuda") for _ in range(3)]
    logits = [torch.empty((B, Q, Z, Y, X), device="cuda") for _ in range(2)]
    for out, img in ((labels[0], P(image.data_ptr())), (labels[1], None)):
        assert lib.veon_voxel_text_argmax(P(feat.data_ptr()), P(w.data_ptr()), P(cls.data_ptr()),
                                          P(bin_occ.data_ptr()), B, C, Q, Z, Y, X, 17,
                                          P(out.data_ptr()), img, st) == 0
    torch.cuda.synchronize()
    assert lib.veon_voxel_text_argmax(P(feat.data_ptr()), P(w.data_ptr()), P(cls.data_ptr()),
                                      P(bin_occ.data_ptr()), B, C, Q, Z, Y, X, 17,
                                      P(labels[2].data_ptr()), None, st) == 0
    for out, img in ((logits[0], P(image.data_ptr())), (logits[1], None)):
        assert lib.veon_semantic_inference_3d(P(w.data_ptr()), P(feat.data_ptr()), B, C, Q, Z, Y, X,
                                              P(out.data_ptr()), img, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(labels[0], labels[1]) and torch.equal(labels[0], labels[2])
    assert torch.equal(logits[0], logits[1])
    want = torch.einsum("qc,bczyx->bqzyx", w.double(), feat.double())
    assert float((logits[0].double() - want).abs().max() / want.abs().max()) < 1e-5
