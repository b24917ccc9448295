"""CPU: the oracle against the golden vectors recorded from the reference
(tests/golden/make_golden.py) and against the reference's known-answer test."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import lift_oracle as O
from veon_b200 import synthetic as S

KEYS = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return np.load(os.path.join(golden_dir, "prepare_tiny.npz"))


@pytest.mark.parametrize("case", ["rig", "edge", "outside", "one_voxel"])
def test_prepare_oracle_matches_reference_fixture(tiny, case):
    out = O.prepare_v2(tiny[f"{case}.coor"], tiny["grid.lower"], tiny["grid.interval"],
                       tiny["grid.size"])
    if bool(tiny[f"{case}.none"]):
        assert out == (None,) * 5
        return
    for key, arr in zip(KEYS, out):
        assert arr.dtype == np.int32
        np.testing.assert_array_equal(arr, tiny[f"{case}.{key}"], err_msg=key)


def test_edge_fixture_really_has_the_edge_cases(tiny):
    """(-1,0) coordinates are KEPT (trunc, not floor) and collisions exist."""
    coor = tiny["edge.coor"].reshape(-1, 3)
    vox = (coor - tiny["grid.lower"]) / tiny["grid.interval"]
    neg = np.where((vox[:, 0] > -1) & (vox[:, 0] < 0) & np.isfinite(vox).all(1))[0]
    assert neg.size > 0
    kept_ids = set(tiny["edge.ranks_depth"].tolist())
    # the hand-placed (-0.5, 3.2, 1.1) point is kept
    assert any(int(i) in kept_ids for i in neg)
    assert tiny["edge.interval_lengths"].max() >= 100   # 300 collisions split over B=2


@pytest.mark.parametrize("name", ["C1_B1", "C1_B2", "small_B2", "C3_B1", "C4_B1", "C1_B27"])
def test_prepare_oracle_matches_reference_hashes(golden_dir, name):
    with open(os.path.join(golden_dir, "prepare_hashes.json")) as f:
        g = json.load(f)[name]
    cfg = S.CONFIGS[g["config"]]
    coor = S.lidar_coor_np(cfg, batch=g["batch"])
    assert sha(coor) == g["sha256"]["coor"], "synthetic coor is not bit-reproducible here"
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    out = O.prepare_v2(coor, lower, interval, size)
    assert out[0].size == g["n_kept"] and out[3].size == g["n_intervals"]
    for key, arr in zip(KEYS, out):
        assert sha(arr) == g["sha256"][key], key


def test_fp32_rank_collapse_is_reproduced(golden_dir):
    """B=27 > 26: float32 ranks merge neighbouring voxels (SURVEY 7 trap ii)."""
    with open(os.path.join(golden_dir, "prepare_hashes.json")) as f:
        g = json.load(f)
    assert g["C1_B27"]["n_intervals"] < 27 * (g["C1_B2"]["n_intervals"] // 2) 


def _kat(golden_dir):
    with open(os.path.join(golden_dir, "kat_bev_pool_v2.json")) as f:
        return json.load(f)


def test_pool_oracle_known_answer(golden_dir):
    k = _kat(golden_dir)
    depth = np.array(k["depth"], np.float32).reshape(k["depth_shape"])
    feat = np.ones(k["feat_shape"], np.float32)
    rd, rf, rb = (np.array(k[n], np.int32) for n in ("ranks_depth", "ranks_feat", "ranks_bev"))
    starts, lengths = np.array([0, 2], np.int32), np.array([2, 2], np.int32)
    out = O.bev_pool_v2(depth, feat, rd, rf, rb, tuple(k["bev_feat_shape"]), starts, lengths)
    assert out.shape == (1, 2, 1, 2, 2)
    assert abs(float(out.sum()) - k["loss"]) < 1e-6
    dg, fg = O.bev_pool_v2_backward(np.ones_like(out), depth, feat, rd, rf, rb)
    np.testing.assert_allclose(dg.reshape(-1), k["grad_depth"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(fg.reshape(-1), k["grad_feat"], rtol=0, atol=1e-6)


def test_pool_oracle_vs_float64_and_autograd():
    rng = np.random.RandomState(0)
    cfg = S.CONFIGS["tiny"]
    coor = S.lidar_coor_np(cfg)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    rb, rd, rf, st, ln = O.prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    C = 8
    depth = rng.rand(B, N, D, H, W).astype(np.float32)
    feat = rng.randn(B, N, H, W, C).astype(np.float32)
    shape = (B, 16, 200, 200, C)
    out = O.bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)
    ref = O.bev_pool_v2_f64(depth, feat, rd, rf, rb, shape)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)
    # backward against torch autograd of the scatter-add formulation
    import torch
    g = rng.randn(*out.shape).astype(np.float32)
    dg, fg = O.bev_pool_v2_backward(g, depth, feat, rd, rf, rb)
    td = torch.from_numpy(depth).double().requires_grad_()
    tf = torch.from_numpy(feat).double().requires_grad_()
    vol = torch.zeros(B * 16 * 200 * 200, C, dtype=torch.double).index_add_(
        0, torch.from_numpy(rb).long(),
        td.reshape(-1)[torch.from_numpy(rd).long()].unsqueeze(1)
        * tf.reshape(-1, C)[torch.from_numpy(rf).long()])
    (vol.view(B, 16, 200, 200, C).permute(0, 4, 1, 2, 3) * torch.from_numpy(g).double()).sum().backward()
    np.testing.assert_allclose(dg, td.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fg, tf.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_torch_cpu_lift_matches_c_oracle():
    """the timed CPU baseline computes the same thing as the checker"""
    import torch
    cfg = S.CONFIGS["tiny"]
    coor = S.lidar_coor_np(cfg)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    rb, rd, rf, st, ln = O.prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    C = 4
    g = torch.Generator().manual_seed(1)
    depth = torch.rand(B, N, D, H, W, generator=g)
    feat = torch.randn(B, N, C, H, W, generator=g)
    og = torch.randn(B, C, 16, 200, 200, generator=g)
    bev, dg, fg = O.torch_cpu_lift(torch.from_numpy(coor), depth, feat, lower, interval, size, og)
    feat_last = feat.permute(0, 1, 3, 4, 2).contiguous().numpy()
    out = O.bev_pool_v2(depth.numpy(), feat_last, rd, rf, rb, (B, 16, 200, 200, C), st, ln)
    np.testing.assert_allclose(bev.numpy(), out, rtol=1e-5, atol=1e-6)
    dgo, fgo = O.bev_pool_v2_backward(og.numpy(), depth.numpy(), feat_last, rd, rf, rb)
    np.testing.assert_allclose(dg.numpy(), dgo, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fg.permute(0, 1, 3, 4, 2).numpy(), fgo, rtol=1e-4, atol=1e-5)


def test_class_groups_and_tail_oracle():
    refl = [0] * 3 + [1] + [2] * 2 + [3]
    cls = O.class_groups(refl)
    assert cls.tolist() == [0, 0, 0, 1, 2, 2, 3, 4]       # background row = own group
    # the real nuScenes-brief group sizes: 66 prompts + bg = 67 rows, 18 classes
    sizes = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
    refl = [k for k, n in enumerate(sizes) for _ in range(n)]
    cls = O.class_groups(refl)
    assert cls.size == 67 and cls.max() == 17
    rng = np.random.RandomState(0)
    B, C, Z, Y, X = 1, 16, 2, 3, 5
    feat = rng.randn(B, C, Z, Y, X).astype(np.float32)
    w = rng.randn(67, C).astype(np.float32)
    bin_occ = rng.randn(B, 2, Z, Y, X).astype(np.float32)
    lab = O.voxel_text_labels(feat, w, cls, bin_occ)
    assert lab.shape == (B, X, Y, Z) and lab.dtype == np.uint8
    # brute force
    for z in range(Z):
        for y in range(Y):
            for x in range(X):
                logits = w @ feat[0, :, z, y, x]
                merged = [logits[cls == k].max() for k in range(18)]
                want = int(np.argmax(merged)) if bin_occ[0, 0, z, y, x] > bin_occ[0, 1, z, y, x] else 17
                assert lab[0, x, y, z] == want


# ---- tail at the decoder's resolution -------------------------------------
@pytest.mark.parametrize("shape,size", [
    ((2, 3, 8, 10, 12), (16, 20, 24)),      # the 2x case of the model
    ((1, 2, 3, 5, 7), (7, 9, 20)),          # non-integer scales
    ((1, 2, 6, 9, 9), (3, 4, 5)),           # down-sampling
    ((1, 2, 4, 4, 4), (4, 4, 4)),           # identity
    ((1, 1, 8, 100, 100), (16, 200, 200)),  # one channel at the real size
])
def test_trilinear_oracle_matches_torch_interpolate(shape, size):
    """The reference calls F.interpolate(..., mode="trilinear", align_corners=False)
    (san_in_veon_temporal.py:196-207); torch is that function, so it pins the restatement."""
    import torch
    import torch.nn.functional as F
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape)))
    want = F.interpolate(x, size=size, mode="trilinear", align_corners=False).numpy()
    got = O.trilinear_upsample(x.numpy(), size)
    assert got.shape == want.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    if shape[2:] == tuple(size):
        np.testing.assert_array_equal(got, x.numpy())


def test_lowres_tail_oracle_follows_the_reference_expressions():
    """interpolate -> einsum -> merge -> label rule, written with the reference's own torch
    expressions (san_in_veon_temporal.py:196-208,257-259; san_in_veon_entry_temporal.py:273-297;
    veon_temporal.py:223-229), against the numpy oracle."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    refl = [0, 0, 1, 2, 2, 2, 3]
    cls = O.class_groups(refl)
    B, C, size = 2, 24, (6, 10, 14)
    feat = torch.sigmoid(torch.randn(B, C, 3, 5, 7, generator=g)) - 0.5
    gate = torch.randn(B, 2, 3, 5, 7, generator=g)
    w = torch.randn(len(refl) + 1, C, generator=g)
    feat_occ = F.interpolate(feat, size=size, mode="trilinear", align_corners=False)
    bin_occ = F.interpolate(gate, size=size, mode="trilinear", align_corners=False)
    sem = torch.einsum("qc,bczhw->bqzhw", w, feat_occ)
    merged = torch.stack([sem[:, torch.from_numpy(np.where(cls == k)[0])].max(dim=1).values
                          for k in range(int(cls.max()) + 1)], dim=1)
    mx = torch.max(torch.softmax(merged, dim=1), dim=1)
    sel = (mx.values > 0.0) & (torch.softmax(bin_occ, dim=1)[:, 0] > 0.5)
    want = torch.where(sel, mx.indices, torch.ones_like(mx.indices) * 17)
    want = want.permute(0, 3, 2, 1).contiguous().numpy().astype(np.uint8)
    got = O.voxel_text_labels_lowres(feat.numpy(), w.numpy(), cls, gate.numpy(), size)
    assert got.shape == want.shape == (B, size[2], size[1], size[0])
    assert (got == want).mean() >= 0.999       # float near-ties only
    assert 0.2 < (got == 17).mean() < 0.8


def test_classifier_commutes_with_the_interpolation():
    """the identity the low-resolution route rests on, in float64"""
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((1, 6, 4, 5, 6))
    w = rng.standard_normal((3, 6))
    size = (8, 10, 12)
    a = np.einsum("qc,bczyx->bqzyx", w, O.trilinear_upsample(feat, size).astype(np.float64))
    b = O.trilinear_upsample(np.einsum("qc,bczyx->bqzyx", w, feat), size)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)


# ---- tail: golden vectors from the reference's own functions (tests/golden/make_golden_tail.py)
@pytest.fixture(scope="module")
def tail_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "tail_reference.npz"))


@pytest.mark.parametrize("voc", ["nuscenes_brief", "nuscenes_default"])
def test_tail_oracle_matches_reference_functions(tail_golden, voc):
    """class groups, logits, merged logits and labels against `_add_vocabulary_nuscenes`,
    `semantic_inference_3d`, `_merge_classes_prob` (run from /root/reference) and the label rule
    of veon_temporal.py:223-229,240 applied literally."""
    from veon_b200.tail import class_of_prompt
    t = tail_golden
    refl = t[f"{voc}.class_reflection"]
    cls = O.class_groups(refl)
    assert cls.size == refl.size + 1 == t[f"{voc}.w"].shape[0]
    assert class_of_prompt(refl.tolist()).numpy().tolist() == cls.tolist()   # the product's host logic
    merged = t[f"{voc}.merged"]
    assert int(cls.max()) + 1 == merged.shape[1] == 18
    assert cls[-1] == 17 and (cls[:-1] == refl).all()       # background row is its own, last class
    feat, w = t[f"{voc}.feat"], t[f"{voc}.w"]
    sem = np.einsum("qc,bczyx->bqzyx", w, feat)
    np.testing.assert_allclose(sem, t[f"{voc}.sem_occ"], rtol=0, atol=2e-5 * np.abs(sem).max())
    ref_sem = t[f"{voc}.sem_occ"]
    for k in range(18):                                      # the merge itself is exact
        np.testing.assert_array_equal(ref_sem[:, cls == k].max(axis=1), merged[:, k])
    got = O.voxel_text_labels(feat, w, cls, t[f"{voc}.bin_occ"])
    np.testing.assert_array_equal(got, t[f"{voc}.labels"])
    if voc == "nuscenes_brief":
        assert np.bincount(refl).tolist() == [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]


@pytest.mark.needs_reference
def test_tail_fixture_is_what_the_reference_functions_return_today(golden_dir, tmp_path):
    """build container only: re-run the generator (its own process: it stubs detectron2 & co. in
    sys.modules) and compare with the committed fixture"""
    import subprocess
    import sys
    out = tmp_path / "tail.npz"
    res = subprocess.run([sys.executable, os.path.join(golden_dir, "make_golden_tail.py"), str(out)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:]
    new, old = np.load(out), np.load(os.path.join(golden_dir, "tail_reference.npz"))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        np.testing.assert_array_equal(new[k], old[k], err_msg=k)
