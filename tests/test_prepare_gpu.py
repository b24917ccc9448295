"""GPU parity: CUDA voxel_pooling_prepare_v2 (through the C ABI) against the
oracle, the reference fixtures and the full-size reference hashes.  Integer
work => bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import lift_oracle as O
from veon_b200 import synthetic as S

pytestmark = pytest.mark.gpu
KEYS = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gpu_prepare(coor, lower, interval, size):
    from veon_b200.bev_pool import voxel_pooling_prepare_v2
    out = voxel_pooling_prepare_v2(torch.from_numpy(np.ascontiguousarray(coor)).cuda(),
                                   lower, interval, size)
    torch.cuda.synchronize()
    if out[0] is None:
        return out
    for t in out:
        assert t.dtype == torch.int32 and t.is_contiguous() and t.is_cuda
    return tuple(t.cpu().numpy() for t in out)


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return np.load(os.path.join(golden_dir, "prepare_tiny.npz"))


@pytest.mark.parametrize("case", ["rig", "edge", "outside", "one_voxel"])
def test_prepare_matches_reference_fixture(tiny, case):
    out = gpu_prepare(tiny[f"{case}.coor"], tiny["grid.lower"], tiny["grid.interval"],
                      tiny["grid.size"])
    if bool(tiny[f"{case}.none"]):
        assert out == (None,) * 5
        return
    for key, arr in zip(KEYS, out):
        np.testing.assert_array_equal(arr, tiny[f"{case}.{key}"], err_msg=key)


def test_prepare_empty_input_returns_nones(tiny):
    coor = np.zeros((0, 2, 3, 4, 5, 3), np.float32)
    from veon_b200.bev_pool import voxel_pooling_prepare_v2
    out = voxel_pooling_prepare_v2(torch.from_numpy(coor).cuda(), tiny["grid.lower"],
                                   tiny["grid.interval"], tiny["grid.size"])
    assert out == (None,) * 5


@pytest.mark.parametrize("name", ["C1_B1", "C1_B2", "small_B2", "C3_B1", "C4_B1", "C1_B27"])
def test_prepare_matches_reference_hashes_full_size(golden_dir, name):
    with open(os.path.join(golden_dir, "prepare_hashes.json")) as f:
        g = json.load(f)[name]
    cfg = S.CONFIGS[g["config"]]
    coor = S.lidar_coor_np(cfg, batch=g["batch"])
    assert sha(coor) == g["sha256"]["coor"], "synthetic coor is not bit-reproducible here"
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    out = gpu_prepare(coor, lower, interval, size)
    assert out[0].size == g["n_kept"] and out[3].size == g["n_intervals"]
    for key, arr in zip(KEYS, out):
        assert sha(arr) == g["sha256"][key], key


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_prepare_random_vs_oracle(seed):
    """random clouds incl. out-of-grid points, odd shapes, non-square grid"""
    rng = np.random.RandomState(seed)
    B, N, D, H, W = [(2, 3, 7, 5, 9), (1, 1, 1, 1, 1), (3, 2, 40, 6, 11)][seed]
    lower = np.array([-10.0, -6.0, -2.0], np.float32)
    interval = np.array([0.5, 0.25, 1.0], np.float32)
    size = np.array([40.0, 48.0, 4.0], np.float32) if seed != 1 else np.array([3.0, 5.0, 2.0], np.float32)
    coor = (rng.rand(B, N, D, H, W, 3).astype(np.float32) - 0.5) * np.array([26, 16, 7], np.float32)
    want = O.prepare_v2(coor, lower, interval, size)
    got = gpu_prepare(coor, lower, interval, size)
    if want[0] is None:
        assert got == (None,) * 5
        return
    for key, a, b in zip(KEYS, got, want):
        np.testing.assert_array_equal(a, b, err_msg=key)


def test_prepare_pathological_collisions():
    """every point in ONE voxel (> the in-kernel rank-by-counting limit):
    exercises the bitonic long-segment path; order must still be ascending."""
    B, N, D, H, W = 1, 2, 25, 10, 12   # 6000 points
    lower = np.array([0, 0, 0], np.float32)
    interval = np.array([1, 1, 1], np.float32)
    size = np.array([4, 4, 2], np.float32)
    coor = np.full((B, N, D, H, W, 3), 1.5, np.float32)
    rb, rd, rf, st, ln = gpu_prepare(coor, lower, interval, size)
    P = B * N * D * H * W
    assert st.tolist() == [0] and ln.tolist() == [P]
    np.testing.assert_array_equal(rd, np.arange(P, dtype=np.int32))
    want = O.prepare_v2(coor, lower, interval, size)
    np.testing.assert_array_equal(rf, want[2])
    np.testing.assert_array_equal(rb, want[0])


def test_prepare_properties_full_size():
    """size-independent invariants at C2 (B=8): sortedness, intervals tile the
    points exactly, ranks_depth unique and ascending inside every interval,
    ranks_feat consistent with ranks_depth."""
    cfg = S.CONFIGS["C2"]
    coor = S.lidar_coor_np(cfg)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    rb, rd, rf, st, ln = gpu_prepare(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    assert np.all(np.diff(rb) >= 0)
    assert st[0] == 0 and np.array_equal(st[1:], np.cumsum(ln)[:-1]) and st[-1] + ln[-1] == rb.size
    heads = np.flatnonzero(np.r_[True, rb[1:] != rb[:-1]])
    np.testing.assert_array_equal(heads, st)
    inside = np.ones(rd.size, bool); inside[st] = False
    assert np.all(np.diff(rd)[inside[1:]] > 0)
    assert np.unique(rd).size == rd.size
    HW = H * W
    np.testing.assert_array_equal(rf, (rd // (D * HW)) * HW + rd % HW)
    assert rb.min() >= 0 and rb.max() < B * 640000


def test_prepare_is_deterministic():
    cfg = S.CONFIGS["C1"]
    coor = S.lidar_coor_np(cfg, batch=2)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    a = gpu_prepare(coor, lower, interval, size)
    b = gpu_prepare(coor, lower, interval, size)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
