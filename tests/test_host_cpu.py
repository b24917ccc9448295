"""CPU: C-ABI surface, host-side mirror logic, multi-process sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "veon_lift.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(veon_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from veon_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"libveonlift.so does not export {n}"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes table and header disagree"
    assert lib.veon_abi_version() == 1
    assert b"bad argument" in lib.veon_error_string(-1)


def test_argument_errors_do_not_need_a_gpu():
    from veon_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    assert lib.veon_bev_pool_v2(64, 10, null, null, null, null, null, null, null, null, null) == -1
    assert lib.veon_bev_pool_v2_fwd_planar(null, null, null, null, null, null, 1, 64, 640000, 4224, null, null) == -1
    gs = _lib.float3([200, 200, 16])
    assert lib.veon_prepare_v2_workspace_bytes(8, 6, 88, 16, 44, gs) > 0
    assert lib.veon_prepare_v2_workspace_bytes(0, 6, 88, 16, 44, gs) == 0
    assert lib.veon_pool_num_tiles(8, 640000) == 8 * 20000
    assert lib.veon_pool_num_tiles(2, 33) == 4


def test_cpu_tensors_are_refused_loudly():
    from veon_b200.bev_pool import bev_pool_v2, voxel_pooling_prepare_v2
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voxel_pooling_prepare_v2(torch.zeros(1, 1, 2, 2, 2, 3), [0, 0, 0], [1, 1, 1], [4, 4, 2])
    z = torch.zeros(1, 1, 2, 2, 2)
    i = torch.zeros(2, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bev_pool_v2(z, z, i, i, i, (1, 1, 2, 2, 2), i, i)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from veon_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VeonLibraryError):
        _lib.load()


def test_view_transformer_mirror_geometry(golden_dir):
    """frustum and get_lidar_coor against the reference's outputs"""
    from veon_b200 import synthetic as S
    from veon_b200.view_transformer import LSSViewTransformer
    geo = np.load(os.path.join(golden_dir, "geometry_tiny.npz"))
    for name in ("tiny", "C1"):
        cfg = S.CONFIGS[name]
        neck = LSSViewTransformer(cfg.grid_config, cfg.input_size, cfg.downsample, 8, cfg.channels)
        assert neck.D == int(geo[f"{name}.D"]) == cfg.D
        np.testing.assert_array_equal(neck.frustum.numpy(), geo[f"{name}.frustum"])
        lower, interval, size = S.grid_vectors(cfg.grid_config)
        np.testing.assert_array_equal(neck.grid_size.numpy(), size)
        np.testing.assert_array_equal(neck.grid_lower_bound.numpy(), lower)
        np.testing.assert_array_equal(neck.grid_interval.numpy(), interval)
        cal = S.calibration(cfg, batch=1)
        coor = neck.get_lidar_coor(*[torch.from_numpy(cal[k]) for k in
                                     ("sensor2ego", "ego2global", "intrins", "post_rots",
                                      "post_trans", "bda")]).numpy()
        if name == "tiny":
            np.testing.assert_allclose(coor, geo["tiny.coor"], rtol=0, atol=2e-5)
        else:
            np.testing.assert_allclose(coor[:, :, ::11, ::5, ::7], geo["C1.coor_sample"],
                                       rtol=0, atol=2e-4)
        # and the bit-reproducible rig generator describes the same geometry
        np.testing.assert_allclose(S.lidar_coor_np(cfg, batch=1), coor, rtol=0, atol=2e-4)


def test_tail_class_of_prompt_matches_oracle():
    from oracle.lift_oracle import class_groups
    from veon_b200.tail import class_of_prompt
    sizes = [16, 1, 1, 1, 8, 1, 1, 3, 1, 1, 1, 1, 5, 3, 5, 13, 4]
    refl = [k for k, n in enumerate(sizes) for _ in range(n)]
    assert class_of_prompt(refl).tolist() == class_groups(refl).tolist()
    assert class_of_prompt(list(range(17))).tolist() == list(range(18))


def test_shard_samples():
    from veon_b200.dist import shard_samples
    for n, g in ((8, 1), (8, 2), (9, 4), (3, 8), (64, 8)):
        seen = sorted(i for r in range(g) for i in shard_samples(n, g, r))
        assert seen == list(range(n))


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from veon_b200.dist import shard_samples, all_gather_occupancy
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
r = dist.get_rank()
for n in (5, 4, 1):
    mine = shard_samples(n, 2, r)
    local = torch.stack([torch.full((3, 4, 2), 10 * i + 1, dtype=torch.uint8) for i in mine]) if mine \
        else torch.zeros((0, 3, 4, 2), dtype=torch.uint8)
    full = all_gather_occupancy(local, n)
    assert full.shape == (n, 3, 4, 2), full.shape
    for i in range(n):
        assert int(full[i, 0, 0, 0]) == 10 * i + 1 and bool((full[i] == 10 * i + 1).all())
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_all_gather_occupancy_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
