"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own `view_transformer.py` (unmodified, stubbed
third-party imports -- tests/golden/_ref_loader.py) and records what
`LSSViewTransformer.voxel_pooling_prepare_v2` / `get_lidar_coor` /
`create_frustum` return on seeded synthetic inputs:

  prepare_tiny.npz      full input + output arrays of small cases, including
                        the edge cases (coords in (-1,0), on cell borders,
                        NaN/inf/huge, everything outside, heavy collisions)
  prepare_hashes.json   sha256 of the outputs at full BASELINE sizes; `coor` is
                        regenerated bit-identically from veon_b200.synthetic
                        (its sha256 is stored too, so a mismatch is detected)
  geometry_tiny.npz     frustum + get_lidar_coor of the reference for the rig
  kat_bev_pool_v2.json  the known-answer test written in
                        mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176

The reference's argsort is unstable, so ranks_depth / ranks_feat are stored in
the canonical in-voxel order (ascending ranks_depth); ranks_bev and the
intervals are stored exactly as returned.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from _ref_loader import load_reference_view_transformer  # noqa: E402
from oracle.lift_oracle import canonical_order  # noqa: E402
from veon_b200 import synthetic as S  # noqa: E402

warnings.simplefilter("ignore")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_neck(mod, cfg):
    return mod.LSSViewTransformer(grid_config=cfg.grid_config, input_size=cfg.input_size,
                                  downsample=cfg.downsample, in_channels=8,
                                  out_channels=cfg.channels, collapse_z=False)


def run_prepare(neck, coor):
    out = neck.voxel_pooling_prepare_v2(torch.from_numpy(coor))
    if out[0] is None:
        return None
    rb, rd, rf, st, ln = (t.numpy() for t in out)
    rb_c, rd_c, rf_c = canonical_order(rb, rd, rf)
    assert np.array_equal(rb_c, rb)          # reference output is sorted by ranks_bev
    return rb, rd_c, rf_c, st, ln


def edge_case_coor(cfg, seed):
    """tiny-config coor with hand-placed edge values (in voxel units)."""
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor = S.lidar_coor_np(cfg, batch=cfg.batch, sample_offset=seed)
    flat = coor.reshape(-1, 3)
    rng = np.random.RandomState(seed)
    n = flat.shape[0]

    def put(i, vx, vy, vz):
        flat[i] = np.array([vx, vy, vz], dtype=np.float32) * interval + lower

    specials = [(-0.5, 3.2, 1.1), (-0.999, -0.001, -0.5), (0.0, 0.0, 0.0),
                (199.999, 199.5, 15.9), (200.0, 5.0, 5.0), (5.0, 200.0, 5.0),
                (5.0, 5.0, 16.0), (-1.0, 5.0, 5.0), (-1.0001, 5.0, 5.0),
                (7.0, 7.0, 7.0), (7.999, 7.001, 7.5), (7.0, 7.0, 7.0)]
    idx = rng.choice(n, size=len(specials) + 6, replace=False)
    for i, sp in zip(idx, specials):
        put(i, *sp)
    k = len(specials)
    flat[idx[k + 0]] = [np.nan, 0.0, 0.0]
    flat[idx[k + 1]] = [0.0, np.inf, 0.0]
    flat[idx[k + 2]] = [0.0, 0.0, -np.inf]
    flat[idx[k + 3]] = [1e30, 0.0, 0.0]
    flat[idx[k + 4]] = [-1e30, 0.0, 0.0]
    flat[idx[k + 5]] = [3e9, 3e9, 3e9]
    # heavy collisions: 300 points into one voxel, 40 into another
    coll = rng.choice(n, size=340, replace=False)
    for j, i in enumerate(coll):
        base = (11.0, 12.0, 3.0) if j < 300 else (150.0, 60.0, 9.0)
        put(i, base[0] + rng.rand() * 0.9, base[1] + rng.rand() * 0.9, base[2] + rng.rand() * 0.9)
    return coor


def main():
    mod = load_reference_view_transformer(lambda *a, **k: None)

    # ---------------------------------------------------------------- tiny
    tiny = {}
    cfg = S.CONFIGS["tiny"]
    neck = ref_neck(mod, cfg)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    assert np.array_equal(neck.grid_lower_bound.numpy(), lower)
    assert np.array_equal(neck.grid_interval.numpy(), interval)
    assert np.array_equal(neck.grid_size.numpy(), size)
    cases = {
        "rig": S.lidar_coor_np(cfg, batch=cfg.batch),
        "edge": edge_case_coor(cfg, 3),
        "outside": S.lidar_coor_np(cfg, batch=1) + np.float32(500.0),
        "one_voxel": np.full((1, 1, 3, 2, 5, 3), 2.0, dtype=np.float32),
    }
    for name, coor in cases.items():
        res = run_prepare(neck, coor)
        tiny[f"{name}.coor"] = coor
        tiny[f"{name}.none"] = np.array(res is None)
        if res is not None:
            for key, arr in zip(("ranks_bev", "ranks_depth", "ranks_feat",
                                 "interval_starts", "interval_lengths"), res):
                tiny[f"{name}.{key}"] = arr
        print(f"tiny/{name}: points={coor[..., 0].size} ->",
              "None" if res is None else f"kept={res[0].size} intervals={res[3].size}")
    tiny["grid.lower"], tiny["grid.interval"], tiny["grid.size"] = lower, interval, size
    np.savez_compressed(os.path.join(HERE, "prepare_tiny.npz"), **tiny)

    # ------------------------------------------------------------ geometry
    geo = {}
    for name in ("tiny", "C1"):
        cfg = S.CONFIGS[name]
        neck = ref_neck(mod, cfg)
        cal = S.calibration(cfg, batch=1)
        args = [torch.from_numpy(cal[k]) for k in
                ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")]
        coor = neck.get_lidar_coor(*args).numpy()
        geo[f"{name}.frustum"] = neck.frustum.numpy()
        if name == "tiny":
            geo[f"{name}.coor"] = coor
        else:   # keep the fixture small: a strided sample + checksum
            geo[f"{name}.coor_sample"] = coor[:, :, ::11, ::5, ::7].copy()
        geo[f"{name}.D"] = np.array(neck.D)
    np.savez_compressed(os.path.join(HERE, "geometry_tiny.npz"), **geo)

    # -------------------------------------------------------- full-size hashes
    hashes = {}
    for name, batch in (("C1", 1), ("C1", 2), ("small", 2), ("C3", 1), ("C4", 1), ("C1", 27)):
        cfg = S.CONFIGS[name]
        neck = ref_neck(mod, cfg)
        coor = S.lidar_coor_np(cfg, batch=batch)
        rb, rd, rf, st, ln = run_prepare(neck, coor)
        hashes[f"{name}_B{batch}"] = {
            "config": name, "batch": batch, "points": int(coor[..., 0].size),
            "n_kept": int(rb.size), "n_intervals": int(st.size),
            "max_interval": int(ln.max()),
            "sha256": {"coor": sha(coor), "ranks_bev": sha(rb), "ranks_depth": sha(rd),
                       "ranks_feat": sha(rf), "interval_starts": sha(st),
                       "interval_lengths": sha(ln)},
        }
        print(name, batch, hashes[f"{name}_B{batch}"]["n_kept"],
              hashes[f"{name}_B{batch}"]["n_intervals"])
    with open(os.path.join(HERE, "prepare_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)

    # ------------------------------------------------------------------- KAT
    kat = {
        "source": "mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176 (test_bev_pool_v2)",
        "depth": [0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], "depth_shape": [1, 1, 2, 2, 2],
        "feat": "ones", "feat_shape": [1, 1, 2, 2, 2],
        "ranks_depth": [0, 4, 1, 6], "ranks_feat": [0, 0, 1, 2], "ranks_bev": [0, 0, 1, 1],
        "bev_feat_shape": [1, 1, 2, 2, 2],
        "loss": 4.4,
        "grad_depth": [2.0, 2.0, 0.0, 0.0, 2.0, 0.0, 2.0, 0.0],
        "grad_feat": [1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0.0, 0.0],
    }
    with open(os.path.join(HERE, "kat_bev_pool_v2.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()
