"""Synthetic nuScenes-shaped inputs for the lifting path.

Everything here is *bit-reproducible across machines*: only IEEE +,-,*,/ on
explicitly typed numpy scalars/arrays, tabulated trigonometry and an integer
LCG -- no libm, no BLAS, no torch RNG.  That is what lets golden hashes of
the reference's `voxel_pooling_prepare_v2` outputs (made in the build
container, where /root/reference exists) be checked on the GPU box from a
regenerated `coor` (tests/golden/make_golden.py).

Geometry follows the reference:
  frustum        mmdet3d/models/necks/view_transformer.py:84-112
  lidar coords   mmdet3d/models/necks/view_transformer.py:114-152
  calibration    tests/test_models/test_necks/test_necks.py:139-158 (constants),
                 mmdet3d/datasets/pipelines/loading.py:1173-1186 (test-mode resize/crop)
  grid           configs/veon/veon-temporal-base-512x1408-zoe-nodepthcache.py:33-38
"""
from dataclasses import dataclass, field

import numpy as np

# (cos, sin) of the six nuScenes camera yaws [55, 0, -55, 110, 180, -110] deg,
# tabulated so no libm call is involved.
_YAW_COS_SIN = (
    (0.573576436351046, 0.819152044288992),
    (1.0, 0.0),
    (0.573576436351046, -0.819152044288992),
    (-0.342020143325669, 0.939692620785908),
    (-1.0, 0.0),
    (-0.342020143325669, -0.939692620785908),
)
_CAM_T = (
    (1.52, 0.49, 1.51),
    (1.70, 0.02, 1.51),
    (1.55, -0.49, 1.50),
    (1.04, 0.48, 1.56),
    (0.03, 0.00, 1.56),
    (1.01, -0.48, 1.56),
)
_FOCAL = (1266.4, 1266.4, 1266.4, 1266.4, 809.2, 1266.4)
_CX, _CY = 816.3, 491.5


@dataclass
class LiftConfig:
    """One lifting workload (BASELINE.json `configs`)."""
    name: str = "C2"
    n_cams: int = 6
    input_size: tuple = (256, 704)
    downsample: int = 16
    channels: int = 64
    batch: int = 8
    grid_config: dict = field(default_factory=lambda: {
        "x": [-40.0, 40.0, 0.4],
        "y": [-40.0, 40.0, 0.4],
        "z": [-1.0, 5.4, 0.4],
        "depth": [1.0, 45.0, 0.5],
    })

    @property
    def feat_hw(self):
        return (self.input_size[0] // self.downsample,
                self.input_size[1] // self.downsample)

    @property
    def D(self):
        lo, hi, st = self.grid_config["depth"]
        return int(np.arange(lo, hi, st, dtype=np.float32).shape[0])


def _depth(lo, hi, st):
    return {"x": [-40.0, 40.0, 0.4], "y": [-40.0, 40.0, 0.4],
            "z": [-1.0, 5.4, 0.4], "depth": [lo, hi, st]}


CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "C1": LiftConfig("C1", 6, (256, 704), 16, 64, 1),
    # configs[1]: the configuration the metric is quoted on
    "C2": LiftConfig("C2", 6, (256, 704), 16, 64, 8),
    # configs[2]: VEON ViT-B CLIP-dim lift
    "C3": LiftConfig("C3", 6, (512, 1408), 16, 512, 1),
    # configs[3]: VEON* ViT-L high-res
    "C4": LiftConfig("C4", 6, (512, 1408), 16, 768, 16, _depth(1.0, 60.0, 0.5)),
    # small shapes for tests / smoke
    "tiny": LiftConfig("tiny", 2, (64, 96), 16, 32, 2, _depth(1.0, 13.0, 1.0)),
    "small": LiftConfig("small", 6, (128, 352), 16, 64, 2, _depth(1.0, 45.0, 1.0)),
}


class _LCG:
    """64-bit LCG (Knuth MMIX constants); uniform() in [-1, 1)."""

    def __init__(self, seed):
        self.s = (int(seed) * 2862933555777941757 + 3037000493) & (2**64 - 1)

    def uniform(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & (2**64 - 1)
        return ((self.s >> 11) / float(2**53)) * 2.0 - 1.0


def grid_vectors(grid_config):
    """grid_lower_bound, grid_interval, grid_size as float32, computed the way
    `create_grid_infos` does (view_transformer.py:66-82): python-float
    arithmetic then a float32 cast."""
    axes = [grid_config[k] for k in ("x", "y", "z")]
    lower = np.array([a[0] for a in axes], dtype=np.float32)
    interval = np.array([a[2] for a in axes], dtype=np.float32)
    size = np.array([(a[1] - a[0]) / a[2] for a in axes], dtype=np.float32)
    return lower, interval, size


def frustum_np(cfg: LiftConfig):
    """[D, H, W, 3] float32 frustum (x_img, y_img, depth); linspace done in
    float64 with explicit arithmetic then cast (deterministic)."""
    H_in, W_in = cfg.input_size
    H, W = cfg.feat_hw
    lo, hi, st = cfg.grid_config["depth"]
    D = cfg.D
    d = (np.float64(lo) + np.arange(D, dtype=np.float64) * np.float64(st)).astype(np.float32)
    xs = (np.arange(W, dtype=np.float64) * (np.float64(W_in - 1) / max(W - 1, 1))).astype(np.float32)
    ys = (np.arange(H, dtype=np.float64) * (np.float64(H_in - 1) / max(H - 1, 1))).astype(np.float32)
    fr = np.empty((D, H, W, 3), dtype=np.float32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = d[:, None, None]
    return fr


def calibration(cfg: LiftConfig, batch=None, jitter=True, sample_offset=0):
    """Per-sample camera calibration.

    Returns dict of float32 arrays: sensor2ego [B,N,4,4], ego2global [B,N,4,4]
    (identity, unused by the path), intrins [B,N,3,3], post_rots [B,N,3,3],
    post_trans [B,N,3], bda [B,3,3].  Sample `i` is jittered (yaw +-2 deg,
    t +-5 cm) with seed = sample_offset + i.
    """
    B = cfg.batch if batch is None else batch
    N = cfg.n_cams
    H_in, W_in = cfg.input_size
    s = W_in / 1600.0
    crop_h = int(900 * s) - H_in
    s2e = np.zeros((B, N, 4, 4), dtype=np.float64)
    K = np.zeros((B, N, 3, 3), dtype=np.float64)
    for b in range(B):
        rng = _LCG(sample_offset + b)
        for n in range(N):
            c0, s0 = _YAW_COS_SIN[n % 6]
            t = list(_CAM_T[n % 6])
            if jitter:
                dth = rng.uniform() * (2.0 * 3.141592653589793 / 180.0)
                # polynomial small-angle rotation: deterministic, no libm
                d2 = dth * dth
                cd = 1.0 - d2 / 2.0 + d2 * d2 / 24.0
                sd = dth - d2 * dth / 6.0
                c, sn = c0 * cd - s0 * sd, s0 * cd + c0 * sd
                t = [t[k] + 0.05 * rng.uniform() for k in range(3)]
            else:
                c, sn = c0, s0
            R = np.array([[sn, 0.0, c], [-c, 0.0, sn], [0.0, -1.0, 0.0]])
            s2e[b, n, :3, :3] = R
            s2e[b, n, :3, 3] = t
            s2e[b, n, 3, 3] = 1.0
            f = _FOCAL[n % 6]
            K[b, n] = [[f, 0.0, _CX], [0.0, f, _CY], [0.0, 0.0, 1.0]]
    post_rots = np.zeros((B, N, 3, 3), dtype=np.float32)
    post_rots[..., 0, 0] = s
    post_rots[..., 1, 1] = s
    post_rots[..., 2, 2] = 1.0
    post_trans = np.zeros((B, N, 3), dtype=np.float32)
    post_trans[..., 1] = -float(crop_h)
    eye4 = np.broadcast_to(np.eye(4, dtype=np.float32), (B, N, 4, 4)).copy()
    bda = np.broadcast_to(np.eye(3, dtype=np.float32), (B, 3, 3)).copy()
    return {
        "sensor2ego": s2e.astype(np.float32), "ego2global": eye4,
        "intrins": K.astype(np.float32), "post_rots": post_rots,
        "post_trans": post_trans, "bda": bda,
    }


def lidar_coor_np(cfg: LiftConfig, calib=None, batch=None, sample_offset=0):
    """[B,N,D,H,W,3] float32 ego-frame frustum points with a FIXED float32
    operation order (elementwise only), so the bits are machine-independent.
    Same geometry as `get_lidar_coor` (view_transformer.py:114-152) for the
    rig's diagonal post_rots / upper-triangular K / identity bda."""
    if calib is None:
        calib = calibration(cfg, batch=batch, sample_offset=sample_offset)
    fr = frustum_np(cfg)
    s2e = calib["sensor2ego"].astype(np.float64)
    K = calib["intrins"].astype(np.float64)
    B, N = s2e.shape[:2]
    D, H, W, _ = fr.shape
    out = np.empty((B, N, D, H, W, 3), dtype=np.float32)
    f32 = np.float32
    for b in range(B):
        for n in range(N):
            sx = f32(calib["post_rots"][b, n, 0, 0])
            sy = f32(calib["post_rots"][b, n, 1, 1])
            tx, ty, _ = (f32(v) for v in calib["post_trans"][b, n])
            px = (fr[..., 0] - tx) / sx
            py = (fr[..., 1] - ty) / sy
            pz = fr[..., 2]
            xc, yc = px * pz, py * pz
            f_, cx, cy = K[b, n, 0, 0], K[b, n, 0, 2], K[b, n, 1, 2]
            Kinv = np.array([[1.0 / f_, 0.0, -cx / f_], [0.0, 1.0 / f_, -cy / f_], [0.0, 0.0, 1.0]])
            M = np.empty((3, 3), dtype=np.float64)
            for i in range(3):          # explicit 3x3 product, fixed order
                for j in range(3):
                    M[i, j] = (s2e[b, n, i, 0] * Kinv[0, j] + s2e[b, n, i, 1] * Kinv[1, j]) \
                        + s2e[b, n, i, 2] * Kinv[2, j]
            M = M.astype(np.float32)
            t = s2e[b, n, :3, 3].astype(np.float32)
            for i in range(3):
                out[b, n, ..., i] = ((M[i, 0] * xc + M[i, 1] * yc) + M[i, 2] * pz) + t[i]
    return out


def metric_depth_np(cfg: LiftConfig, batch=None, seed=0):
    """Deterministic pseudo metric depth in [1, 45) per feature pixel
    [B,N,H,W] float32 (input of the two-hot depth producer, §8f-2)."""
    B = cfg.batch if batch is None else batch
    H, W = cfg.feat_hw
    n = B * cfg.n_cams * H * W
    idx = np.arange(n, dtype=np.uint64) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & (2**64 - 1))
    x = idx * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
    x ^= x >> np.uint64(29)
    x = x * np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(32)
    u = (x >> np.uint64(40)).astype(np.float64) / float(2**24)
    return (1.0 + 44.0 * u).astype(np.float32).reshape(B, cfg.n_cams, H, W)
