"""Golden vectors for the depth-distribution producer, made by the
reference's OWN `downsample_depth` / `get_two_hot_depth`
(mmdet3d/models/necks/view_transformer_raw.py:393-429), executed unmodified from
/root/reference with the stub loader.  Run in the build container only:

    python tests/golden/make_golden_depth.py      ->  tests/golden/two_hot_depth.npz
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from _ref_loader import REF_ROOT, load_reference_view_transformer  # noqa: E402


def load_raw_module():
    load_reference_view_transformer(lambda *a, **k: None)      # installs the stubs
    name = "mmdet3d.models.necks.view_transformer_raw"
    path = os.path.join(REF_ROOT, "mmdet3d/models/necks/view_transformer_raw.py")
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    module.__package__ = "mmdet3d.models.necks"
    sys.modules[name] = module
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(module)
    return module


def metric_depth(B, N, H, W, seed):
    """synthetic metric depth maps: smooth ramps + noise, holes (0 = no measurement),
    values below the first and beyond the last bin"""
    g = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    d = 2.0 + 40.0 * yy[None, None] * (0.5 + 0.5 * np.sin(6.0 * xx[None, None] + g.rand(B, N, 1, 1) * 6))
    d = d + g.randn(B, N, H, W) * 0.3
    d[g.rand(B, N, H, W) < 0.15] = 0.0
    d[0, 0, :4, :4] = 0.0            # an all-zero 4x4 / part of an all-zero 8x8 block
    d[0, 0, 8:16, 8:16] = 0.0        # a whole 8x8 block without a measurement
    d[0, 0, 20, 20] = 0.3            # closer than the first bin
    d[0, 0, 30, 30] = 80.0           # beyond the last bin
    return d.astype(np.float32)


def main():
    mod = load_raw_module()
    Raw = mod.LSSViewTransformerRaw
    out = {}
    for tag, depth_cfg, ds, shape in (("c2", [1.0, 45.0, 0.5], 8, (2, 2, 64, 96)),
                                      ("coarse", [2.0, 58.0, 2.0], 4, (1, 3, 32, 40))):
        lo, hi, st = depth_cfg
        fake = types.SimpleNamespace()
        fake.D = int(torch.arange(lo, hi, st).numel())
        fake.downsample = ds
        fake.grid_config = {"depth": depth_cfg}
        fake.downsample_depth = lambda d, s, _f=fake: Raw.downsample_depth(_f, d, s)
        B, N, H, W = shape
        d = torch.from_numpy(metric_depth(B, N, H, W, seed=len(out)))
        with torch.no_grad():
            small = Raw.downsample_depth(fake, d, ds)
            dist_ds = Raw.get_two_hot_depth(fake, d, gamma=4, downsample=True)
            dist_nd = Raw.get_two_hot_depth(fake, small, gamma=2.5, downsample=False)
        out[f"{tag}.depth_cfg"] = np.array(depth_cfg, dtype=np.float64)
        out[f"{tag}.downsample"] = np.array(ds)
        out[f"{tag}.depths"] = d.numpy()
        out[f"{tag}.downsampled"] = small.numpy()
        out[f"{tag}.two_hot_g4_ds"] = dist_ds.contiguous().numpy()
        out[f"{tag}.two_hot_g2p5"] = dist_nd.contiguous().numpy()
        print(tag, "D", fake.D, "in", tuple(d.shape), "->", tuple(dist_ds.shape),
              "row sums", float(dist_ds.sum(2).min()), float(dist_ds.sum(2).max()))
    np.savez_compressed(os.path.join(HERE, "two_hot_depth.npz"), **out)


if __name__ == "__main__":
    main()
