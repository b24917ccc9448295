"""Depth-distribution producer: oracle vs the reference's own output (golden,
made by tests/golden/make_golden_depth.py from view_transformer_raw.py:393-429), the torch
autograd route vs the oracle (CPU), and the CUDA kernel vs both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import lift_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "two_hot_depth.npz"))
TAGS = ("c2", "coarse")
TOL = 2e-6      # float32 softmax on the CPU: exp implementations differ by a few ulp
TOL_GPU = 1e-5  # expf + summation order on the GPU; north_star's float bound is 1e-3 relative


def neck_for(tag):
    from veon_b200.view_transformer import LSSViewTransformerRaw
    cfg = [float(v) for v in GOLD[f"{tag}.depth_cfg"]]
    ds = int(GOLD[f"{tag}.downsample"])
    grid = {"x": [-40.0, 40.0, 0.4], "y": [-40.0, 40.0, 0.4], "z": [-1.0, 5.4, 0.4], "depth": cfg}
    return LSSViewTransformerRaw(grid, (64, 96), ds, 8, 64), cfg, ds


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_output(tag):
    cfg, ds = GOLD[f"{tag}.depth_cfg"], int(GOLD[f"{tag}.downsample"])
    assert np.array_equal(O.downsample_depth(GOLD[f"{tag}.depths"], ds), GOLD[f"{tag}.downsampled"])
    a = O.two_hot_depth(GOLD[f"{tag}.depths"], cfg, gamma=4, downsample=ds)
    b = O.two_hot_depth(GOLD[f"{tag}.downsampled"], cfg, gamma=2.5)
    assert a.shape == GOLD[f"{tag}.two_hot_g4_ds"].shape
    assert np.abs(a - GOLD[f"{tag}.two_hot_g4_ds"]).max() <= TOL
    assert np.abs(b - GOLD[f"{tag}.two_hot_g2p5"]).max() <= TOL


@pytest.mark.parametrize("tag", TAGS)
def test_autograd_route_matches_reference_output_on_cpu(tag):
    neck, cfg, ds = neck_for(tag)
    d = torch.from_numpy(GOLD[f"{tag}.depths"]).requires_grad_()
    out = neck.get_two_hot_depth(d, gamma=4, downsample=True)
    assert np.abs(out.detach().numpy() - GOLD[f"{tag}.two_hot_g4_ds"]).max() <= TOL
    out.square().sum().backward()
    assert torch.isfinite(d.grad).all() and float(d.grad.abs().sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_cuda_kernel_matches_reference_output(tag):
    neck, cfg, ds = neck_for(tag)
    d = torch.from_numpy(GOLD[f"{tag}.depths"]).cuda()
    a = neck.get_two_hot_depth(d, gamma=4, downsample=True)
    small = torch.from_numpy(GOLD[f"{tag}.downsampled"]).cuda()
    b = neck.get_two_hot_depth(small, gamma=2.5)
    assert a.is_contiguous() and a.shape == GOLD[f"{tag}.two_hot_g4_ds"].shape
    assert np.abs(a.cpu().numpy() - GOLD[f"{tag}.two_hot_g4_ds"]).max() <= TOL_GPU
    assert np.abs(b.cpu().numpy() - GOLD[f"{tag}.two_hot_g2p5"]).max() <= TOL_GPU
    # the gradient route on the GPU gives the same values
    dg = d.clone().requires_grad_()
    c = neck.get_two_hot_depth(dg, gamma=4, downsample=True)
    assert (c - a).abs().max().item() <= TOL_GPU


@pytest.mark.gpu
def test_cuda_kernel_full_size_vs_oracle():
    """C2-sized input: 8 x 6 maps of 256x704 -> [8,6,88,32,88] (downsample 8)"""
    from veon_b200 import bev_pool as BP
    g = torch.Generator().manual_seed(2)
    d = (torch.rand(2, 6, 256, 704, generator=g) * 50).float()
    d[torch.rand(d.shape, generator=g) < 0.3] = 0.0
    out = BP.two_hot_depth(d.cuda(), [1.0, 45.0, 0.5], 88, gamma=4, downsample=8).cpu().numpy()
    want = O.two_hot_depth(d.numpy(), [1.0, 45.0, 0.5], gamma=4, downsample=8)
    assert out.shape == want.shape == (2, 6, 88, 32, 88)
    assert np.abs(out - want).max() <= TOL_GPU
    s = out.sum(axis=2)
    assert s.max() <= 1.0 + 1e-5
