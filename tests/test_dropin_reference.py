"""DROP-IN checks: the reference's own, UNMODIFIED files executed on top of
our operators, on the GPU.

The three reference files of the path (view_transformer.py, view_transformer_raw.py,
ops/bev_pool_v2/bev_pool.py) are run from /root/reference in the build container and from
oracle/_ref/ on the GPU box (`make -C oracle ref` stages them there next to the reference's
compiled kernels; tests/golden/_ref_loader.py).  Third-party imports are inert stubs.

  route A  (INTEGRATION.md):  reference necks  +  veon_b200.bev_pool.bev_pool_v2
                              (+ our voxel_pooling_prepare_v2 patched in)
  route B:                    reference QuickCumsumCuda (bev_pool.py:11-83, with its memset,
                              permute, argsort)  +  veon_b200.bev_pool_v2_ext
  oracle:                     the same reference QuickCumsumCuda on the reference's OWN
                              kernels (oracle/_ref/libbev_pool_v2_ref.so) -- the unmodified
                              reference operator, end to end, on this GPU.
"""
import ctypes
import os
import types

import numpy as np
import pytest
import torch

from oracle import lift_oracle as O
from veon_b200 import synthetic as S

pytestmark = [pytest.mark.gpu, pytest.mark.needs_reference]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbev_pool_v2_ref.so")
KEYS = ("sensor2ego", "ego2global", "intrins", "post_rots", "post_trans", "bda")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def close_fraction(a, b, atol=1e-4):
    """Share of elements within atol -- the criterion of the reference's own neck test
    (tests/test_models/test_necks/test_necks.py:190-195).  Used where the two sides compute the
    frustum geometry with different float expressions: a point within rounding of a voxel
    border may land in the neighbouring voxel."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) <= atol).mean())


def reference_ext():
    """A `bev_pool_v2_ext` made of the reference's own two launchers (bev_pool_cuda.cu:125-140,
    compiled unmodified): same role as bev_pool.cpp:30-57,74-104."""
    if not os.path.isfile(REF_SO):
        pytest.skip("oracle/_ref not built (make -C oracle ref, needs /root/reference)")
    lib = ctypes.CDLL(REF_SO)
    fwd = getattr(lib, "_Z11bev_pool_v2iiPKfS0_PKiS2_S2_S2_S2_Pf")
    bwd = getattr(lib, "_Z16bev_pool_v2_gradiiPKfS0_S0_PKiS2_S2_S2_S2_PfS3_")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731

    def bev_pool_v2_forward(depth, feat, out, rd, rf, rb, ln, st):
        torch.cuda.synchronize()            # the reference launches on the legacy default stream
        fwd(ctypes.c_int(feat.size(4)), ctypes.c_int(ln.size(0)), vp(depth), vp(feat), vp(rd),
            vp(rf), vp(rb), vp(st), vp(ln), vp(out))
        torch.cuda.synchronize()

    def bev_pool_v2_backward(og, dg, fg, depth, feat, rd, rf, rb, ln, st):
        torch.cuda.synchronize()
        bwd(ctypes.c_int(og.size(4)), ctypes.c_int(ln.size(0)), vp(og), vp(depth), vp(feat),
            vp(rd), vp(rf), vp(rb), vp(st), vp(ln), vp(dg), vp(fg))
        torch.cuda.synchronize()

    return types.SimpleNamespace(bev_pool_v2_forward=bev_pool_v2_forward,
                                 bev_pool_v2_backward=bev_pool_v2_backward, __name__="ref_ext")


def neck_inputs(cfg, B, C, seed):
    g = torch.Generator().manual_seed(seed)
    H, W = cfg.feat_hw
    cal = S.calibration(cfg, batch=B)
    metas = [torch.from_numpy(cal[k]) for k in KEYS]
    depth = torch.softmax(torch.randn(B * cfg.n_cams, cfg.D, H, W, generator=g) * 4, dim=1)
    feat = torch.randn(B * cfg.n_cams, C, H, W, generator=g)
    img = torch.zeros(B, cfg.n_cams, 8, H, W)
    return img, metas, depth, feat


def test_reference_neck_runs_unchanged_on_our_operators():
    """Route A: the reference's LSSViewTransformer (unmodified file) with our bev_pool_v2 bound
    in place of its extension and our prepare patched in; compared with the same reference
    neck on the CPU oracle pooling."""
    from _ref_loader import load_reference_view_transformer
    from veon_b200 import bev_pool as BP

    def cpu_pool(depth, feat, rd, rf, rb, shape, st, ln):
        out = O.bev_pool_v2(depth.numpy(), feat.contiguous().numpy(), rd.numpy(), rf.numpy(),
                            rb.numpy(), tuple(shape), st.numpy(), ln.numpy())
        return torch.from_numpy(out)

    cfg = S.CONFIGS["small"]
    B, C = 1, 16
    img, metas, depth, feat = neck_inputs(cfg, B, C, seed=5)
    kw = dict(grid_config=cfg.grid_config, input_size=cfg.input_size, downsample=cfg.downsample,
              in_channels=8, out_channels=C, collapse_z=False)

    mod = load_reference_view_transformer(cpu_pool)
    ref_neck = mod.LSSViewTransformer(**kw)
    want, _ = ref_neck.view_transform([img] + metas, depth, feat)

    mod = load_reference_view_transformer(BP.bev_pool_v2)
    neck = mod.LSSViewTransformer(**kw)
    cu = [img.cuda()] + [m.cuda() for m in metas]
    got, _ = neck.view_transform(cu, depth.cuda(), feat.cuda())
    # get_lidar_coor runs on GPU vs CPU here (ATen both, different BLAS/inverse kernels)
    assert close_fraction(got.cpu().numpy(), want.numpy()) >= 0.999

    # and with the index preparation replaced as well
    mod.LSSViewTransformer.voxel_pooling_prepare_v2 = lambda self, coor: BP.voxel_pooling_prepare_v2(
        coor, self.grid_lower_bound, self.grid_interval, self.grid_size)
    neck2 = mod.LSSViewTransformer(**kw)
    got2, _ = neck2.view_transform(cu, depth.cuda(), feat.cuda())
    assert torch.equal(got2, got)
    neck2.accelerate = True                                  # reference's cache path on our ranks
    got3, _ = neck2.view_transform(cu, depth.cuda(), feat.cuda())
    assert got3.numel() == got.numel() and torch.equal(got3.reshape(got.shape), got)

    # training through the reference neck: gradients of our operator vs autograd of the oracle
    d = depth.cuda().requires_grad_()
    f = feat.cuda().requires_grad_()
    out, _ = neck2.view_transform(cu, d, f)
    og = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    out.backward(og.cuda())
    H, W = cfg.feat_hw
    coor = neck2.get_lidar_coor(*cu[1:7]).cpu()     # the very coordinates the GPU run pooled
    _, dg, fg = O.torch_cpu_lift(coor, depth.view(B, cfg.n_cams, cfg.D, H, W),
                                 feat.view(B, cfg.n_cams, C, H, W), ref_neck.grid_lower_bound,
                                 ref_neck.grid_interval, ref_neck.grid_size, og.reshape(out.shape))
    assert rel(d.grad.cpu().numpy().reshape(dg.shape), dg.numpy()) <= 1e-3
    assert rel(f.grad.cpu().numpy().reshape(fg.shape), fg.numpy()) <= 1e-3


def test_reference_raw_neck_runs_unchanged_on_our_operators():
    """VEON's neck (view_transformer_raw.py:537-555, unmodified) on our operators equals our
    LSSViewTransformerRaw bit for bit -- forward incl. the 2x2x2 reduction, and the gradients
    (the reference's rearrange + torch.max(dim) backward vs our one-node route)."""
    from _ref_loader import load_reference_view_transformer_raw
    from veon_b200 import bev_pool as BP
    from veon_b200.view_transformer import LSSViewTransformerRaw
    cfg = S.CONFIGS["small"]
    B, C = 1, 32
    H, W = cfg.feat_hw
    img, metas, depth, feat = neck_inputs(cfg, B, C, seed=6)
    mod = load_reference_view_transformer_raw(BP.bev_pool_v2)
    ref = mod.LSSViewTransformerRaw(grid_config=cfg.grid_config, input_size=cfg.input_size,
                                    downsample=cfg.downsample, out_channels=C, collapse_z=False)
    ours = LSSViewTransformerRaw(cfg.grid_config, cfg.input_size, cfg.downsample, 8, C)
    cm = [m.cuda() for m in metas]
    f_a = feat.view(B, cfg.n_cams, C, H, W).cuda().requires_grad_()
    d_a = depth.view(B, cfg.n_cams, cfg.D, H, W).cuda().requires_grad_()
    f_b = f_a.detach().clone().requires_grad_()
    d_b = d_a.detach().clone().requires_grad_()
    with torch.no_grad():           # the reference's own geometry (ATen ops): float tolerance
        loose = ref([f_a.detach()] + cm, d_a.detach())
    # same coordinates for both (our geometry kernel agrees with the reference's ATen ops to
    # rounding only, which moves border points between voxels): from here on, bit for bit
    ref.get_lidar_coor = ours.get_lidar_coor
    out_ref = ref([f_a] + cm, d_a)
    out = ours([f_b] + cm, d_b)
    assert close_fraction(loose.cpu().numpy(), out.detach().cpu().numpy()) >= 0.999
    assert out.shape == (B, C, 8, 100, 100)
    assert torch.equal(out, out_ref)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(2)).cuda()
    out_ref.backward(go)
    out.backward(go)
    assert torch.equal(f_b.grad, f_a.grad)
    assert torch.equal(d_b.grad, d_a.grad)


@pytest.mark.parametrize("cfg_name,batch,C", [("small", 2, 64), ("small", 2, 256), ("C1", 1, 512),
                                              ("C3", 1, 256), ("C1", 1, 768)])
def test_reference_operator_on_its_own_kernels_vs_ours(cfg_name, batch, C):
    """The unmodified reference operator (bev_pool.py: QuickCumsumCuda + bev_pool_v2) run
    (1) on the reference's own CUDA kernels, (2) on our bev_pool_v2_ext stand-in (route B),
    against (3) our bev_pool_v2.  Forward bit-identical three ways; gradients element-wise
    within fp32 re-association -- at the channel widths VEON uses (256, 512, 768)."""
    from _ref_loader import load_reference_bev_pool
    from veon_b200 import bev_pool as BP, bev_pool_v2_ext as our_ext
    cfg = S.CONFIGS[cfg_name]
    coor = S.lidar_coor_np(cfg, batch=batch)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    ranks = [torch.from_numpy(r).cuda() for r in O.prepare_v2(coor, lower, interval, size)]
    rb, rd, rf, st, ln = ranks
    B, N, D, H, W, _ = coor.shape
    shape = (B, 16, 200, 200, C)
    assert B * 640000 * C < 2 ** 31          # the reference kernels index with 32-bit ints
    g = torch.Generator().manual_seed(13)
    depth = torch.softmax(torch.randn(B, N, D, H, W, generator=g) * 4, dim=2).cuda()
    feat = torch.randn(B, N, H, W, C, generator=g).cuda()
    og = torch.randn(B, C, 16, 200, 200, generator=g).cuda()

    def run(fn):
        d = depth.detach().clone().requires_grad_()
        f = feat.detach().clone().requires_grad_()
        out = fn(d, f, rd, rf, rb, shape, st, ln)
        out.backward(og)
        torch.cuda.synchronize()
        return out.detach(), d.grad, f.grad

    o_ref, dg_ref, fg_ref = run(load_reference_bev_pool(reference_ext()).bev_pool_v2)
    o_b, dg_b, fg_b = run(load_reference_bev_pool(our_ext).bev_pool_v2)
    o_us, dg_us, fg_us = run(BP.bev_pool_v2)
    assert o_ref.shape == (B, C, 16, 200, 200)
    assert torch.equal(o_b, o_ref) and torch.equal(o_us, o_ref)
    for dg, fg in ((dg_b, fg_b), (dg_us, fg_us)):
        assert rel(dg.cpu().numpy(), dg_ref.cpu().numpy()) <= 2e-5
        assert rel(fg.cpu().numpy(), fg_ref.cpu().numpy()) <= 2e-5
