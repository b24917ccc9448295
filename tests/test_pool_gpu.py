"""GPU parity: bev_pool_v2 forward / backward through the C ABI against the C
oracle, the reference's known-answer test, and (when oracle/_ref was built)
the reference's own CUDA kernels compiled unmodified.

Tolerances (north_star): float results within 1e-3 relative max-abs.  The
forward is in fact bit-exact (same fma order as the reference kernel)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import lift_oracle as O
from veon_b200 import synthetic as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbev_pool_v2_ref.so")


def rel_max_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_case(cfg_name, batch, C, seed=0, peaky=False):
    cfg = S.CONFIGS[cfg_name]
    coor = S.lidar_coor_np(cfg, batch=batch)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    ranks = O.prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    g = torch.Generator().manual_seed(seed)
    depth = torch.softmax(torch.randn(B, N, D, H, W, generator=g) * 4, dim=2)
    feat = torch.randn(B, N, H, W, C, generator=g)
    Z, Y, X = int(size[2]), int(size[1]), int(size[0])
    return dict(coor=coor, grid=(lower, interval, size), ranks=ranks, depth=depth, feat=feat,
                shape=(B, Z, Y, X, C), dims=(B, N, D, H, W))


def gpu_ranks(ranks):
    return [torch.from_numpy(r).cuda() for r in ranks]


def run_gpu(case, out_grad=None):
    from veon_b200.bev_pool import bev_pool_v2
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    depth = case["depth"].cuda().requires_grad_(out_grad is not None)
    feat = case["feat"].cuda().requires_grad_(out_grad is not None)
    out = bev_pool_v2(depth, feat, rd, rf, rb, case["shape"], st, ln)
    if out_grad is None:
        return out.detach().cpu().numpy(), None, None
    out.backward(out_grad.cuda())
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), depth.grad.cpu().numpy(), feat.grad.cpu().numpy()


def test_known_answer_test(golden_dir):
    """mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176, verbatim values"""
    from veon_b200.bev_pool import bev_pool_v2
    with open(os.path.join(golden_dir, "kat_bev_pool_v2.json")) as f:
        k = json.load(f)
    depth = torch.tensor(k["depth"]).float().cuda().view(*k["depth_shape"]).requires_grad_()
    feat = torch.ones(size=k["feat_shape"], dtype=torch.float, device="cuda").requires_grad_()
    rd, rf, rb = (torch.tensor(k[n]).int().cuda() for n in ("ranks_depth", "ranks_feat", "ranks_bev"))
    kept = torch.ones(rb.shape[0], device="cuda", dtype=torch.bool)
    kept[1:] = rb[1:] != rb[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb.shape[0] - starts[-1]
    bev = bev_pool_v2(depth, feat, rd, rf, rb, tuple(k["bev_feat_shape"]), starts, lengths)
    assert bev.shape == (1, 2, 1, 2, 2) and bev.is_contiguous()
    loss = torch.sum(bev)
    loss.backward()
    assert abs(loss.item() - k["loss"]) < 1e-6
    assert depth.grad.allclose(torch.tensor(k["grad_depth"]).cuda().view(1, 1, 2, 2, 2))
    assert feat.grad.allclose(torch.tensor(k["grad_feat"]).cuda().view(1, 1, 2, 2, 2))


@pytest.mark.parametrize("cfg_name,batch,C", [("tiny", 2, 32), ("tiny", 2, 7), ("small", 2, 64),
                                              ("tiny", 1, 100), ("C1", 1, 64), ("small", 1, 256),
                                              # narrow rows: the lane-per-voxel kernel (C <= 32, C % 4 == 0)
                                              ("tiny", 2, 20), ("small", 2, 4), ("small", 2, 20),
                                              ("C1", 2, 28)])
def test_forward_bit_exact_vs_oracle(cfg_name, batch, C):
    case = make_case(cfg_name, batch, C)
    rb, rd, rf, st, ln = case["ranks"]
    want = O.bev_pool_v2(case["depth"].numpy(), case["feat"].numpy(), rd, rf, rb, case["shape"], st, ln)
    got, _, _ = run_gpu(case)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("cfg_name,batch,C", [("tiny", 2, 32), ("tiny", 2, 7), ("small", 2, 64),
                                              ("C1", 1, 64), ("small", 1, 160),
                                              # the widths VEON uses (its neck: 256; CLIP dims 512 / 768),
                                              # several samples each: the multi-chunk, multi-sample backward
                                              ("small", 2, 256), ("small", 3, 512), ("tiny", 2, 768),
                                              ("C1", 2, 256), ("tiny", 3, 96), ("small", 2, 100)])
def test_backward_vs_oracle(cfg_name, batch, C):
    case = make_case(cfg_name, batch, C, seed=1)
    rb, rd, rf, st, ln = case["ranks"]
    B, Z, Y, X, _ = case["shape"]
    og = torch.randn(B, C, Z, Y, X, generator=torch.Generator(device="cuda").manual_seed(2),
                     device="cuda").cpu()
    dg_want, fg_want = O.bev_pool_v2_backward(og.numpy(), case["depth"].numpy(),
                                              case["feat"].numpy(), rd, rf, rb)
    _, dg, fg = run_gpu(case, og)
    assert rel_max_err(dg, dg_want) <= 1e-3      # north_star tolerance
    assert rel_max_err(fg, fg_want) <= 1e-3
    # and much tighter in practice (fp32 re-association only)
    assert rel_max_err(dg, dg_want) <= 2e-5 and rel_max_err(fg, fg_want) <= 2e-5
    # dropped points get exactly zero depth gradient
    kept = np.zeros(dg.size, bool); kept[rd] = True
    assert np.all(dg.reshape(-1)[~kept] == 0)


def test_backward_is_bitwise_deterministic():
    case = make_case("small", 2, 64, seed=3)
    B, Z, Y, X, C = case["shape"]
    og = torch.randn(B, C, Z, Y, X, generator=torch.Generator().manual_seed(4))
    a = run_gpu(case, og)
    b = run_gpu(case, og)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_noncontiguous_feat_and_channels_last_grad():
    """feat arrives as a permuted view (view_transformer.py:190); the upstream
    gradient may arrive with arbitrary strides"""
    from veon_b200.bev_pool import bev_pool_v2
    case = make_case("tiny", 2, 16, seed=5)
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    feat_bnchw = case["feat"].permute(0, 1, 4, 2, 3).contiguous().cuda().requires_grad_()
    depth = case["depth"].cuda().requires_grad_()
    out = bev_pool_v2(depth, feat_bnchw.permute(0, 1, 3, 4, 2), rd, rf, rb, case["shape"], st, ln)
    B, Z, Y, X, C = case["shape"]
    og = torch.randn(B, Z, Y, X, C, generator=torch.Generator().manual_seed(6)).cuda()
    out.backward(og.permute(0, 4, 1, 2, 3))     # non-contiguous channels-first view
    r = case["ranks"]
    dg_want, fg_want = O.bev_pool_v2_backward(og.permute(0, 4, 1, 2, 3).cpu().numpy(),
                                              case["depth"].numpy(), case["feat"].numpy(),
                                              r[1], r[2], r[0])
    assert rel_max_err(depth.grad.cpu().numpy(), dg_want) <= 2e-5
    assert rel_max_err(feat_bnchw.grad.permute(0, 1, 3, 4, 2).cpu().numpy(), fg_want) <= 2e-5


def test_generic_path_for_unsorted_ranks():
    """shuffle the INTERVAL order: ranks_bev no longer sorted -> plan rejects,
    literal interval-driven kernels are used; result must not change."""
    from veon_b200 import bev_pool as BP
    case = make_case("tiny", 2, 24, seed=7)
    rb, rd, rf, st, ln = case["ranks"]
    rng = np.random.RandomState(0)
    perm = rng.permutation(st.size)
    new_rb, new_rd, new_rf, new_st, new_ln = [], [], [], [], []
    pos = 0
    for k in perm:
        s, l = int(st[k]), int(ln[k])
        new_rb.append(rb[s:s + l]); new_rd.append(rd[s:s + l]); new_rf.append(rf[s:s + l])
        new_st.append(pos); new_ln.append(l); pos += l
    shuf = (np.concatenate(new_rb), np.concatenate(new_rd), np.concatenate(new_rf),
            np.array(new_st, np.int32), np.array(new_ln, np.int32))
    want = O.bev_pool_v2(case["depth"].numpy(), case["feat"].numpy(), rd, rf, rb, case["shape"], st, ln)
    B, Z, Y, X, C = case["shape"]
    og = torch.randn(B, C, Z, Y, X, generator=torch.Generator().manual_seed(8))
    dg_want, fg_want = O.bev_pool_v2_backward(og.numpy(), case["depth"].numpy(), case["feat"].numpy(), rd, rf, rb)
    case2 = dict(case, ranks=shuf)
    got, dg, fg = run_gpu(case2, og)
    np.testing.assert_array_equal(got, want)
    assert rel_max_err(dg, dg_want) <= 2e-5 and rel_max_err(fg, fg_want) <= 2e-5
    srb, srd, srf, sst, sln = gpu_ranks(shuf)
    plan = BP._plan_for(srd, srf, srb, sst, sln, case["dims"], Z * Y * X)
    assert plan.flags & 1 and not plan.ok


def _ref_lib():
    if not os.path.isfile(REF_SO):
        pytest.skip("oracle/_ref not built (make -C oracle ref, needs /root/reference)")
    lib = ctypes.CDLL(REF_SO)
    return (getattr(lib, "_Z11bev_pool_v2iiPKfS0_PKiS2_S2_S2_S2_Pf"),
            getattr(lib, "_Z16bev_pool_v2_gradiiPKfS0_S0_PKiS2_S2_S2_S2_PfS3_"))


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("cfg_name,batch,C", [("small", 2, 64), ("C1", 2, 64), ("C3", 1, 128)])
def test_against_reference_cuda_kernels(cfg_name, batch, C):
    """The reference's own bev_pool_cuda.cu (compiled unmodified into
    oracle/_ref) run on the same ranks: forward must be bit-identical,
    backward within fp32 re-association."""
    fwd, bwd = _ref_lib()
    from veon_b200.bev_pool import bev_pool_v2
    case = make_case(cfg_name, batch, C, seed=9)
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    depth = case["depth"].cuda().requires_grad_()
    feat = case["feat"].cuda().requires_grad_()
    B, Z, Y, X, _ = case["shape"]
    # --- reference forward: channels-last, caller-zeroed, default stream
    ref_out = torch.zeros(B, Z, Y, X, C, device="cuda")
    torch.cuda.synchronize()
    fwd(ctypes.c_int(C), ctypes.c_int(st.numel()), _vp(depth), _vp(feat), _vp(rd), _vp(rf), _vp(rb),
        _vp(st), _vp(ln), _vp(ref_out))
    torch.cuda.synchronize()
    out = bev_pool_v2(depth, feat, rd, rf, rb, case["shape"], st, ln)
    assert torch.equal(out, ref_out.permute(0, 4, 1, 2, 3))
    # --- reference backward with its Python-side index work (bev_pool.py:47-57)
    og = torch.randn(B, C, Z, Y, X, generator=torch.Generator().manual_seed(10)).cuda()
    out.backward(og)
    order = rf.argsort()
    rf_s, rd_s, rb_s = rf[order].contiguous(), rd[order].contiguous(), rb[order].contiguous()
    kept = torch.ones(rb_s.shape[0], device="cuda", dtype=torch.bool)
    kept[1:] = rf_s[1:] != rf_s[:-1]
    st_bp = torch.where(kept)[0].int()
    ln_bp = torch.zeros_like(st_bp)
    ln_bp[:-1] = st_bp[1:] - st_bp[:-1]
    ln_bp[-1] = rb_s.shape[0] - st_bp[-1]
    dgr, fgr = torch.zeros_like(depth), torch.zeros_like(feat)
    og_cl = og.permute(0, 2, 3, 4, 1).contiguous()
    torch.cuda.synchronize()
    bwd(ctypes.c_int(C), ctypes.c_int(st_bp.numel()), _vp(og_cl), _vp(depth), _vp(feat), _vp(rd_s),
        _vp(rf_s), _vp(rb_s), _vp(st_bp), _vp(ln_bp), _vp(dgr), _vp(fgr))
    torch.cuda.synchronize()
    assert rel_max_err(depth.grad.cpu().numpy(), dgr.cpu().numpy()) <= 2e-5
    assert rel_max_err(feat.grad.cpu().numpy(), fgr.cpu().numpy()) <= 2e-5


def test_literal_c_abi_entry_points_match_reference_layout():
    """veon_bev_pool_v2 / veon_bev_pool_v2_grad: channels-last, caller-zeroed,
    intervals as given -- what bev_pool.cpp:7-14 would bind."""
    from veon_b200 import _lib
    lib = _lib.load()
    case = make_case("tiny", 2, 20, seed=11)
    rbn, rdn, rfn, stn, lnn = case["ranks"]
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    depth, feat = case["depth"].cuda(), case["feat"].cuda()
    B, Z, Y, X, C = case["shape"]
    out = torch.zeros(B, Z, Y, X, C, device="cuda")
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.veon_bev_pool_v2(C, st.numel(), _vp(depth), _vp(feat), _vp(rd), _vp(rf), _vp(rb),
                              _vp(st), _vp(ln), _vp(out), s)
    assert rc == 0
    want = O.bev_pool_v2_channels_last(case["depth"].numpy(), case["feat"].numpy(), rdn, rfn, rbn,
                                       case["shape"], stn, lnn)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    og = torch.randn(B, C, Z, Y, X, generator=torch.Generator().manual_seed(12))
    bp = O.bp_intervals(rdn, rfn, rbn)
    rd2, rf2, rb2, st2, ln2 = gpu_ranks(bp)
    dg, fg = torch.zeros_like(depth), torch.zeros_like(feat)
    og_cl = og.permute(0, 2, 3, 4, 1).contiguous().cuda()
    rc = lib.veon_bev_pool_v2_grad(C, st2.numel(), _vp(og_cl), _vp(depth), _vp(feat), _vp(rd2),
                                   _vp(rf2), _vp(rb2), _vp(st2), _vp(ln2), _vp(dg), _vp(fg), s)
    assert rc == 0
    dg_want, fg_want = O.bev_pool_v2_backward(og.numpy(), case["depth"].numpy(),
                                              case["feat"].numpy(), rdn, rfn, rbn)
    np.testing.assert_array_equal(fg.cpu().numpy(), fg_want)   # same fma order as the reference
    assert rel_max_err(dg.cpu().numpy(), dg_want) <= 2e-5


def test_full_size_properties_c2():
    """BASELINE configs[1] (6 cams 16x44, D=88, C=64, B=8): linearity and the
    adjoint identity <pool(d,f), g> = <d, dgrad> = <f, fgrad> -- size-independent
    checks that need no CPU oracle at this size."""
    from veon_b200.bev_pool import bev_pool_v2, voxel_pooling_prepare_v2
    cfg = S.CONFIGS["C2"]
    coor = torch.from_numpy(S.lidar_coor_np(cfg)).cuda()
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    C = cfg.channels
    shape = (B, 16, 200, 200, C)
    g = torch.Generator(device="cuda").manual_seed(0)
    depth = torch.rand(B, N, D, H, W, device="cuda", generator=g).requires_grad_()
    feat = torch.randn(B, N, H, W, C, device="cuda", generator=g).requires_grad_()
    out = bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)
    assert out.shape == (B, C, 16, 200, 200) and out.is_contiguous()
    # empty voxels are exactly zero; occupied count matches the intervals
    occ = (out.detach().abs().sum(1) > 0).sum().item()
    assert occ <= st.numel() and occ >= 0.99 * st.numel()
    # linearity in feat
    out2 = bev_pool_v2(depth, 2.0 * feat, rd, rf, rb, shape, st, ln)
    assert torch.equal(out2, 2.0 * out)
    # adjoint identity
    og = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(og)
    lhs = (out.detach().double() * og.double()).sum().item()
    via_feat = (feat.grad.double() * feat.detach().double()).sum().item()
    via_depth = (depth.grad.double() * depth.detach().double()).sum().item()
    assert abs(lhs - via_feat) <= 1e-6 * abs(lhs) + 1e-3
    assert abs(lhs - via_depth) <= 1e-6 * abs(lhs) + 1e-3


def test_full_size_c2_backward_elementwise_per_sample():
    """BASELINE configs[1] at full size (B=8, C=64): the gradients of the first, a middle and the
    last sample against the C oracle run on that sample alone (every sample is independent:
    the batch index is only the top digit of ranks_bev, view_transformer.py:241)."""
    from veon_b200.bev_pool import bev_pool_v2, voxel_pooling_prepare_v2
    cfg = S.CONFIGS["C2"]
    coor_np = S.lidar_coor_np(cfg)
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor = torch.from_numpy(coor_np).cuda()
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor_np.shape
    C = cfg.channels
    g = torch.Generator(device="cuda").manual_seed(21)
    depth = torch.softmax(torch.randn(B, N, D, H, W, device="cuda", generator=g) * 4, dim=2).requires_grad_()
    feat = torch.randn(B, N, H, W, C, device="cuda", generator=g).requires_grad_()
    og = torch.randn(B, C, 16, 200, 200, device="cuda", generator=g)
    out = bev_pool_v2(depth, feat, rd, rf, rb, (B, 16, 200, 200, C), st, ln)
    out.backward(og)
    for b in (0, 3, B - 1):
        r1 = O.prepare_v2(coor_np[b:b + 1], lower, interval, size)
        d1 = depth.detach()[b:b + 1].cpu().numpy()
        f1 = feat.detach()[b:b + 1].cpu().numpy()
        want = O.bev_pool_v2(d1, f1, r1[1], r1[2], r1[0], (1, 16, 200, 200, C), r1[3], r1[4])
        np.testing.assert_array_equal(out.detach()[b:b + 1].cpu().numpy(), want)
        dg, fg = O.bev_pool_v2_backward(og[b:b + 1].cpu().numpy(), d1, f1, r1[1], r1[2], r1[0])
        assert rel_max_err(depth.grad[b:b + 1].cpu().numpy(), dg) <= 2e-5
        assert rel_max_err(feat.grad[b:b + 1].cpu().numpy(), fg) <= 2e-5


@pytest.mark.parametrize("name,batch", [("C3", 1), ("C4", 1)])
def test_full_size_properties_wide_channels(name, batch):
    """BASELINE configs[2]/[3] shapes (32x88 feats, C=512 / D=118, C=768; one sample):
    multi-chunk channel path at full size.  Checked by size-independent properties
    (empty voxels exactly zero, linearity, adjoint identity) and, on a slice of
    channels, bit-exactness against the C oracle."""
    from veon_b200.bev_pool import bev_pool_v2, voxel_pooling_prepare_v2
    cfg = S.CONFIGS[name]
    coor_np = S.lidar_coor_np(cfg, batch=batch)
    coor = torch.from_numpy(coor_np).cuda()
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    rb, rd, rf, st, ln = voxel_pooling_prepare_v2(coor, lower, interval, size)
    B, N, D, H, W, _ = coor.shape
    C = cfg.channels
    shape = (B, 16, 200, 200, C)
    g = torch.Generator(device="cuda").manual_seed(0)
    depth = torch.rand(B, N, D, H, W, device="cuda", generator=g).requires_grad_()
    feat = torch.randn(B, N, H, W, C, device="cuda", generator=g).requires_grad_()
    out = bev_pool_v2(depth, feat, rd, rf, rb, shape, st, ln)
    assert out.shape == (B, C, 16, 200, 200) and out.is_contiguous()
    occ_vox = torch.zeros(B * 640000, dtype=torch.bool, device="cuda")
    occ_vox[rb.long()] = True
    assert float(out.detach().view(B, C, -1)[:, :, ~occ_vox.view(B, -1)[0]].abs().max()) == 0.0
    # a slice of channels against the oracle (bit-exact), incl. the last partial chunk
    sel = [0, 1, 63, 64, 200, C - 65, C - 1]
    feat_sel = feat.detach()[..., sel].contiguous()
    want = O.bev_pool_v2(depth.detach().cpu().numpy(), feat_sel.cpu().numpy(), rd.cpu().numpy(),
                         rf.cpu().numpy(), rb.cpu().numpy(), (B, 16, 200, 200, len(sel)),
                         st.cpu().numpy(), ln.cpu().numpy())
    np.testing.assert_array_equal(out.detach()[:, sel].cpu().numpy(), want)
    og = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(og)
    lhs = (out.detach().double() * og.double()).sum().item()
    via_feat = (feat.grad.double() * feat.detach().double()).sum().item()
    via_depth = (depth.grad.double() * depth.detach().double()).sum().item()
    assert abs(lhs - via_feat) <= 1e-6 * abs(lhs) + 1e-2
    assert abs(lhs - via_depth) <= 1e-6 * abs(lhs) + 1e-2
    # ... and element by element against the C oracle (bev_pool_cuda.cu:67-121 restated)
    dg, fg = O.bev_pool_v2_backward(og.cpu().numpy(), depth.detach().cpu().numpy(),
                                    feat.detach().cpu().numpy(), rd.cpu().numpy(),
                                    rf.cpu().numpy(), rb.cpu().numpy())
    assert rel_max_err(depth.grad.cpu().numpy(), dg) <= 2e-5
    assert rel_max_err(feat.grad.cpu().numpy(), fg) <= 2e-5


# ---------------------------------------------------------------- heavy tiles
def _heavy_ids(plan):
    h = plan.tile_heavy.cpu().numpy()
    return int(h[0]), int(h[1]), np.sort(h[2:2 + int(h[0])])


def test_heavy_tile_list_matches_point_counts():
    """The plan lists exactly the tiles whose point count reaches the threshold
    (veon_lift.h: tile_heavy), from both plan builders."""
    from veon_b200 import bev_pool as BP
    case = make_case("C1", 2, 8)
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    B, Z, Y, X, _ = case["shape"]
    plan = BP._plan_for(rd, rf, rb, st, ln, case["dims"], Z * Y * X)
    n, thr, ids = _heavy_ids(plan)
    cnt = np.bincount(case["ranks"][0] // 32, minlength=B * Z * Y * X // 32)
    want = np.nonzero(cnt >= thr)[0]
    assert thr >= 32 and n == want.size and n > 0
    assert np.array_equal(ids, want)
    lower, interval, size = case["grid"]
    prep = BP.prepare_ranks(torch.from_numpy(case["coor"]).cuda(), lower, interval, size)
    n2, thr2, ids2 = _heavy_ids(prep.plan)
    assert (n2, thr2) == (n, thr) and np.array_equal(ids2, want)


@pytest.mark.parametrize("C", [64, 80, 6, 200, 20, 32])
def test_heavy_tile_kernel_is_bit_identical_to_the_warp_path(C):
    """The CTA-per-tile kernel for heavy tiles only re-schedules the loads: with and
    without the heavy list the volume is the same bit for bit; the call without the list (and
    without a workspace) takes the general lane-per-channel kernel, the planned call the
    streaming two-role kernel at C=64 -- so this also pins those two against each other (C=80: ragged last
    channel chunk; C=6: rows not 16-byte sized, the heavy list is ignored; C=20, 32: with the
    list the main grid is the lane-per-voxel kernel for narrow rows, without it the
    lane-per-channel one, so this also checks those two against each other)."""
    from veon_b200 import _lib, bev_pool as BP
    case = make_case("C1", 1, C, seed=3)
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    B, Z, Y, X, _ = case["shape"]
    V = Z * Y * X
    plan = BP._plan_for(rd, rf, rb, st, ln, case["dims"], V)
    assert _heavy_ids(plan)[0] > 0
    depth, feat = case["depth"].cuda(), case["feat"].cuda()
    with_heavy = BP._fwd_planar(depth, feat, rd, rf, rb, plan, B, C, V, (B, C, Z, Y, X))
    lib = _lib.load()
    without = torch.full((B, C, Z, Y, X), float("nan"), device="cuda")
    rc = lib.veon_bev_pool_v2_fwd_planar(
        BP._ptr(depth), BP._ptr(feat), BP._ptr(rd), BP._ptr(rf), BP._ptr(rb),
        BP._ptr(plan.tile_start), None, None, None, 0, B, C, V, feat.numel() // C,
        BP._ptr(without), None, 0, BP._stream_ptr(depth.device))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(with_heavy, without)
    want = O.bev_pool_v2(case["depth"].numpy(), case["feat"].numpy(), *case["ranks"][1:3],
                         case["ranks"][0], case["shape"], *case["ranks"][3:])
    assert np.array_equal(with_heavy.cpu().numpy(), np.ascontiguousarray(want))


def test_heavy_tile_with_thousands_of_points_in_one_voxel():
    """Many rounds of 128 points, all in one voxel + a second voxel that starts in the
    middle of a round: partial sums carried across rounds keep the rank order."""
    from veon_b200.bev_pool import bev_pool_v2
    B, N, D, H, W, C = 1, 1, 40, 8, 16, 64
    Z, Y, X = 1, 2, 64
    g = torch.Generator().manual_seed(5)
    depth = torch.rand(B, N, D, H, W, generator=g)
    feat = torch.randn(B, N, H, W, C, generator=g)
    P = D * H * W
    n_a = 3000
    rd = torch.randperm(P, generator=g)[:n_a + 333].int()
    rd[:n_a] = rd[:n_a].sort().values
    rd[n_a:] = rd[n_a:].sort().values
    rf = (rd % (H * W)).int()
    rb = torch.cat([torch.full((n_a,), 37), torch.full((333,), 41)]).int()
    starts = torch.tensor([0, n_a]).int()
    lengths = torch.tensor([n_a, 333]).int()
    out = bev_pool_v2(depth.cuda(), feat.cuda(), rd.cuda(), rf.cuda(), rb.cuda(),
                      (B, Z, Y, X, C), starts.cuda(), lengths.cuda())
    want = O.bev_pool_v2(depth.numpy(), feat.numpy(), rd.numpy(), rf.numpy(), rb.numpy(),
                         (B, Z, Y, X, C), starts.numpy(), lengths.numpy())
    assert np.array_equal(out.cpu().numpy(), np.ascontiguousarray(want))


@pytest.mark.parametrize("C", [64, 20])
def test_next_kernel_in_the_stream_sees_the_whole_volume(C):
    """The forward is two grids, the second a programmatic dependent of the first that does not
    wait for it.  Whatever is launched next on the stream must still find every tile written:
    clone the volume with no synchronisation in between, many times, at full size.  (C=20: the
    narrow-row pair, where the heavy grid is the first launch and the main grid the dependent.)"""
    from veon_b200 import bev_pool as BP
    cfg = S.CONFIGS["C2"]
    B = 8
    lower, interval, size = S.grid_vectors(cfg.grid_config)
    coor = torch.from_numpy(S.lidar_coor_np(cfg, batch=B)).cuda()
    _, N, D, H, W, _ = coor.shape
    g = torch.Generator(device="cuda").manual_seed(0)
    depth = torch.softmax(torch.randn(B, N, D, H, W, device="cuda", generator=g) * 4, dim=2)
    feat = torch.randn(B, N, H, W, C, device="cuda", generator=g)
    prep = BP.prepare_ranks(coor, lower, interval, size)
    assert int(prep.plan.tile_heavy[0]) > 0
    shape = (B, 16, 200, 200, C)
    ref = BP.pool_prepared(depth, feat, prep, shape).clone()
    torch.cuda.synchronize()
    scratch = torch.empty(64 << 20, device="cuda")
    for i in range(12):
        scratch.normal_()                               # unrelated work in front
        vol = BP.pool_prepared(depth, feat, prep, shape)
        copy = vol.clone()                              # the very next kernel reads the volume
        torch.cuda.synchronize()
        assert torch.equal(copy, ref), f"iteration {i}: a consumer ran before the volume was complete"


@pytest.mark.parametrize("C", [64, 20])
def test_forward_inside_a_cuda_graph(C):
    """Stream capture: the two forward grids and their programmatic edge go into a CUDA graph
    (the heavy grid then joins the main grid before it completes); replays must reproduce the
    eager volume, including through the node captured right behind them.  (C=20: narrow rows.)"""
    from veon_b200 import bev_pool as BP
    case = make_case("C1", 2, C, seed=7)
    rb, rd, rf, st, ln = gpu_ranks(case["ranks"])
    B, Z, Y, X, C = case["shape"]
    V = Z * Y * X
    plan = BP._plan_for(rd, rf, rb, st, ln, case["dims"], V)
    assert plan.ok and int(plan.tile_heavy[0]) > 0
    depth, feat = case["depth"].cuda(), case["feat"].cuda()
    eager = BP._fwd_planar(depth, feat, rd, rf, rb, plan, B, C, V, (B, C, Z, Y, X)).clone()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        BP._fwd_planar(depth, feat, rd, rf, rb, plan, B, C, V, (B, C, Z, Y, X))   # warm-up
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        vol = BP._fwd_planar(depth, feat, rd, rf, rb, plan, B, C, V, (B, C, Z, Y, X))
        copy = vol.clone()
    for _ in range(5):
        copy.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(copy, eager)
