"""TEST INFRASTRUCTURE ONLY -- numpy / ctypes CPU restatement of the lifting
path.  See oracle/__init__.py for the rules and the pinning status.

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32


# --------------------------------------------------------------------------
# a3  voxel_pooling_prepare_v2      mmdet3d/models/necks/view_transformer.py:202-260
# --------------------------------------------------------------------------
def prepare_v2(coor, grid_lower_bound, grid_interval, grid_size):
    """coor [B,N,D,H,W,3] float32 -> (ranks_bev, ranks_depth, ranks_feat,
    interval_starts, interval_lengths), all int32, or 5x None.

    Every arithmetic step keeps the dtype the reference's torch code has:
      :225-227  (coor - lower) / interval in float32, `.long()` = truncation
                toward zero (NaN / overflow -> INT64_MIN on x86, i.e. dropped)
      :233-235  bounds test of the int64 index against the FLOAT32 grid_size
      :241-244  rank accumulated in FLOAT32 (int64 * 0-dim float32 tensor)
      :245      argsort -- unstable in the reference; the stable order is used
                here (`canonical_order` maps any valid order onto it)
      :249-257  head flags on the float ranks, interval starts / lengths
      :258-260  `.int()`
    """
    coor = np.ascontiguousarray(coor, dtype=_f32)
    B, N, D, H, W, _ = coor.shape
    P = B * N * D * H * W
    if P == 0:                                        # :236-237 `len(kept) == 0`
        return (None,) * 5
    lo = np.asarray(grid_lower_bound, dtype=_f32)
    iv = np.asarray(grid_interval, dtype=_f32)
    gs = np.asarray(grid_size, dtype=_f32)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        vox = (coor.reshape(P, 3) - lo) / iv          # float32, IEEE
        idx = vox.astype(np.int64)                    # cvttss2si semantics
    bad = ~np.isfinite(vox) | (np.abs(vox) >= _f32(2.0**63))
    idx[bad] = np.iinfo(np.int64).min
    kept = np.ones(P, dtype=bool)
    for a in range(3):
        kept &= (idx[:, a] >= 0) & (idx[:, a].astype(_f32) < gs[a])
    ranks_depth = np.arange(P, dtype=np.int32)[kept]
    # :218-223  feature-pixel index of point (b,n,d,h,w) = (b*N+n)*H*W + h*W + w
    HW = H * W
    pt = np.arange(P, dtype=np.int64)
    ranks_feat = ((pt // (D * HW)) * HW + pt % HW).astype(np.int32)[kept]
    batch = (pt // (P // B))[kept].astype(_f32)
    x, y, z = (idx[kept, a].astype(_f32) for a in range(3))
    g_zyx = _f32(_f32(gs[2] * gs[1]) * gs[0])
    g_yx = _f32(gs[1] * gs[0])
    r = batch * g_zyx                                  # float32 throughout
    r = r + z * g_yx
    r = r + (y * gs[0] + x)
    order = np.argsort(r, kind="stable")
    r, ranks_depth, ranks_feat = r[order], ranks_depth[order], ranks_feat[order]
    if r.shape[0] == 0:                                # :253-254
        return (None,) * 5
    head = np.ones(r.shape[0], dtype=bool)
    head[1:] = r[1:] != r[:-1]
    starts = np.flatnonzero(head).astype(np.int32)
    lengths = np.empty_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = r.shape[0] - starts[-1]
    return (r.astype(np.int32), ranks_depth.astype(np.int32), ranks_feat.astype(np.int32),
            starts, lengths)


def canonical_order(ranks_bev, ranks_depth, ranks_feat):
    """Re-order a (possibly unstably sorted) rank triple into the canonical
    order: by ranks_bev, then ascending ranks_depth.  The reference's argsort
    (:245) leaves the in-voxel order arbitrary, so parity of ranks_depth /
    ranks_feat is a per-interval multiset statement."""
    order = np.lexsort((ranks_depth, ranks_bev))
    return ranks_bev[order], ranks_depth[order], ranks_feat[order]


# --------------------------------------------------------------------------
# a6/a7 + a8/a9  pooling kernels, via the C restatement (pool_oracle.c)
# --------------------------------------------------------------------------
_clib = None


def _pool_lib():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "libpool_oracle.so")
        if not os.path.isfile(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        _clib = ctypes.CDLL(path)
    return _clib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def bev_pool_v2_channels_last(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                              bev_feat_shape, interval_starts, interval_lengths):
    """QuickCumsumCuda.forward (bev_pool.py:17-41): zeros + kernel
    (bev_pool_cuda.cu:21-48).  Returns [B,Z,Y,X,C] float32."""
    depth = np.ascontiguousarray(depth, dtype=_f32)
    feat = np.ascontiguousarray(feat, dtype=_f32)
    c = feat.shape[-1]
    out = np.zeros(bev_feat_shape, dtype=_f32)
    ints = [np.ascontiguousarray(a, dtype=np.int32)
            for a in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)]
    _pool_lib().oracle_bev_pool_v2(ctypes.c_int(c), ctypes.c_int(ints[3].shape[0]), _p(depth),
                                   _p(feat), *[_p(a) for a in ints], _p(out))
    return out


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                interval_starts, interval_lengths):
    """bev_pool_v2 (bev_pool.py:86-92): the above + permute(0,4,1,2,3)."""
    out = bev_pool_v2_channels_last(depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                    bev_feat_shape, interval_starts, interval_lengths)
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))


def bp_intervals(ranks_depth, ranks_feat, ranks_bev):
    """QuickCumsumCuda.backward index work (bev_pool.py:47-57): sort the
    points by ranks_feat, head flags, starts / lengths."""
    order = np.argsort(ranks_feat, kind="stable")
    rf, rd, rb = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    head = np.ones(rf.shape[0], dtype=bool)
    head[1:] = rf[1:] != rf[:-1]
    starts = np.flatnonzero(head).astype(np.int32)
    lengths = np.empty_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rf.shape[0] - starts[-1]
    return rd, rf, rb, starts, lengths


def bev_pool_v2_backward(out_grad_bczyx, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """QuickCumsumCuda.backward (bev_pool.py:43-83) for a channels-first
    upstream gradient: `.contiguous()` to channels-last (:69), zeros (:67-68),
    kernel (bev_pool_cuda.cu:67-121).  Returns (depth_grad, feat_grad)."""
    depth = np.ascontiguousarray(depth, dtype=_f32)
    feat = np.ascontiguousarray(feat, dtype=_f32)
    g = np.ascontiguousarray(np.asarray(out_grad_bczyx, dtype=_f32).transpose(0, 2, 3, 4, 1))
    c = feat.shape[-1]
    rd, rf, rb, starts, lengths = bp_intervals(
        np.asarray(ranks_depth, np.int32), np.asarray(ranks_feat, np.int32),
        np.asarray(ranks_bev, np.int32))
    ints = [np.ascontiguousarray(a, dtype=np.int32) for a in (rd, rf, rb, starts, lengths)]
    depth_grad = np.zeros_like(depth)
    feat_grad = np.zeros_like(feat)
    _pool_lib().oracle_bev_pool_v2_grad(ctypes.c_int(c), ctypes.c_int(starts.shape[0]), _p(g),
                                        _p(depth), _p(feat), *[_p(a) for a in ints],
                                        _p(depth_grad), _p(feat_grad))
    return depth_grad, feat_grad


def bev_pool_v2_f64(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape):
    """Order-independent float64 evaluation of the same sums (tolerance
    reference): out[rank, c] = sum depth[rd] * feat[rf, c]; [B,C,Z,Y,X]."""
    B, Z, Y, X, C = bev_feat_shape
    out = np.zeros((B * Z * Y * X, C), dtype=np.float64)
    contrib = depth.reshape(-1).astype(np.float64)[ranks_depth][:, None] * \
        feat.reshape(-1, C).astype(np.float64)[ranks_feat]
    np.add.at(out, ranks_bev, contrib)
    return out.reshape(B, Z, Y, X, C).transpose(0, 4, 1, 2, 3)


# --------------------------------------------------------------------------
# a13-a15  open-vocabulary tail
# --------------------------------------------------------------------------
def class_groups(class_reflection):
    """Merged class id of every classifier row.  `class_reflection` has one
    entry per text prompt; the classifier has one extra trailing background
    row that always forms its own group
    (san_in_veon_entry_temporal.py:273-286: runs of equal class_reflection,
    `right < dim_length - 2` keeps the last row apart)."""
    refl = list(class_reflection)
    n = len(refl) + 1
    cls = np.empty(n, dtype=np.int32)
    left, k = 0, 0
    while left < n:
        right = left
        while right < n - 2 and refl[left] == refl[right + 1]:
            right += 1
        cls[left:right + 1] = k
        k += 1
        left = right + 1
    return cls


def voxel_text_labels(feat_occ, text_w, class_of_prompt, bin_occ, free_label=17,
                      dtype=np.float32):
    """feat_occ [B,C,Z,Y,X], text_w [Q,C], bin_occ [B,2,Z,Y,X] -> uint8
    labels [B,X,Y,Z].

      logits   einsum "qc,bczhw->bqzhw"            san_in_veon_temporal.py:257-259
      merge    max over each class's prompt rows   san_in_veon_entry_temporal.py:273-297
      label    argmax softmax == argmax logits (first index on ties); gate
               softmax(bin_occ)[:,0] > 0.5 and score > 0; free label;
               permute(0,3,2,1); uint8             veon_temporal.py:223-229,240
    """
    B, C, Z, Y, X = feat_occ.shape
    w = np.asarray(text_w, dtype=dtype)
    f = np.asarray(feat_occ, dtype=dtype).reshape(B, C, -1)
    logits = np.einsum("qc,bcv->bqv", w, f, optimize=True)
    cls = np.asarray(class_of_prompt)
    n_cls = int(cls.max()) + 1
    merged = np.full((B, n_cls, logits.shape[-1]), -np.inf, dtype=logits.dtype)
    for q in range(w.shape[0]):
        merged[:, cls[q]] = np.maximum(merged[:, cls[q]], logits[:, q])
    best = merged.argmax(axis=1)
    m = merged - merged.max(axis=1, keepdims=True)
    with np.errstate(invalid="ignore", over="ignore"):
        e = np.exp(m)
        score = (e / e.sum(axis=1, keepdims=True)).max(axis=1)
        b = np.asarray(bin_occ, dtype=np.float32).reshape(B, 2, -1)
        bm = b.max(axis=1, keepdims=True)
        be = np.exp(b - bm)
        p0 = be[:, 0] / be.sum(axis=1)
    sel = (score > 0.0) & (p0 > 0.5)
    lab = np.where(sel, best, free_label).reshape(B, Z, Y, X)
    return np.ascontiguousarray(lab.transpose(0, 3, 2, 1)).astype(np.uint8)


def trilinear_upsample(x, size):
    """F.interpolate(x, size=size, mode="trilinear", align_corners=False) for x [B,C,Zi,Yi,Xi]
    (san_in_veon_temporal.py:196-207), float32, restated from ATen's upsample_trilinear3d:
    src = max(scale*(dst+0.5)-0.5, 0) with scale = in/out, i0 = int(src), i1 = i0 + (i0 < in-1),
    l1 = src - i0, l0 = 1 - l1;  value = lz0*(ly0*(lx0*a + lx1*b) + ly1*(lx0*c + lx1*d)) + lz1*(...).
    Pinned against torch's own F.interpolate in tests/test_oracle_golden.py."""
    x = np.asarray(x, dtype=np.float32)

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = scale * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5)
        src = np.maximum(src, np.float32(0)).astype(np.float32)
        i0 = src.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1).astype(np.float32), l1

    Z, Y, X = (int(v) for v in size)
    z0, z1, lz0, lz1 = axis(x.shape[2], Z)
    y0, y1, ly0, ly1 = axis(x.shape[3], Y)
    x0, x1, lx0, lx1 = axis(x.shape[4], X)

    def plane(zi):
        p = x[:, :, zi]                                    # [B,C,Z,Yi,Xi]
        r0, r1 = p[:, :, :, y0], p[:, :, :, y1]            # [B,C,Z,Y,Xi]
        top = lx0 * r0[..., x0] + lx1 * r0[..., x1]
        bot = lx0 * r1[..., x0] + lx1 * r1[..., x1]
        return ly0[:, None] * top + ly1[:, None] * bot

    out = lz0[:, None, None] * plane(z0) + lz1[:, None, None] * plane(z1)
    return out.astype(np.float32)


def voxel_text_labels_lowres(feat_occ_lr, text_w, class_of_prompt, bin_occ_lr, occ_size,
                             free_label=17, dtype=np.float32):
    """The reference's order of operations: up-sample the decoder's feat_occ and bin_occ
    (san_in_veon_temporal.py:196-207), then classify the up-sampled volume."""
    feat = trilinear_upsample(feat_occ_lr, occ_size)
    gate = trilinear_upsample(bin_occ_lr, occ_size)
    return voxel_text_labels(feat, text_w, class_of_prompt, gate, free_label, dtype)


# --------------------------------------------------------------------------
# BASELINE.json configs[0]: the reference's pure-PyTorch CPU lift (timed leg)
# --------------------------------------------------------------------------
def torch_cpu_lift(coor, depth, feat, grid_lower_bound, grid_interval, grid_size,
                   out_grad=None):
    """prepare (torch CPU ops, same steps as view_transformer.py:202-260) +
    scatter-add pooling (`index_add_` restating bev_pool_cuda.cu:39-47) +
    autograd backward.  This is the CPU baseline `bench.py` times; torch uses
    every host thread it is given.  Returns (bev_feat, depth_grad, feat_grad).
    """
    import torch
    B, N, D, H, W, _ = coor.shape
    P = B * N * D * H * W
    lo = torch.as_tensor(grid_lower_bound, dtype=torch.float32)
    iv = torch.as_tensor(grid_interval, dtype=torch.float32)
    gs = torch.as_tensor(grid_size, dtype=torch.float32)
    vox = ((coor.reshape(P, 3) - lo) / iv).long()
    inside = ((vox >= 0) & (vox < gs)).all(dim=1)
    pid = torch.nonzero(inside).squeeze(1)
    vox = vox[pid]
    HW = H * W
    rf = (pid // (D * HW)) * HW + pid % HW
    rank = (pid // (P // B)).float() * (gs[2] * gs[1] * gs[0])
    rank = rank + vox[:, 2] * (gs[1] * gs[0])
    rank = rank + (vox[:, 1] * gs[0] + vox[:, 0])
    order = rank.argsort()
    rb, rd, rf = rank[order].long(), pid[order], rf[order]
    Z, Y, X = int(gs[2]), int(gs[1]), int(gs[0])
    C = feat.shape[2]
    depth = depth.detach().requires_grad_(out_grad is not None)
    feat = feat.detach().requires_grad_(out_grad is not None)
    rows = feat.permute(0, 1, 3, 4, 2).reshape(-1, C)
    vol = torch.zeros(B * Z * Y * X, C, dtype=rows.dtype).index_add_(
        0, rb, depth.reshape(-1)[rd].unsqueeze(1) * rows[rf])
    bev = vol.view(B, Z, Y, X, C).permute(0, 4, 1, 2, 3).contiguous()
    if out_grad is None:
        return bev, None, None
    bev.backward(out_grad)
    return bev.detach(), depth.grad, feat.grad


# ---------------------------------------------------------------- depth producer (8f-2)
def downsample_depth(depths, downsample):
    """LSSViewTransformerRaw.downsample_depth (view_transformer_raw.py:393-404): minimum over
    every downsample x downsample block, a zero (= no measurement) counting as 1e5."""
    d = np.asarray(depths, dtype=_f32)
    B, N, H, W = d.shape
    s = int(downsample)
    blk = d.reshape(B, N, H // s, s, W // s, s)
    blk = np.where(blk == 0.0, _f32(1e5), blk)
    return blk.min(axis=(3, 5)).astype(_f32)


def two_hot_depth(depths, depth_cfg, gamma=4, downsample=0):
    """LSSViewTransformerRaw.get_two_hot_depth (view_transformer_raw.py:406-429):
    softmax over D+1 bin centres of -|depth - centre| * gamma clamped at -16, last bin dropped;
    returns [B,N,D,H,W] float32.  `downsample` > 0 applies downsample_depth first (:413-414)."""
    d = np.asarray(depths, dtype=_f32)
    if downsample:
        d = downsample_depth(d, downsample)
    lo, hi, st = (float(v) for v in depth_cfg)
    D = int(np.arange(lo, hi, st, dtype=np.float32).shape[0])
    # torch.arange(D+1) * step + (lo + step/2): int64 * python float -> float32 tensor
    centers = np.arange(D + 1).astype(_f32) * _f32(st) + _f32(lo + st / 2)
    gap = -np.abs(d[..., None] - centers) * _f32(gamma)
    gap = np.where(gap >= _f32(-16), gap, _f32(-16)).astype(_f32)
    m = gap.max(axis=-1, keepdims=True)
    e = np.exp(gap - m, dtype=_f32)
    dist = (e / e.sum(axis=-1, keepdims=True, dtype=_f32)).astype(_f32)
    return np.ascontiguousarray(np.moveaxis(dist[..., :D], -1, 2))


def lidar_coor_torch(frustum, sensor2ego, ego2global, cam2imgs, post_rots, post_trans, bda):
    """`LSSViewTransformer.get_lidar_coor` (view_transformer.py:114-152) written with torch ops in
    the reference's operation order: the float reference the CUDA geometry kernels are validated
    against (tests only; `frustum` [D,H,W,3] from the neck)."""
    import torch
    fr = frustum.to(sensor2ego)                           # [D,H,W,3]
    undo_aug = torch.inverse(post_rots)                   # [B,N,3,3]
    cam2ego = sensor2ego[..., :3, :3] @ torch.inverse(cam2imgs)
    ego_t = sensor2ego[..., :3, 3]
    img = fr[None, None] - post_trans[:, :, None, None, None, :]
    img = torch.einsum("bnij,bndhwj->bndhwi", undo_aug, img)
    depth = img[..., 2:3]
    cam = torch.cat((img[..., :2] * depth, depth), dim=-1)   # pixel * depth, depth
    ego = torch.einsum("bnij,bndhwj->bndhwi", cam2ego, cam)
    ego = ego + ego_t[:, :, None, None, None, :]
    return torch.einsum("bij,bndhwj->bndhwi", bda, ego)
