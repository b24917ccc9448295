"""Lift + classify as one pipeline (BASELINE.json configs[2]).

When the classifier is applied to the pooled volume directly (north_star's "2D-to-3D lifting
followed by the open-vocabulary voxel-text classification tail"), both steps are linear in the
image features:

    sem_occ[b,q,v] = sum_c W[q,c] * sum_{p in v} depth[p] * feat[pixel(p), c]
                   = sum_{p in v} depth[p] * (sum_c W[q,c] * feat[pixel(p), c])

so the classifier can run on the [B,N,C,H,W] image features (17 k pixels per sample instead of
640 k voxels) and the pooling then moves Q + 2 channels instead of C = 512: the 1.31 GB/sample
feature volume is never formed.  `lift_classify` does that with the same operators as the
separate steps (`semantic_inference_3d` on tcgen05, `view_transform`'s index preparation and
pooling kernels, `classify_logits`); `lift_then_classify` is the feature-space order of the
reference (`view_transform` -> `voxel_text_argmax`) kept for comparison and for the tests.

The occupancy gate is a linear head here (`gate_weight [2,C]`, bin_occ = gate_weight . volume):
the reference's gate comes out of its 3D decoder, which sits between the two steps in the full
model and is out of scope; with a non-linear decoder in between only
`voxel_text_argmax_lowres` applies.
"""
import torch

from . import tail as _tail

__all__ = ["lift_classify", "lift_logits", "lift_then_classify"]


def _heads(ov_classifier_weight, gate_weight, channel_pad=4):
    w = ov_classifier_weight.detach().float()
    g = gate_weight.detach().float()
    if g.shape != (2, w.shape[1]):
        raise ValueError("gate_weight must be [2, C]")
    rows = torch.cat((w, g), 0)
    pad = (-rows.shape[0]) % channel_pad   # the pooling kernels stage rows with 16-byte copies
    if pad:
        rows = torch.cat((rows, rows.new_zeros(pad, rows.shape[1])), 0)
    return rows.contiguous()


def lift_logits(neck, input, depth, tran_feat, ov_classifier_weight, gate_weight, channel_pad=4,
                cameras=None):
    """The pooled LOGIT volume [B, Q', Z, Y, X] (Q' = Q + 2 gate channels, padded to a multiple of
    `channel_pad`): classifier on the image features, then the lift of those Q' channels.

    `cameras`: optional list of camera indices -- only their pixels are lifted.  Pooling is a sum
    over points, so the volumes of disjoint camera groups ADD UP to the full volume (up to float
    re-association): that is what camera-group sharding all-reduces (veon_b200.dist)."""
    B, N = input[0].shape[:2]
    BN, C, H, W = tran_feat.shape
    rows = _heads(ov_classifier_weight, gate_weight, channel_pad)
    with torch.no_grad():
        metas = list(input[1:7])
        d5 = depth.reshape(B, N, neck.D, H, W)
        f5 = tran_feat.reshape(B, N, C, H, W)
        if cameras is not None:
            idx = torch.as_tensor(list(cameras), device=tran_feat.device, dtype=torch.long)
            # every calibration tensor but bda [B,3,3] carries the camera axis
            metas = [m.index_select(1, idx.to(m.device)).contiguous() if i < 5 else m
                     for i, m in enumerate(metas)]
            d5 = d5.index_select(1, idx).contiguous()
            f5 = f5.index_select(1, idx).contiguous()
            N = len(cameras)
        # per-pixel logits: the classifier over each camera's feature map, [B*N, Q', 1, H, W]
        px = _tail.semantic_inference_3d(rows, f5.reshape(B * N, C, 1, H, W))
        px5 = px.view(B, N, rows.shape[0], H, W)
        if metas[0].is_cuda and neck.fuse_geometry:
            vol = neck._voxel_pooling_calib(metas, d5, px5)
        else:
            vol = neck.voxel_pooling_v2(neck.get_lidar_coor(*metas), d5, px5)
        if vol.dim() != 5:
            raise RuntimeError("lift_logits needs collapse_z=False and points inside the grid")
        return vol


def _heads_gate_first(ov_classifier_weight, gate_weight):
    """[gate 0, gate 1, prompt rows, zero padding to a multiple of 4]: the row order of the fused
    lift + classify kernel (the gate channels sit at fixed register positions)."""
    w = ov_classifier_weight.detach().float()
    g = gate_weight.detach().float()
    if g.shape != (2, w.shape[1]):
        raise ValueError("gate_weight must be [2, C]")
    rows = torch.cat((g, w), 0)
    # the kernel walks the row in 1, 2 or 3 passes of equal width (<= 32 channels each)
    n = rows.shape[0]
    step = 4 if n <= 32 else (8 if n <= 64 else 12)
    pad = max(-n % step, (40 if step == 8 else 72 if step == 12 else 0) - n)
    if pad:
        rows = torch.cat((rows, rows.new_zeros(pad, rows.shape[1])), 0)
    return rows.contiguous()


def lift_classify(neck, input, depth, tran_feat, ov_classifier_weight, prompt_class, gate_weight,
                  free_label=17, channel_pad=4, fused=True):
    """neck: LSSViewTransformer; input = (img [B,N,*,H,W], sensor2ego, ego2global, cam2imgs,
    post_rots, post_trans, bda); depth [B*N,D,H,W]; tran_feat [B*N,C,H,W];
    ov_classifier_weight [Q,C]; prompt_class [Q]; gate_weight [2,C] -> uint8 [B,X,Y,Z].

    fused=True (default): the pooling kernel classifies each tile itself and the pooled logit
    volume is never written (`veon_lift_classify_fwd`, Q <= 94 prompt rows); otherwise, or when
    that kernel does not take the shape, the volume is pooled and `classify_logits` reads it.
    Both give the same labels (the same sums, bit for bit, through the same rule)."""
    Q = ov_classifier_weight.shape[0]
    if (fused and Q <= 94 and tran_feat.is_cuda and neck.fuse_geometry
            and not neck.collapse_z):
        B, N = input[0].shape[:2]
        BN, C, H, W = tran_feat.shape
        with torch.no_grad():
            rows = _heads_gate_first(ov_classifier_weight, gate_weight)
            px = _tail.semantic_inference_3d(rows, tran_feat.reshape(B * N, C, 1, H, W))
            labels = neck.lift_labels_calib(list(input[1:7]), depth.reshape(B, N, neck.D, H, W),
                                            px.view(B, N, rows.shape[0], H, W), Q, prompt_class,
                                            free_label)
        if labels is not None:
            return labels
    vol = lift_logits(neck, input, depth, tran_feat, ov_classifier_weight, gate_weight, channel_pad)
    with torch.no_grad():
        return _tail.classify_logits(vol[:, :Q], vol[:, Q:Q + 2], prompt_class, free_label)


def lift_then_classify(neck, input, depth, tran_feat, ov_classifier_weight, prompt_class,
                       gate_weight, free_label=17):
    """The same result in the reference's order: pool the C-channel features, then classify the
    volume (einsum for the gate head: it is a stand-in for the decoder, not a path kernel)."""
    with torch.no_grad():
        vol, _ = neck.view_transform(input, depth, tran_feat)
        gate = torch.einsum("kc,bczyx->bkzyx", gate_weight.float(), vol)
        return _tail.voxel_text_argmax(vol, ov_classifier_weight, prompt_class, gate, free_label)
