"""Sample sharding over the GPUs of one box + the one collective of the path.

The lift and the tail are independent per sample (the batch index is only the
top digit of ranks_bev, view_transformer.py:241), so samples are dealt
round-robin to ranks and nothing is exchanged until the finished uint8
occupancy volumes are all-gathered -- the B200 counterpart of mmdet's
`multi_gpu_test` result collection used by the reference (tools/test.py:247).
Feature volumes never cross NVLink.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_samples", "all_gather_occupancy", "shard_cameras", "all_reduce_logit_volume",
           "lift_classify_camera_sharded", "reserve_sms", "gathered_row"]


def reserve_sms(n):
    """Leave `n` SMs free of the path's persistent grids (`veon_reserve_sms`) so that a collective
    running on another stream finds room for its CTAs at once instead of at the next kernel
    boundary; returns the previous value.  0 = use every SM (the default)."""
    from . import _lib
    return int(_lib.load().veon_reserve_sms(int(n)))


def shard_samples(n_samples, world_size, rank):
    """Sample i -> rank i mod world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_samples, world_size))


def all_gather_occupancy(local_labels, n_samples, group=None, sample_order=True):
    """local_labels: uint8 [B_local, X, Y, Z] for samples shard_samples(n_samples, G, r)
    (in that order).  Returns uint8 [n_samples, X, Y, Z] identical on every rank.
    One all_gather_into_tensor; shards are padded to the largest B_local.

    sample_order=False: skip the device-side reordering pass and return the gathered buffer as it
    arrives, [G * per, X, Y, Z] rank-major (sample g + j*G at row g*per + j, `gathered_row`): the
    reference's collector puts its results into dataset order on the HOST as well
    (mmdet `collect_results_gpu`), and the pass re-reads and re-writes the whole gathered buffer
    in HBM while the next step's kernels run."""
    if local_labels.dtype != torch.uint8 or local_labels.dim() != 4:
        raise ValueError("local_labels must be uint8 [B_local, X, Y, Z]")
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local_labels.shape[0] != n_samples:
            raise ValueError("single process must hold every sample")
        return local_labels
    G, r = dist.get_world_size(group), dist.get_rank(group)
    mine = shard_samples(n_samples, G, r)
    if local_labels.shape[0] != len(mine):
        raise ValueError(f"rank {r} holds {local_labels.shape[0]} samples, expected {len(mine)}")
    per = (n_samples + G - 1) // G
    vol = tuple(local_labels.shape[1:])
    send = local_labels
    if len(mine) < per:
        send = torch.zeros((per,) + vol, dtype=torch.uint8, device=local_labels.device)
        send[:len(mine)] = local_labels
    send = send.contiguous()
    recv = torch.empty((G * per,) + vol, dtype=torch.uint8, device=local_labels.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if not sample_order:
        return recv
    # recv[g*per + j] is sample g + j*G  ->  put back in sample order
    recv = recv.view(G, per, *vol).transpose(0, 1).reshape(G * per, *vol)
    return recv[:n_samples].contiguous()


def gathered_row(sample, n_samples, world_size):
    """Row of `sample` in the rank-major buffer all_gather_occupancy(sample_order=False) returns."""
    per = (n_samples + world_size - 1) // world_size
    return (sample % world_size) * per + sample // world_size


# ---- camera-group sharding (fewer samples than GPUs) ---------------------------------
def shard_cameras(n_cams, world_size, rank):
    """Camera c -> rank c mod world_size.  Within a sample the batch digit is the only coupling
    between points (view_transformer.py:241) and pooling is a plain sum over points, so disjoint
    camera groups can be lifted on different GPUs and their volumes added."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_cams, world_size))


def all_reduce_logit_volume(vol, group=None):
    """Sum the per-rank [B, Q', Z, Y, X] logit volumes in place (the path's one real exchange
    step in this mode: 51 MB per sample at Q' = 20 instead of the 1.31 GB C = 512 feature
    volume).  A no-op for a single process."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vol, op=dist.ReduceOp.SUM, group=group)
    return vol


def lift_classify_camera_sharded(neck, input, depth, tran_feat, ov_classifier_weight, prompt_class,
                                 gate_weight, free_label=17, group=None, channel_pad=4):
    """`pipeline.lift_classify` with the CAMERAS of every sample dealt over the ranks: each rank
    classifies and lifts the pixels of its cameras into a logit volume, one NCCL all-reduce sums
    the volumes, every rank finishes with the same labels (uint8 [B,X,Y,Z]).  All ranks pass the
    same full inputs.  Labels equal the unsharded result up to float re-association of the
    per-voxel sums (>= 99.99 % of voxels; tests/test_tail_gpu.py)."""
    from . import tail as _tail
    from .pipeline import lift_logits
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    r = dist.get_rank(group) if dist.is_initialized() else 0
    n_cams = input[0].shape[1]
    mine = shard_cameras(n_cams, G, r)
    Q = ov_classifier_weight.shape[0]
    if mine:
        vol = lift_logits(neck, input, depth, tran_feat, ov_classifier_weight, gate_weight,
                          channel_pad, cameras=mine)
    else:       # more ranks than cameras: contribute zeros
        Qp = (Q + 2 + channel_pad - 1) // channel_pad * channel_pad
        B, Z, Y, X, _ = neck._bev_shape(input[0], Qp)
        vol = torch.zeros((B, Qp, Z, Y, X), dtype=torch.float32, device=tran_feat.device)
    all_reduce_logit_volume(vol, group)
    with torch.no_grad():
        return _tail.classify_logits(vol[:, :Q], vol[:, Q:Q + 2], prompt_class, free_label)
