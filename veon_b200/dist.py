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

__all__ = ["shard_samples", "all_gather_occupancy"]


def shard_samples(n_samples, world_size, rank):
    """Sample i -> rank i mod world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_samples, world_size))


def all_gather_occupancy(local_labels, n_samples, group=None):
    """local_labels: uint8 [B_local, X, Y, Z] for samples shard_samples(n_samples, G, r)
    (in that order).  Returns uint8 [n_samples, X, Y, Z] identical on every rank.
    One all_gather_into_tensor; shards are padded to the largest B_local."""
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
    # recv[g*per + j] is sample g + j*G  ->  put back in sample order
    recv = recv.view(G, per, *vol).transpose(0, 1).reshape(G * per, *vol)
    return recv[:n_samples].contiguous()
