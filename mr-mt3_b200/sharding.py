"""Track sharding across the GPUs of one box, and the gradient exchange of the fine-tune step
(SURVEY 8e).

Tracks are independent (reference test.py:45-64 walks them one by one); the segments of one
MR-MT3 track are sequential (models/t5_segmem_v2_with_prev.py:241-294).  So the unit of
partitioning is the track, there is no collective on the data path, and the only exchange is a
final gather of the int token rows.  The reference has no counterpart: it runs one process on
one GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_tracks(seg_counts, world_size):
    """Longest-processing-time assignment of tracks to ranks by segment count.
    -> list (per rank) of track-index lists; deterministic (ties by track index)."""
    seg_counts = np.asarray(seg_counts)
    order = sorted(range(len(seg_counts)), key=lambda i: (-int(seg_counts[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(seg_counts[i])
    return [sorted(s) for s in shards]


def gather_token_rows(local_rows, local_track_ids, seg_counts, max_length, group=None, dst=0):
    """Gather every rank's (n_local_segments, max_length) int64 token rows on `dst` and put them
    back in global track order.  Works with the nccl (CUDA tensors) and gloo (CPU) backends.
    Returns (total_segments, max_length) on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    seg_counts = np.asarray(seg_counts)
    shards = [None] * world
    dist.all_gather_object(shards, list(local_track_ids), group=group)
    n_rows = [int(sum(seg_counts[t] for t in s)) for s in shards]
    cap = max(n_rows) if n_rows else 0
    dev = local_rows.device
    padded = torch.zeros((cap, max_length), dtype=torch.int64, device=dev)
    padded[:local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    base = np.concatenate([[0], np.cumsum(seg_counts)])
    out = torch.zeros((int(base[-1]), max_length), dtype=torch.int64, device=dev)
    for r, tracks in enumerate(shards):
        off = 0
        for t in tracks:
            n = int(seg_counts[t])
            out[int(base[t]):int(base[t]) + n] = bufs[r][off:off + n]
            off += n
    return out


def allreduce_mean_(flat_grad, group=None):
    """Data-parallel gradient exchange of the fine-tune step: ONE all-reduce (sum) of the flat fp32
    gradient buffer, divided by the world size in place -- what DDP's bucketed all-reduce amounts
    to for the reference (every parameter receives a gradient), without buckets because
    the whole gradient already is one contiguous 194 MB buffer.  nccl (CUDA) or gloo (CPU).
    A no-op when torch.distributed is not initialised or the group has one rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad /= world
    return flat_grad
