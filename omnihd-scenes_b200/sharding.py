"""Frame-batch sharding across GPUs.

Frames are independent — `ranks_bev` carries the batch index as its most significant term
(cam_stream_lss_bevpoolv2.py:332-333) — so rank r of G takes frames [r*B/G, (r+1)*B/G) and runs
geometry + prepare + forward + backward locally. There is NO collective on the path.
The only exchange is optional: an all-gather of the per-rank BEV-grid shards for consumers that
want the whole batch on every GPU (the high-resolution occupancy grid, BASELINE config 5),
done with `all_gather_into_tensor` (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def frame_shard(batch, rank, world_size):
    """[lo, hi) frame range of `rank`; the first `batch % world_size` ranks take one extra frame."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_frames(tensors, rank, world_size):
    """Slice dim 0 (frames) of every tensor for this rank."""
    B = tensors[0].shape[0]
    lo, hi = frame_shard(B, rank, world_size)
    return tuple(t[lo:hi] for t in tensors)


def all_gather_bev(bev_shard, group=None, async_op=False):
    """All-gather [B/G, C, Z, Y, X] shards into [B, C, Z, Y, X] (equal shard sizes).
    Returns (full, work) — `work` is None unless async_op, so the gather can overlap the next
    step's pooling on another stream."""
    world = dist.get_world_size(group)
    bev_shard = bev_shard.contiguous()
    full = bev_shard.new_empty((bev_shard.shape[0] * world,) + tuple(bev_shard.shape[1:]))
    work = dist.all_gather_into_tensor(full, bev_shard, group=group, async_op=async_op)
    return full, (work if async_op else None)
