"""Data parallelism: one process per GPU (torchrun), NCCL over NVLink for the ONE collective the path needs -- the
SUM all-reduce of the flat f32 gradient buffer (7 158 592 floats, 28.6 MB) -- plus a parameter broadcast at start.
The global-norm clip is computed after the reduction (crnn_clip_adam_step with grad_mul = 1/world, wd_mul = world).
The forward path has no collective: BatchNorm uses per-replica batch statistics."""
import numpy as np
import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def allreduce_sum_(flat):
    """In-place SUM all-reduce of a flat tensor (no-op for world size 1)."""
    if world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def broadcast_(flat, src=0):
    if world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def shard_batch(data, labels, label_len, time_step_len, rank_, world):
    """Contiguous sample shard of one global batch (data-layer tuple, gen.py:67): returns this rank's
    (data, flat_labels, label_len, time_step_len).  N must be divisible by world."""
    data = np.asarray(data); labels = np.asarray(labels); label_len = np.asarray(label_len); time_step_len = np.asarray(time_step_len)
    N = data.shape[0]
    if N % world != 0:
        raise ValueError(f"global batch {N} is not divisible by world size {world}")
    per = N // world
    lo, hi = rank_ * per, (rank_ + 1) * per
    offs = np.concatenate([[0], np.cumsum(label_len)])
    return data[lo:hi], labels[offs[lo]:offs[hi]], label_len[lo:hi], time_step_len[lo:hi]
