"""Frame sharding for multi-GPU inference.

Frames / clouds are independent units: rank r takes a contiguous chunk of the
batch (like ObjectDetectBatch.scatter, ml3d/torch/dataloaders/concat_batcher.py:538-553),
weights are replicated, and the forward needs NO collective.  The only exchange is
after the batch: an all_gather of fixed-size per-frame results (label maps) or an
all_reduce of the confusion matrix for whole-batch metrics -- KB..MB messages over
NCCL/NVLink (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(num_items, rank, world_size):
    """Contiguous [lo, hi) chunk of `num_items` for `rank` (ceil-sized chunks, last may be short)."""
    per = (num_items + world_size - 1) // world_size
    lo = min(rank * per, num_items)
    return lo, min(lo + per, num_items)


_GATHER_BUFFERS = {}


def gather_frame_results(local, num_items, group=None):
    """local: [n_local, ...] tensor of per-frame results of this rank's shard.  Returns the
    [num_items, ...] tensor of all frames on every rank (ragged tails are zero-padded internally).
    One all_gather_into_tensor on persistent buffers (cached per shape): nothing is allocated or freed
    around the collective after the first call, and the result is a view of the cached buffer -- valid until
    the next call with the same shape."""
    world = dist.get_world_size(group)
    per = (num_items + world - 1) // world
    key = (tuple(local.shape[1:]), local.dtype, str(local.device), per, world, id(group))
    buf = _GATHER_BUFFERS.get(key)
    if buf is None:
        if len(_GATHER_BUFFERS) > 16:
            _GATHER_BUFFERS.clear()
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf = _GATHER_BUFFERS[key] = (pad, out)
    pad, out = buf
    n = local.shape[0]
    pad[:n].copy_(local)
    if n < per:
        pad[n:].zero_()
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:num_items]


def reduce_confusion(conf, group=None):
    """Sum of per-rank [C, C] int64 confusion matrices (SemSegMetric, semseg_metric.py:95-126)."""
    dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)
    return conf
