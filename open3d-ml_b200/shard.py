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


def gather_frame_results(local, num_items, group=None):
    """local: [n_local, ...] tensor of per-frame results of this rank's shard.  Returns the
    [num_items, ...] tensor of all frames on every rank (pads ragged tails internally)."""
    world = dist.get_world_size(group)
    per = (num_items + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out, 0)[:num_items]


def reduce_confusion(conf, group=None):
    """Sum of per-rank [C, C] int64 confusion matrices (SemSegMetric, semseg_metric.py:95-126)."""
    dist.all_reduce(conf, op=dist.ReduceOp.SUM, group=group)
    return conf
