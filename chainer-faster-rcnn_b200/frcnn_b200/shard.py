"""Multi-GPU plumbing for the inference path: images shard one per GPU, replicas only.

The reference hard-wires batch size 1 (models/faster_rcnn.py:77) and its forward path has no
cross-image state, so N GPUs run N independent replicas and there is NO data-path collective
(SURVEY.md 8e).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
barrier around the timed region, the max-over-ranks of the elapsed time and gathering per-rank counts.

The one real collective of the repository is on the TRAINING path (train_rpn.py:169-174, Chainer ParallelUpdater:
every device back-propagates its own image, gradients are ADDED into the main model, one optimizer update): here a
single all-reduce(SUM) of the flat float32 gradient bucket, `allreduce_sum_`, followed by the identical local update
on every rank (weights stay replicated without a broadcast).
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items, rank, world):
    """Image i -> rank i mod world (the partition SURVEY.md 8e names).  Returns the list for `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_items, world))


def image_seed(rank, step, base=0):
    """Distinct synthetic image per (rank, step): no two replicas ever process the same pixels."""
    return base + 1000 * rank + step


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ints(value, device=None):
    """Per-rank integer (e.g. proposals kept) gathered to every rank, ordered by rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(value)]
    t = torch.tensor([int(value)], dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]


def aggregate_throughput(items_per_rank, seconds_this_rank, device=None):
    """Whole-job throughput: all ranks' items divided by the SLOWEST rank's time."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    return world * items_per_rank / max_over_ranks(seconds_this_rank, device)


def allreduce_sum_(flat, group=None):
    """In-place SUM over ranks of one flat gradient bucket (no-op for a single process).  Returns `flat`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
