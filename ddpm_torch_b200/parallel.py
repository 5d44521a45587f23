"""Data-parallel host logic (one process per GPU), mirroring what the reference does with DDP / mp.spawn:

* training (train.py:110 DDP, utils/train.py:149-153): every rank runs the same step on ``batch_size // world_size``
  samples; gradients are AVERAGED across ranks.  Here that is ONE all-reduce of the engine's flat fp32 gradient buffer
  (NCCL over NVLink on GPUs, gloo in the CPU tests) instead of DDP's 25 MB buckets over 304 tensors.
* sampling (generate.py:105-110,168-172): images are split across ranks, ``total // world`` each and one extra for the
  first ``total % world`` ranks; no collective.
"""
import torch
import torch.distributed as dist


def shard_size(total: int, rank: int, world: int) -> int:
    """generate.py:105-110."""
    return total // world + (1 if rank < total % world else 0)


def per_rank_batch(batch_size: int, world: int) -> int:
    """datasets.py:244-245: the global batch is divided evenly; remainders are dropped exactly like the reference."""
    return batch_size // world


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean all-reduce of a flat gradient buffer (the one collective of the training step)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if flat.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:                                   # gloo has no AVG (CPU tensors, or CUDA tensors over a gloo group)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def rank_seeds(rank: int):
    """utils/train.py:115-117: per-rank generator seeds for (t, noise) and for preview sampling."""
    return 8191 + rank, 131071 + rank
