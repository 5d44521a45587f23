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


def grad_chunks(model):
    """[(lo, hi)] element ranges of the model's flat gradient buffer in the order the backward pass completes them."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    lo, hi = (C.c_longlong * 64)(), (C.c_longlong * 64)()
    n = L.ddpm_unet_grad_chunks(model._h, 64, lo, hi)
    if n < 0:
        _lib.check(n, "grad_chunks")
    return [(int(lo[i]), int(hi[i])) for i in range(n)]


def allreduce_grads_overlapped_(model, group=None):
    """Mean all-reduce of the flat gradient buffer, chunk by chunk, OVERLAPPED with the tail of the backward pass: call it right
    after the (asynchronous) backward launch.  Every chunk is reduced on a communication stream that waits only for that chunk's
    completion event inside the engine's backward (ddpm_unet_wait_grad_chunk); the current stream finally waits for the
    communication stream.  This is what DDP's bucketed reducer does for the reference (train.py:110), with the engine's own
    level-group chunks as buckets and the flat buffer as the message."""
    from . import _lib
    flat = model._grads
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or flat is None:
        return flat
    if model._acc_pending:                      # gradient accumulation in flight: reduce the folded buffer in one piece
        return allreduce_mean_(model.flat_grads, group)
    L = _lib.lib()
    dev = flat.device
    chunks = getattr(model, "_grad_chunks", None)
    if chunks is None or model._grad_chunks_key != model._plan_key:
        chunks = model._grad_chunks = grad_chunks(model)
        model._grad_chunks_key = model._plan_key
        assert sum(h - l for l, h in chunks) == flat.numel() and min(l for l, _ in chunks) == 0
    comm = getattr(model, "_comm_stream", None)
    if comm is None or comm.device != dev:
        comm = model._comm_stream = torch.cuda.Stream(device=dev)
    with torch.cuda.device(dev):
        for i, (lo, hi) in enumerate(chunks):
            _lib.check(L.ddpm_unet_wait_grad_chunk(model._h, i, comm.cuda_stream), "wait_grad_chunk")
            with torch.cuda.stream(comm):
                allreduce_mean_(flat[lo:hi], group)
        torch.cuda.current_stream(dev).wait_stream(comm)
    return flat
