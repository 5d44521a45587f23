"""N>1 host logic on CPU: world_size-2 gloo process group, 127.0.0.1 rendezvous.  The training step shards by batch
with ONE collective (mean all-reduce of the flat gradient buffer); the sampler shards by image count with none."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddpm_torch_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(10_000, generator=g)
    mine = flat.clone()
    parallel.allreduce_mean_(flat)
    # every rank holds the mean of all ranks' buffers
    ref = sum(torch.randn(10_000, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ok = torch.allclose(flat, ref, atol=1e-6) and not torch.equal(flat, mine)
    # batch / image sharding
    ok = ok and parallel.per_rank_batch(128, world) == 64
    ok = ok and sum(parallel.shard_size(1001, r, world) for r in range(world)) == 1001
    ok = ok and parallel.rank_seeds(rank) == (8191 + rank, 131071 + rank)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_mean_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world))


def test_shard_size_matches_reference_rule():
    # generate.py:105-110: total // world, +1 for the first `total % world` ranks
    assert [parallel.shard_size(10, r, 4) for r in range(4)] == [3, 3, 2, 2]
    assert [parallel.shard_size(8, r, 8) for r in range(8)] == [1] * 8
    assert parallel.allreduce_mean_(torch.ones(3)).tolist() == [1.0, 1.0, 1.0]      # no process group: identity
