"""the reference flow wraps the model in DistributedDataParallel (train.py:110) and calls
``diffusion.train_losses(self.model, ...)`` with the WRAPPER (utils/train.py:144).  The fused path calls the inner UNet
directly, so DDP.forward never runs and its reducer never fires; the engine must average the flat gradient itself.
Two ranks share cuda:0 over a gloo group (NCCL refuses two ranks on one device; gloo all-reduces CUDA tensors)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import ddpm_torch_b200 as D
    from oracle import ddpm_ref as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cfg = R.SMALL64_CFG
    c = R.normalize_cfg(cfg)
    inner = D.UNet(3, c["hid_channels"], 3, c["ch_multipliers"], c["num_res_blocks"], c["apply_attn"], drop_rate=0.0)
    inner.load_state_dict(R.make_state_dict(cfg, 12))
    inner = inner.to(dev).train()
    model = DDP(inner, device_ids=[0])
    diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-large", "mse")
    g = torch.Generator(dev).manual_seed(100 + rank)              # different data per rank
    x0 = torch.randn(4, 3, 32, 32, device=dev, generator=g); t = torch.randint(1000, (4,), device=dev, generator=g)
    nz = torch.randn(4, 3, 32, 32, device=dev, generator=g)
    # un-reduced local gradient through the bare model
    inner.zero_grad()
    diff.train_losses(inner, x0, t, nz).mean().backward()
    local = torch.cat([p.grad.flatten() for p in inner.parameters()]).clone()
    # reference flow: the DDP wrapper is the denoise_fn
    inner.zero_grad()
    diff.train_losses(model, x0, t, nz).mean().backward()
    red = torch.cat([p.grad.flatten() for p in inner.parameters()]).clone()
    locs = [torch.empty_like(local) for _ in range(world)]; reds = [torch.empty_like(red) for _ in range(world)]
    dist.all_gather(locs, local); dist.all_gather(reds, red)
    mean = sum(locs) / world
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    ret[rank] = dict(equal=bool(torch.equal(reds[0], reds[1])), vs_mean=rel(red, mean), vs_local=rel(red, local),
                     local_diff=rel(locs[0], locs[1]))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_wrapped_train_losses_averages_gradients():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(world):
        d = ret[r]
        print(f"\n[ddp rank {r}] {d}")
        assert d["equal"], "replicas hold different gradients"
        assert d["vs_mean"] < 2e-2              # run-to-run reduction-order noise of two engine passes
        assert d["local_diff"] > 0.1 and d["vs_local"] > 0.05
