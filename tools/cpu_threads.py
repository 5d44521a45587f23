import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ddpm_ref as R
cfg = dict(R.CIFAR10_CFG); cfg["drop_rate"] = 0.0
sd = {k: v.requires_grad_(True) for k, v in R.make_state_dict(cfg, 1234).items()}
diff = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-large")
for bs in (8, 32):
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(bs, 3, 32, 32, generator=g); t = torch.randint(1000, (bs,), generator=g); noise = torch.randn(bs, 3, 32, 32, generator=g)
    for nt in (8, 16, 32, 64, 128):
        torch.set_num_threads(nt)
        def step():
            for p in sd.values(): p.grad = None
            diff.train_losses(lambda x, tt: R.unet_forward(sd, cfg, x, tt), x0, t, noise).mean().backward()
        step()
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < 4 and n < 10: step(); n += 1
        dt = (time.perf_counter() - t0) / n
        print(f"bs {bs} threads {nt}: {dt*1e3:.0f} ms/step {bs/dt:.1f} img/s", flush=True)
