"""One CIFAR bs=128 training step + one fused optimizer step between cudaProfilerStart/Stop (for `ncu --profile-from-start off -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = [sys.argv[0], "train", "128"]
here = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(here, "profile_step.py")).read().split("for i in range(2):")[0]
exec(src)
from ddpm_torch_b200.optim import EMA, FusedAdam
opt = FusedAdam(model, lr=2e-4, betas=(0.9, 0.999), warmup=5000, grad_norm=1.0, ema=EMA(model, 0.9999))
for i in range(2):
    step(i); opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(2); opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
