#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 300 2>&1 | tail -3
timeout 600 python - <<PY
import os, sys, json
sys.path.insert(0, ".")
import torch, bench
import ddpm_torch_b200 as D
dev = torch.device("cuda")
model = D.UNet(**bench.CIFAR).to(dev).eval()
with torch.no_grad():
    gi = torch.Generator(device=dev).manual_seed(7)
    for n_, p in model.named_parameters():
        if p.ndim >= 2:
            p.copy_((torch.rand(p.shape, device=dev, generator=gi) * 2 - 1) * (3.0 / p[0].numel()) ** 0.5)
print(json.dumps(bench.bench_sampler(D, model, dev)))
PY
