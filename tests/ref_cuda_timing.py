"""Time the reference's STOCK torch-CUDA path on this box's B200 (the denominator of the north-star's
">=10x the reference's stock torch-CUDA UNet step").  /root/reference cannot travel to the GPU box, so this
runs the oracle restatement — the same ATen/cuDNN/cuBLAS calls in the same order — with the reference's
settings (fp32 params, TF32 convs on, cudnn.benchmark=True as train.py:227-228).  Informational only.

    python tests/ref_cuda_timing.py [--bs 128] [--out gpurun_out/ref_cuda.json]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ddpm_ref as R  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=128)
    ap.add_argument("--sample-bs", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/ref_cuda.json")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    cfg = dict(R.CIFAR10_CFG); cfg["drop_rate"] = 0.0
    sd = {k: v.to(dev).requires_grad_(True) for k, v in R.make_state_dict(cfg, 1234).items()}
    diff = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-large")
    g = torch.Generator(device=dev).manual_seed(0)
    B = args.bs
    x0 = torch.randn(B, 3, 32, 32, device=dev, generator=g)
    t = torch.randint(1000, (B,), device=dev, generator=g)
    noise = torch.randn(B, 3, 32, 32, device=dev, generator=g)
    res = {"gpu": torch.cuda.get_device_name(0), "bs": B}

    def step(autocast):
        for p in sd.values():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = diff.train_losses(lambda x, tt: R.unet_forward(sd, cfg, x, tt), x0, t, noise).mean()
        loss.backward()

    for name, tf32, ac in (("tf32_default", True, False), ("fp32_no_tf32", False, False), ("bf16_autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        ms = timeit(lambda: step(ac))
        res[f"train_step_ms_{name}"] = ms
        res[f"train_img_s_{name}"] = B / ms * 1e3
        print(name, f"{ms:.2f} ms  {B / ms * 1e3:.0f} img/s", flush=True)
    torch.backends.cudnn.allow_tf32 = True

    # sampler: one ancestral step (UNet forward + the 41-op tail) at bs=256
    Bs = args.sample_bs
    sdn = {k: v.detach() for k, v in sd.items()}
    xs = torch.randn(Bs, 3, 32, 32, device=dev, generator=g)
    tt = torch.full((Bs,), 500, device=dev, dtype=torch.int64)

    def sstep(ac):
        with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            z = torch.empty_like(xs).normal_()
            return diff.p_sample_step(lambda x, q: R.unet_forward(sdn, cfg, x, q).float(), xs, tt, z)

    for name, ac in (("tf32_default", False), ("bf16_autocast", True)):
        ms = timeit(lambda: sstep(ac))
        res[f"sampler_step_ms_{name}"] = ms
        res[f"sampler_T1000_img_s_{name}"] = Bs / (ms * 1000) * 1e3
        res[f"sampler_ddim50_img_s_{name}"] = Bs / (ms * 50) * 1e3
        print("sampler", name, f"{ms:.2f} ms/step", flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
