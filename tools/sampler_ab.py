"""ancestral sampler ms/step at bs=256 (graph replay), for A/B runs under env knobs"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import ddpm_torch_b200 as D
dev = torch.device("cuda")
model = D.UNet(**bench.CIFAR).to(dev).eval()
with torch.no_grad():
    gi = torch.Generator(device=dev).manual_seed(7)
    for n_, p in model.named_parameters():
        if p.ndim >= 2:
            p.copy_((torch.rand(p.shape, device=dev, generator=gi) * 2 - 1) * (3.0 / p[0].numel()) ** 0.5)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 300
diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, S), "eps", "fixed-large", "mse")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = diff.p_sample(model, shape=(256, 3, 32, 32), device=dev, seed=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("ms/step %.4f" % (dt / S * 1e3), {k: v for k, v in os.environ.items() if k.startswith("DDPM_")})
