"""One training step (CIFAR bs=128) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_torch_b200 as D
from ddpm_torch_b200 import _lib
import bench
mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda")
model = D.UNet(**bench.CIFAR).to(dev).train()
with torch.no_grad():
    gi = torch.Generator(device=dev).manual_seed(7)
    for n_, p in model.named_parameters():
        if p.ndim >= 2:
            p.copy_((torch.rand(p.shape, device=dev, generator=gi) * 2 - 1) * (3.0 / p[0].numel()) ** 0.5)
diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-large", "mse")
g = torch.Generator(device=dev).manual_seed(1)
x0 = torch.rand(B, 3, 32, 32, device=dev, generator=g) * 2 - 1
t = torch.randint(1000, (B,), device=dev, generator=g); nz = torch.randn(B, 3, 32, 32, device=dev, generator=g)
L = _lib.lib()
if mode == "train":
    h = model.prepare(B, 32, 32, training=True)
    ta, tsb = diff._dev_tables(dev); losses = torch.empty(B, device=dev); gs = torch.full((B,), 1.0 / B, device=dev)
    def step(i):
        sp = _lib.stream_ptr()
        _lib.check(L.ddpm_unet_repack(h, sp))
        _lib.check(L.ddpm_train_forward(h, x0.data_ptr(), t.data_ptr(), nz.data_ptr(), ta.data_ptr(), tsb.data_ptr(), losses.data_ptr(), 5 + i, sp))
        _lib.check(L.ddpm_train_backward(h, gs.data_ptr(), sp))
else:
    model.eval()
    h = model.prepare(B, 32, 32, training=False)
    out = torch.empty_like(x0)
    def step(i):
        _lib.check(L.ddpm_unet_forward(h, x0.data_ptr(), t.data_ptr(), out.data_ptr(), 0, _lib.stream_ptr()))
for i in range(2):
    step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
