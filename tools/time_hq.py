"""Informational timing of the CelebA-HQ config (configs/celebahq.json): training step at 4 images/GPU, forward at 8."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_torch_b200 as D
from ddpm_torch_b200 import _lib
HQ = dict(in_channels=3, hid_channels=128, out_channels=3, ch_multipliers=(1, 1, 2, 2, 4, 4), num_res_blocks=2,
          apply_attn=(False, False, False, False, True, False), drop_rate=0.0)
dev = torch.device("cuda")
model = D.UNet(**HQ).to(dev).train()
with torch.no_grad():
    gi = torch.Generator(device=dev).manual_seed(7)
    for n_, p in model.named_parameters():
        if p.ndim >= 2:
            p.copy_((torch.rand(p.shape, device=dev, generator=gi) * 2 - 1) * (3.0 / p[0].numel()) ** 0.5)
diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-small", "mse")
L = _lib.lib()
res = {}
for mode, B in (("train", 4), ("fwd", 8)):
    g = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.rand(B, 3, 256, 256, device=dev, generator=g) * 2 - 1
    t = torch.randint(1000, (B,), device=dev, generator=g); nz = torch.randn(B, 3, 256, 256, device=dev, generator=g)
    if mode == "train":
        h = model.prepare(B, 256, 256, training=True)
        ta, tsb = diff._dev_tables(dev); losses = torch.empty(B, device=dev); gs = torch.full((B,), 1.0 / B, device=dev)
        def step():
            sp = _lib.stream_ptr()
            _lib.check(L.ddpm_unet_repack(h, sp))
            _lib.check(L.ddpm_train_forward(h, x0.data_ptr(), t.data_ptr(), nz.data_ptr(), ta.data_ptr(), tsb.data_ptr(), losses.data_ptr(), 0, sp))
            _lib.check(L.ddpm_train_backward(h, gs.data_ptr(), sp))
    else:
        model.eval()
        h = model.prepare(B, 256, 256, training=False)
        out = torch.empty_like(x0)
        def step():
            _lib.check(L.ddpm_unet_forward(h, x0.data_ptr(), t.data_ptr(), out.data_ptr(), 0, _lib.stream_ptr()))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gf = 497.03 * B * (3 if mode == "train" else 1)
    res[mode] = {"B": B, "ms": ms, "img_s": B / ms * 1e3, "tflops": gf / ms}
    print(mode, res[mode], flush=True)
assert L.ddpm_device_error_flag() == 0
json.dump(res, open("gpurun_out/hq_timing.json", "w"))
