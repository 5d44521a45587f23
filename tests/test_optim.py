"""Parameter update (clip_grad_norm_ + Adam + LambdaLR warm-up + EMA, ddpm_torch/utils/train.py:159-165,280-316).

CPU: the oracle restatement (oracle/optim_ref.py) against the golden produced by the real torch.optim.Adam /
clip_grad_norm_ / LambdaLR and the UNMODIFIED reference EMA class (oracle/gen_golden.py::gen_optim).
GPU: the fused kernels through the C ABI against the same golden, and the FusedAdam/EMA host mirror against the oracle
on a real UNet backward."""
import ctypes as C
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import optim_ref as O  # noqa: E402
from oracle import ddpm_ref as R  # noqa: E402

RTOL, ATOL = 2e-6, 1e-9        # fp32 elementwise arithmetic; differences are fma-contraction / division-rounding ulps


def test_oracle_update_matches_reference_golden(golden):
    fx = golden("optim.pt")
    h = fx["hyper"]
    p = [t.clone() for t in fx["p0"]]
    m = [torch.zeros_like(t) for t in p]; v = [torch.zeros_like(t) for t in p]
    sh = [t.clone() for t in p]                                  # EMA.__init__ clones the parameters (utils/train.py:292)
    for k in range(len(fx["grads"])):
        lr = O.lr_at(h["lr"], h["warmup"], k)
        assert lr == pytest.approx(fx["lrs"][k], rel=1e-12)
        tn, _ = O.update(p, fx["grads"][k], m, v, sh, step=k + 1, lr=lr, beta1=h["beta1"], beta2=h["beta2"],
                         max_norm=h["grad_norm"], ema_decay=h["ema_decay"], ema_num_updates=k)
        assert tn == pytest.approx(fx["norms"][k], rel=1e-6)
        for a, b in zip(p, fx["params"][k]):
            torch.testing.assert_close(a, b, rtol=RTOL, atol=ATOL)
        for a, b in zip(sh, fx["shadow"][k]):
            torch.testing.assert_close(a, b, rtol=RTOL, atol=ATOL)
    assert fx["ema_num_updates"] == len(fx["grads"]) - 1


def test_schedules():
    assert [O.lr_at(2e-4, 4, k) for k in range(5)] == [2e-4 * 0.25, 2e-4 * 0.5, 2e-4 * 0.75, 2e-4, 2e-4]
    assert O.ema_decay_at(0.9999, 0) == 0.1 and O.ema_decay_at(0.9999, 90) == pytest.approx(0.91)
    assert O.ema_decay_at(0.9999, 10 ** 7) == 0.9999


def _flat(ts, dev):
    return torch.cat([t.reshape(-1) for t in ts]).to(dev).contiguous()


@pytest.mark.gpu
def test_fused_kernels_match_reference_golden(golden):
    from ddpm_torch_b200 import _lib
    L = _lib.lib(); _lib.runtime_check()
    fx = golden("optim.pt"); h = fx["hyper"]; dev = "cuda"
    p = _flat(fx["p0"], dev); n = p.numel()
    assert n % 4 == 0
    m = torch.zeros_like(p); v = torch.zeros_like(p); sh = p.clone()
    state = torch.zeros(16, device=dev); norm = torch.zeros(2, device=dev)
    for k in range(len(fx["grads"])):
        g = _flat(fx["grads"][k], dev)
        cfg = _lib.OptCfg()
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = O.lr_at(h["lr"], h["warmup"], k), h["beta1"], h["beta2"], 1e-8
        cfg.max_grad_norm, cfg.ema_decay, cfg.step, cfg.ema_num_updates = h["grad_norm"], h["ema_decay"], k + 1, k
        _lib.check(L.ddpm_opt_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), sh.data_ptr(), n, C.byref(cfg),
                                   state.data_ptr(), norm.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert norm[0].item() == pytest.approx(fx["norms"][k], rel=1e-6)
        assert norm[1].item() == pytest.approx(min(1.0, h["grad_norm"] / (fx["norms"][k] + 1e-6)), rel=1e-6)
        torch.testing.assert_close(p.cpu(), _flat(fx["params"][k], "cpu"), rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(sh.cpu(), _flat(fx["shadow"][k], "cpu"), rtol=RTOL, atol=ATOL)
    assert L.ddpm_device_error_flag() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("use_ema,clip", [(True, 1.0), (False, 0.0)])
def test_fused_adam_on_unet_backward_matches_oracle(use_ema, clip):
    """FusedAdam(+EMA) after real engine backward passes vs the oracle update applied to copies of the same flat buffers."""
    import ddpm_torch_b200 as D
    from ddpm_torch_b200.optim import EMA, FusedAdam
    dev = torch.device("cuda")
    torch.manual_seed(3)
    cfg = dict(R.TINY_CFG)
    model = D.UNet(**{k: cfg[k] for k in ("in_channels", "hid_channels", "out_channels", "ch_multipliers", "num_res_blocks", "apply_attn")},
                   drop_rate=0.0).to(dev).train()
    with torch.no_grad():
        for p in model.parameters():
            if p.ndim >= 2:
                p.copy_(torch.randn_like(p) * 0.05)
    diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-large", "mse")
    ema = EMA(model, 0.9999) if use_ema else None
    opt = FusedAdam(model, lr=1e-3, betas=(0.9, 0.999), warmup=3, grad_norm=clip, ema=ema)
    p_ref = model.flat_params.detach().clone()
    m_ref = torch.zeros_like(p_ref); v_ref = torch.zeros_like(p_ref); sh_ref = p_ref.clone() if use_ema else None
    g = torch.Generator(device=dev).manual_seed(11)
    B = 4
    for k in range(4):
        x = torch.rand(B, 3, 32, 32, device=dev, generator=g) * 2 - 1
        t = torch.randint(1000, (B,), device=dev, generator=g); nz = torch.randn(B, 3, 32, 32, device=dev, generator=g)
        assert torch.equal(model.flat_params, p_ref)                     # both sides start the step from identical weights
        opt.zero_grad()
        diff.train_losses(model, x, t, nz).mean().backward()
        grads = model.flat_grads.detach().clone()
        lr = opt.lr
        assert lr == pytest.approx(O.lr_at(1e-3, 3, k))
        out = opt.step()
        tn, coef = O.update([p_ref], [grads], [m_ref], [v_ref], [sh_ref] if use_ema else None, step=k + 1, lr=lr,
                            max_norm=clip, ema_decay=0.9999, ema_num_updates=k)
        assert out[0].item() == pytest.approx(tn, rel=1e-5)
        assert out[1].item() == pytest.approx(coef, rel=1e-5)
        torch.testing.assert_close(model.flat_params, p_ref, rtol=1e-5, atol=1e-8)
        # the clip coefficient differs in its last ulp between the two norm reductions, and m' = m + 0.1 (c g - m) cancels:
        # the error is ~1e-7 of the OPERANDS, not of the (possibly tiny) result -> absolute tolerance scaled by max|g|
        gmax = grads.abs().max().item()
        torch.testing.assert_close(opt.exp_avg, m_ref, rtol=1e-5, atol=1e-6 * gmax)
        torch.testing.assert_close(opt.exp_avg_sq, v_ref, rtol=1e-5, atol=1e-6 * gmax * gmax)
        if use_ema:
            torch.testing.assert_close(ema.flat, sh_ref, rtol=1e-5, atol=1e-8)
        with torch.no_grad():
            model.flat_params.copy_(p_ref)                               # remove ulp drift so the next step is comparable
    if use_ema:
        assert ema.num_updates == 3
        before = model.flat_params.clone()
        with ema:                                                        # utils/train.py:179: sampling under the EMA weights
            assert torch.equal(model.flat_params, ema.flat)
            name, p0 = next(iter(model.named_parameters()))
            assert torch.equal(p0, ema.shadow[name])
        assert torch.equal(model.flat_params, before)
        sd = ema.state_dict()
        assert set(sd) == {"decay", "shadow", "num_updates"} and set(sd["shadow"]) == {k for k, _ in model.named_parameters()}
    # the optimizer state has torch.optim.Adam's layout: it loads into a stock Adam over the same parameters and back
    sd = opt.state_dict()
    stock = torch.optim.Adam(model.parameters(), lr=1e-3)
    stock.load_state_dict({"state": sd["state"], "param_groups": [{**stock.state_dict()["param_groups"][0], "lr": sd["param_groups"][0]["lr"]}]})
    assert float(stock.state[next(iter(model.parameters()))]["step"]) == 4.0
    opt2 = FusedAdam(model, lr=1e-3, warmup=3)
    opt2.load_state_dict(stock.state_dict())
    assert opt2.steps == 4 and torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)


def test_optim_refuses_cpu_and_foreign_models():
    from ddpm_torch_b200.optim import FusedAdam
    with pytest.raises(RuntimeError, match="native"):
        FusedAdam(torch.nn.Linear(4, 4))
    if not torch.cuda.is_available():
        import ddpm_torch_b200 as D
        cfg = R.TINY_CFG
        m = D.UNet(**{k: cfg[k] for k in ("in_channels", "hid_channels", "out_channels", "ch_multipliers", "num_res_blocks", "apply_attn")})
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            FusedAdam(m)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 3, 32, 32), (2, 3, 256, 256), (3, 1, 7, 5), (2, 4, 16, 16)])
def test_to_uint8_nhwc_bit_exact(shape):
    """generate.py:129 on the device: bit-exact against the literal reference expression evaluated on the CPU, including
    rounding ties (k + 0.5 -> even), out-of-range values and the exact end points."""
    from ddpm_torch_b200.postprocess import to_uint8_nhwc, to_uint8_host_async
    g = torch.Generator().manual_seed(9)
    x = torch.randn(shape, generator=g) * 0.8
    flat = x.view(-1)
    ties = (torch.arange(0, 256, dtype=torch.float32) + 0.5 - 127.5) / 127.5           # pre-images of k + 0.5
    n = min(flat.numel() // 2, ties.numel())
    flat[:n] = ties[:n]
    flat[n:n + 6] = torch.tensor([-1.0, 1.0, -3.0, 3.0, 0.0, 1.0 - 2 ** -24])
    ref = R.to_uint8_nhwc(x)                                                            # CPU, as the reference runs it
    out = to_uint8_nhwc(x.cuda())
    assert out.dtype == torch.uint8 and out.shape == ref.shape and out.is_contiguous()
    assert torch.equal(out.cpu(), ref.contiguous())
    pinned, ev = to_uint8_host_async(x.cuda())
    ev.synchronize()
    assert pinned.is_pinned() and torch.equal(pinned, ref.contiguous())


def test_to_uint8_refuses_cpu():
    from ddpm_torch_b200.postprocess import to_uint8_nhwc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        to_uint8_nhwc(torch.zeros(1, 3, 4, 4))
