"""Parity of the engine-backed UNet / GaussianDiffusion / DDIM (through the C ABI) against the oracle
(oracle/ddpm_ref.py, fp32, TF32 off) and against the committed golden outputs of the unmodified reference.

Arithmetic: bf16 operands / bf16 inter-layer activations, fp32 accumulation and statistics.  Stated tolerances
(relative L2 unless noted): eps-prediction <= 2e-2 (full UNet), per-sample MSE loss <= 2e-2 relative,
parameter gradients <= 5e-2 over the flat gradient vector and cosine >= 0.995 per tensor with non-trivial norm,
sampler step pixel Linf <= 5e-2 * step scale.  Measured values are printed (run with -s) and recorded in DESIGN.md."""
import math
import os

import pytest
import torch

from oracle import ddpm_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def build(cfg, seed, train=False):
    import ddpm_torch_b200 as D
    c = R.normalize_cfg(cfg)
    m = D.UNet(in_channels=c["in_channels"], hid_channels=c["hid_channels"], out_channels=c["out_channels"],
               ch_multipliers=c["ch_multipliers"], num_res_blocks=c["num_res_blocks"], apply_attn=c["apply_attn"],
               drop_rate=0.0)
    sd = R.make_state_dict(cfg, seed)
    m.load_state_dict(sd)
    m = m.to(DEV)
    m.train(train)
    return m, {k: v.to(DEV) for k, v in sd.items()}


def check_device_flag():
    from ddpm_torch_b200 import _lib
    assert _lib.lib().ddpm_device_error_flag() == 0


CASES = [("tiny", 2e-2), ("small64", 2e-2), ("cifar10_bs4", 2e-2)]


@pytest.mark.parametrize("name,tol", CASES)
def test_forward_vs_oracle_and_golden(golden, name, tol):
    fx = golden(f"unet_{name}.pt")
    m, sd = build(fx["cfg"], fx["seed"])
    x, t = fx["x_t"].to(DEV), fx["t"].to(DEV)
    with torch.no_grad():
        eps = m(x, t)
        ref = R.unet_forward(sd, fx["cfg"], x, t)
    check_device_flag()
    r1, r2 = rel(eps, ref), rel(eps.cpu(), fx["eps"])
    print(f"\n[{name}] eps rel-L2 vs oracle {r1:.3e}, vs reference golden {r2:.3e}, max-abs {(eps - ref).abs().max().item():.3e}")
    assert r1 < tol and r2 < tol


@pytest.mark.parametrize("name", ["tiny", "small64", "cifar10_bs4"])
def test_train_losses_and_grads(golden, name):
    import ddpm_torch_b200 as D
    fx = golden(f"unet_{name}.pt")
    cfg = fx["cfg"]
    m, sd = build(cfg, fx["seed"], train=True)
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    diff = D.GaussianDiffusion(betas, "eps", "fixed-large", "mse")
    x0, t, noise = fx["x0"].to(DEV), fx["t"].to(DEV), fx["noise"].to(DEV)
    losses = diff.train_losses(m, x0, t, noise)
    assert losses.shape == (x0.shape[0],)
    losses.mean().backward()
    check_device_flag()
    # oracle
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rd = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-large")
    lref = rd.train_losses(lambda x, tt: R.unet_forward(sdg, cfg, x, tt), x0, t, noise)
    lref.mean().backward()
    lr = (losses - lref).abs().max().item() / lref.abs().max().item()
    print(f"\n[{name}] loss rel err {lr:.3e}; golden {(losses.cpu() - fx['losses']).abs().max().item():.3e}")
    assert lr < 2e-2
    gm = torch.cat([p.grad.flatten() for p in m.parameters()])
    gr = torch.cat([sdg[k].grad.flatten() for k, _ in m.named_parameters()])
    worst = (1.0, "")
    bad = []
    for k, p in m.named_parameters():
        a, b = p.grad.flatten().double(), sdg[k].grad.flatten().double()
        if b.norm() < 1e-6 * gr.norm():
            continue
        cos = (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()
        if cos < worst[0]:
            worst = (cos, k)
        if cos < 0.995:
            bad.append((k, cos, rel(a, b)))
    print(f"[{name}] flat grad rel-L2 {rel(gm, gr):.3e}; worst per-tensor cosine {worst[0]:.5f} ({worst[1]})")
    for k, cos, r in bad[:20]:
        print("   BAD", k, f"cos {cos:.4f} rel {r:.3e}")
    assert rel(gm, gr) < 5e-2 and not bad
    # golden grad norms of the unmodified reference
    for k, p in m.named_parameters():
        gn = fx["grad_norm"][k]
        if gn > 1e-3:
            assert abs(p.grad.norm().item() - gn) / gn < 0.1, k


def test_plain_forward_backward_autograd(golden):
    """denoise_fn seam: model(x, t) under autograd with an arbitrary upstream gradient."""
    fx = golden("unet_small64.pt")
    cfg = fx["cfg"]
    m, sd = build(cfg, fx["seed"], train=True)
    x, t = fx["x_t"].to(DEV), fx["t"].to(DEV)
    g = torch.randn(x.shape[0], 3, x.shape[2], x.shape[3], device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    out = m(x, t)
    (out * g).sum().backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (R.unet_forward(sdg, cfg, x, t) * g).sum().backward()
    gm = torch.cat([p.grad.flatten() for p in m.parameters()])
    gr = torch.cat([sdg[k].grad.flatten() for k, _ in m.named_parameters()])
    print(f"\n[autograd] flat grad rel-L2 {rel(gm, gr):.3e}")
    assert rel(gm, gr) < 5e-2


@pytest.mark.parametrize("name", ["tiny", "cifar10_bs4"])
def test_sampler_steps_vs_golden(golden, name):
    """Single ancestral steps with injected noise (diffusion.py:152-158) and short DDIM loops with the torch RNG."""
    import ddpm_torch_b200 as D
    fx = golden(f"unet_{name}.pt")
    m, sd = build(fx["cfg"], fx["seed"])
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    B = fx["B"]
    for vt in ("fixed-large", "fixed-small"):
        d = D.GaussianDiffusion(betas, "eps", vt, "mse")
        for tv in (0, 1, 500, 999):
            tt = torch.full((B,), tv, dtype=torch.int64, device=DEV)
            g = torch.Generator(DEV).manual_seed(1)
            with torch.no_grad():
                xs = d.p_sample_step(m, fx["x_t"].to(DEV), tt, generator=g)      # generic tail over the native model
            z = torch.empty_like(xs).normal_(generator=torch.Generator(DEV).manual_seed(1))
            rd = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), vt)
            with torch.no_grad():
                ref = rd.p_sample_step(lambda x, q: R.unet_forward(sd, fx["cfg"], x, q), fx["x_t"].to(DEV), tt, z)
            scale = ref.abs().max().item()
            err = (xs - ref).abs().max().item()
            assert err < 5e-2 * max(scale, 1.0), (vt, tv, err, scale)
    # engine sampler loop (DDIM, 5 steps, eta=0 and 4 steps eta=1) vs golden of the unmodified reference.
    # the reference drew its noise from a CPU generator; reproduce that stream and feed it through rng="torch" on device is
    # not possible bit-for-bit (CUDA Philox != CPU MT), so eta=0 (noise-free up to 1e-10) is compared to the golden and
    # eta=1 is compared to the oracle driven with the same CUDA generator.
    for nm, sched, S, eta in (("ddim5_lin", "linear", 5, 0.0), ("ddim5_quad", "quadratic", 5, 0.0)):
        sub = D.get_selection_schedule(sched, S, 1000)
        dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=eta, subsequence=sub)
        for use_graph in (False, True):
            xs = dd.p_sample(m, shape=tuple(fx["noise"].shape), device=torch.device(DEV), noise=fx["noise"].to(DEV), seed=4321, use_graph=use_graph)
            dlt = (xs.cpu() - fx[nm]).abs()
            err, l1, frac = dlt.max().item(), dlt.mean().item(), (dlt > 0.1).float().mean().item()
            print(f"\n[{name}] {nm} graph={use_graph} vs reference golden: pixel L1 {l1:.3e}, Linf {err:.3e}, frac(|d|>0.1) {frac:.4f}")
            # random-weight models saturate x0 at +-1, so a borderline pixel can flip by 2: bound the mean and the flip rate
            assert l1 < 2e-2 and frac < 0.03, (nm, l1, frac)
    check_device_flag()
    sub = D.get_selection_schedule("linear", 4, 1000)
    dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=1.0, subsequence=sub)
    xs = dd.p_sample(m, shape=tuple(fx["noise"].shape), device=torch.device(DEV), noise=fx["noise"].to(DEV), seed=99)
    g = torch.Generator(DEV).manual_seed(99)
    zs = [torch.empty(fx["noise"].shape, device=DEV).normal_(generator=g) for _ in range(4)]
    rd = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-small", eta=1.0, subsequence=sub)
    with torch.no_grad():
        ref = rd.p_sample(lambda x, q: R.unet_forward(sd, fx["cfg"], x, q), fx["noise"].to(DEV), zs)
    dlt = (xs - ref).abs()
    print(f"[{name}] ddim4 eta=1 torch-rng vs oracle: pixel L1 {dlt.mean().item():.3e}, Linf {dlt.max().item():.3e}")
    assert dlt.mean().item() < 2e-2 and (dlt > 0.1).float().mean().item() < 0.03


def test_philox_sampler_runs():
    import ddpm_torch_b200 as D
    m, _ = build(R.SMALL64_CFG, 5)
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    sub = D.get_selection_schedule("linear", 8, 1000)
    dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=1.0, subsequence=sub)
    xs = dd.p_sample(m, shape=(4, 3, 32, 32), device=torch.device(DEV), seed=7, rng="philox")
    assert torch.isfinite(xs).all()
    xs2 = dd.p_sample(m, shape=(4, 3, 32, 32), device=torch.device(DEV), seed=7, rng="philox")
    # same seed -> same noise stream; GroupNorm statistics use (order-dependent) atomics, so allow last-bit drift
    assert (xs - xs2).abs().mean().item() < 1e-2
    check_device_flag()


def test_dropout_train_mode_statistics():
    """drop_rate>0 in train mode: not bit-comparable to torch's mask; check determinism per seed and that backward runs."""
    import ddpm_torch_b200 as D
    cfg = dict(R.SMALL64_CFG); cfg["drop_rate"] = 0.1
    c = R.normalize_cfg(cfg)
    m = D.UNet(3, 64, 3, c["ch_multipliers"], c["num_res_blocks"], c["apply_attn"], drop_rate=0.1)
    m.load_state_dict(R.make_state_dict(cfg, 5))
    m = m.to(DEV).train()
    x = torch.randn(4, 3, 32, 32, device=DEV); t = torch.randint(1000, (4,), device=DEV)
    with torch.no_grad():
        a = m(x, t); b = m(x, t)
    assert not torch.equal(a, b)                      # different masks on successive calls
    m.eval()
    with torch.no_grad():
        c1 = m(x, t); c2 = m(x, t)
    # eval mode: no dropout.  Reductions use atomics, so two runs differ in summation order; in a bf16 network any 1e-7
    # perturbation re-randomises downstream roundings, i.e. run-to-run drift sits at the bf16 noise floor (measured 3.6e-3),
    # far below the effect of a different dropout mask
    assert rel(c1, c2) < 1e-2 and rel(a, b) > 5 * rel(c1, c2)
    m.train()
    out = m(x, t); out.square().mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())
    check_device_flag()


def test_celebahq_forward_golden(golden):
    fx = golden("unet_celebahq_bs1.pt")
    m, _ = build(fx["cfg"], fx["seed"])
    with torch.no_grad():
        eps = m(fx["x"].to(DEV), fx["t"].to(DEV))
    r = rel(eps.cpu(), fx["eps"].float())
    print(f"\n[celebahq bs1 256x256] eps rel-L2 vs reference golden {r:.3e}")
    assert r < 3e-2
    check_device_flag()


def test_dropout_gradient_directional_derivative():
    """Train mode with dropout (the engine's own Philox masks, saved by the forward and re-used by the backward): the
    masks cannot be compared with torch's, so check the gradient by a directional derivative under the SAME seed:
    L(w - eps*g) ~= L(w) - eps*|g|^2."""
    import ddpm_torch_b200 as D
    cfg = dict(R.SMALL64_CFG); cfg["drop_rate"] = 0.1
    c = R.normalize_cfg(cfg)
    m = D.UNet(3, 64, 3, c["ch_multipliers"], c["num_res_blocks"], c["apply_attn"], drop_rate=0.1)
    m.load_state_dict(R.make_state_dict(cfg, 7))
    m = m.to(DEV).train()
    diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-large", "mse")
    g = torch.Generator(DEV).manual_seed(3)
    x0 = torch.randn(8, 3, 32, 32, device=DEV, generator=g); t = torch.randint(1000, (8,), device=DEV, generator=g)
    nz = torch.randn(8, 3, 32, 32, device=DEV, generator=g)

    def loss_with_seed():
        m._drop_calls = 41                        # same dropout seed on every evaluation
        return diff.train_losses(m, x0, t, nz).mean()

    l0 = loss_with_seed()
    l0.backward()
    grads = [p.grad.clone() for p in m.parameters()]
    g2 = sum((gg.double() ** 2).sum() for gg in grads).item()
    with torch.no_grad():
        l0b = loss_with_seed().item()
    assert abs(l0b - l0.item()) < 2e-3 * abs(l0.item())          # same seed -> same masks (up to reduction-order noise)
    eps = 0.05 * l0.item() / g2                                  # aim at a 5 % first-order decrease
    with torch.no_grad():
        for p, gg in zip(m.parameters(), grads):
            p.sub_(eps * gg)
        l1 = loss_with_seed().item()
    pred, got = eps * g2, l0.item() - l1
    print(f"\n[dropout grad] L0 {l0.item():.5f} predicted decrease {pred:.5f} measured {got:.5f}")
    assert 0.7 * pred < got < 1.3 * pred
    check_device_flag()


@pytest.mark.gpu
def test_sampler_split_batch_two_streams_matches_unsplit(golden):
    """The split sampler (two half batches on two forked streams inside one captured graph) draws the same torch noise
    per image and must reproduce the unsplit loop (up to the engine's run-to-run atomics noise)."""
    import ddpm_torch_b200 as D
    fx = golden("unet_tiny.pt")
    m, sd = build(fx["cfg"], fx["seed"])
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    sub = D.get_selection_schedule("linear", 6, 1000)
    dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=1.0, subsequence=sub)
    H = fx["x_t"].shape[-1]
    noise = torch.randn(64, 3, H, H, generator=torch.Generator().manual_seed(5)).to(DEV)
    outs = {}
    for split in (1, 2):
        for use_graph in (True, False):
            outs[(split, use_graph)] = dd.p_sample(m, shape=tuple(noise.shape), device=torch.device(DEV), noise=noise, seed=77,
                                                   use_graph=use_graph, split=split)
    ref = outs[(1, True)]
    for k, v in outs.items():
        d = (v - ref).abs()
        print(f"\n[split sampler] split={k[0]} graph={k[1]}: mean |d| {d.mean().item():.3e}, frac(|d|>0.1) {(d > 0.1).float().mean().item():.4f}")
        assert d.mean().item() < 1e-2 and (d > 0.1).float().mean().item() < 0.03
    # halves must not be copies of each other (distinct images, distinct noise)
    assert (outs[(2, True)][:32] - outs[(2, True)][32:]).abs().mean().item() > 1e-2
    check_device_flag()


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", [("tiny", 64), ("cifar10_bs4", 64)])
def test_train_batch64_tensor_core_paths(golden, name, B):
    """Batch-size-dependent plan choices that the bs=4 fixtures never reach: the tensor-core timestep projections (one GEMM for
    the 22 fc layers + their MN-major weight gradient, power-of-two batch >= 64), one-wave split-K choices, GroupNorm grids.
    Loss and every gradient tensor against the oracle on the same seeded inputs."""
    import ddpm_torch_b200 as D
    fx = golden(f"unet_{name}.pt")
    cfg = fx["cfg"]
    m, sd = build(cfg, fx["seed"], train=True)
    diff = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 1000), "eps", "fixed-large", "mse")
    g = torch.Generator(DEV).manual_seed(21)
    H = fx["x0"].shape[-1]
    x0 = torch.randn(B, 3, H, H, device=DEV, generator=g); t = torch.randint(1000, (B,), device=DEV, generator=g)
    noise = torch.randn(B, 3, H, H, device=DEV, generator=g)
    losses = diff.train_losses(m, x0, t, noise)
    losses.mean().backward()
    check_device_flag()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rd = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-large")
    lref = rd.train_losses(lambda x, tt: R.unet_forward(sdg, cfg, x, tt), x0, t, noise)
    lref.mean().backward()
    lr = (losses - lref).abs().max().item() / lref.abs().max().item()
    gm = torch.cat([p.grad.flatten() for p in m.parameters()])
    gr = torch.cat([sdg[k].grad.flatten() for k, _ in m.named_parameters()])
    worst = (1.0, "")
    for k, p in m.named_parameters():
        a, b = p.grad.flatten().double(), sdg[k].grad.flatten().double()
        if b.norm() < 1e-6 * gr.norm():
            continue
        cos = (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()
        if cos < worst[0]:
            worst = (cos, k)
    print(f"\n[{name} B={B}] loss rel err {lr:.3e}; flat grad rel-L2 {rel(gm, gr):.3e}; worst per-tensor cosine {worst[0]:.5f} ({worst[1]})")
    assert lr < 2e-2 and rel(gm, gr) < 5e-2 and worst[0] > 0.995, worst
