"""Pin the oracle (oracle/ddpm_ref.py) against outputs of the UNMODIFIED reference
(tests/golden/*.pt, produced by oracle/gen_golden.py in the build container).
CPU only.  Tolerances: tables bit-exact (fp64); network outputs 2e-5 abs / 1e-5 rel-L2
(same ATen kernels, different composition order of a few adds)."""
import math

import pytest
import torch

from oracle import ddpm_ref as R


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


TAB_KEYS = ("betas", "alphas_bar", "sqrt_alphas_bar", "sqrt_one_minus_alphas_bar",
            "sqrt_recip_alphas_bar", "sqrt_recip_m1_alphas_bar", "posterior_var",
            "posterior_logvar_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
            "fixed_model_logvar")


def test_beta_schedules_bit_exact(golden):
    tabs = golden("tables.pt")
    for s, ref in tabs["beta_schedules"].items():
        assert torch.equal(R.get_beta_schedule(s, 1e-4, 0.02, 1000), ref), s


@pytest.mark.parametrize("vt", ["fixed-large", "fixed-small"])
def test_ddpm_tables_bit_exact(golden, vt):
    tabs = golden("tables.pt")
    d = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), vt)
    for k in TAB_KEYS:
        assert torch.equal(getattr(d, k), tabs[vt][k]), k
    # known values, SURVEY.md section 8(c)
    assert abs(d.alphas_bar[0].item() - 0.9999) < 1e-12
    assert abs(d.alphas_bar[999].item() - 4.0358e-05) < 1e-8
    assert d.posterior_var[0].item() == 0.0


@pytest.mark.parametrize("name,sched,S,eta", [
    ("ddim_lin50_eta0", "linear", 50, 0.0), ("ddim_quad50_eta0", "quadratic", 50, 0.0),
    ("ddim_lin100_eta0", "linear", 100, 0.0), ("ddim_lin10_eta1", "linear", 10, 1.0),
    ("ddim_lin20_eta05", "linear", 20, 0.5)])
def test_ddim_tables_bit_exact(golden, name, sched, S, eta):
    tabs = golden("tables.pt")[name]
    sub = R.get_selection_schedule(sched, S, 1000)
    assert torch.equal(sub, tabs["subsequence"])
    d = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-small", eta=eta, subsequence=sub)
    assert d.model_var_type == tabs["model_var_type"]
    for k in TAB_KEYS:
        assert torch.equal(getattr(d, k), tabs[k]), k


def test_mirror_tables_bit_exact(golden):
    """The host mirror (ddpm_torch_b200.GaussianDiffusion / DDIM) derives the same fp64 tables as the unmodified reference."""
    import ddpm_torch_b200 as D
    tabs = golden("tables.pt")
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    for vt in ("fixed-large", "fixed-small"):
        d = D.GaussianDiffusion(betas, "eps", vt, "mse")
        for k in TAB_KEYS:
            assert torch.equal(getattr(d, k), tabs[vt][k]), (vt, k)
    for name, sched, S, eta in (("ddim_lin50_eta0", "linear", 50, 0.0), ("ddim_quad50_eta0", "quadratic", 50, 0.0),
                                ("ddim_lin100_eta0", "linear", 100, 0.0), ("ddim_lin10_eta1", "linear", 10, 1.0),
                                ("ddim_lin20_eta05", "linear", 20, 0.5)):
        sub = D.get_selection_schedule(sched, S, 1000)
        assert torch.equal(sub, tabs[name]["subsequence"])
        dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=eta, subsequence=sub)
        assert dd.model_var_type == tabs[name]["model_var_type"]
        # ddim.py:54-59: eta != 1 forces the small variance whatever the base object used
        forced = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-large", "mse"), eta=eta, subsequence=sub)
        assert forced.model_var_type == ("fixed-large" if eta == 1.0 else "fixed-small")
        for k in TAB_KEYS:
            assert torch.equal(getattr(dd, k), tabs[name][k]), (name, k)


def test_selection_schedule_known_values():
    lin = R.get_selection_schedule("linear", 50, 1000)
    assert lin[:3].tolist() == [0, 20, 40] and lin[-1].item() == 980
    quad = R.get_selection_schedule("quadratic", 50, 1000)
    assert quad[:5].tolist() == [0, 0, 1, 3, 5] and quad[-3:].tolist() == [736, 768, 800]


def test_timestep_embedding(golden):
    tabs = golden("tables.pt")
    t = torch.tensor([0, 1, 500, 999])
    assert torch.equal(R.timestep_embedding(t, 128), tabs["temb128"])
    assert torch.equal(R.timestep_embedding(t, 32), tabs["temb32"])
    e = R.timestep_embedding(torch.tensor([1]), 128)[0]
    assert torch.allclose(e[0:3], torch.tensor([0.84147, 0.76044, 0.67906]), atol=1e-5)
    assert torch.allclose(e[64:67], torch.tensor([0.54030, 0.64941, 0.73409]), atol=1e-5)


def test_param_inventory_matches_reference_counts():
    shp = R.param_shapes(R.CIFAR10_CFG)
    assert len(shp) == 304
    assert sum(math.prod(s) for s in shp.values()) == 35_746_307
    shp = R.param_shapes(R.CELEBAHQ_CFG)
    assert sum(math.prod(s) for s in shp.values()) == 113_673_219


def test_flops_match_survey():
    assert abs(R.fwd_flops_per_image(R.CIFAR10_CFG, 32, 32) / 1e9 - 12.444) < 0.01
    assert abs(R.fwd_flops_per_image(R.CELEBAHQ_CFG, 256, 256) / 1e9 - 497.03) < 0.1


@pytest.mark.parametrize("name", ["tiny", "small64", "cifar10_bs4"])
def test_unet_forward_loss_backward(golden, name):
    fx = golden(f"unet_{name}.pt")
    cfg, seed = fx["cfg"], fx["seed"]
    sd = {k: v.requires_grad_(True) for k, v in R.make_state_dict(cfg, seed).items()}
    d = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 1000), "fixed-large")
    x_t = d.q_sample(fx["x0"], fx["t"], fx["noise"])
    assert torch.equal(x_t, fx["x_t"])
    fn = lambda x, t: R.unet_forward(sd, cfg, x, t)
    eps = fn(x_t, fx["t"])
    assert rel_l2(eps, fx["eps"]) < 1e-5
    assert (eps - fx["eps"]).abs().max().item() < 5e-5
    losses = d.train_losses(fn, fx["x0"], fx["t"], fx["noise"])
    assert torch.allclose(losses, fx["losses"], rtol=1e-5, atol=1e-6)
    losses.mean().backward()
    from oracle.gen_golden import sample_idx
    for i, (k, p) in enumerate(sd.items()):
        gn = p.grad.norm().item()
        assert abs(gn - fx["grad_norm"][k]) <= 2e-4 * max(fx["grad_norm"][k], 1e-3), k
        gs = p.grad.flatten()[sample_idx(p.numel(), 64, i)]
        assert torch.allclose(gs, fx["grad_sample"][k], rtol=2e-3, atol=2e-5 * max(1.0, fx["grad_norm"][k])), k


@pytest.mark.parametrize("name", ["tiny", "cifar10_bs4"])
def test_sampler_steps(golden, name):
    fx = golden(f"unet_{name}.pt")
    cfg, seed, B = fx["cfg"], fx["seed"], fx["B"]
    sd = R.make_state_dict(cfg, seed)
    betas = R.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    with torch.no_grad():
        fn = lambda x, t: R.unet_forward(sd, cfg, x, t)
        for vt in ("fixed-large", "fixed-small"):
            d = R.RefDiffusion(betas, vt)
            for tv in (0, 1, 500, 999):
                tt = torch.full((B,), tv, dtype=torch.int64)
                xs = d.p_sample_step(fn, fx["x_t"], tt, fx[f"pstep_noise_{tv}"])
                assert (xs - fx[f"pstep_{vt}_{tv}"]).abs().max().item() < 1e-4, (vt, tv)
        # full DDIM loops; the per-step normal_ draws come from Generator(seed) in call order
        for nm, sched, S, eta in (("ddim5_lin", "linear", 5, 0.0), ("ddim5_quad", "quadratic", 5, 0.0),
                                  ("ddim4_eta1", "linear", 4, 1.0)):
            sub = R.get_selection_schedule(sched, S, 1000)
            d = R.RefDiffusion(betas, "fixed-small", eta=eta, subsequence=sub)
            g = torch.Generator().manual_seed(4321)
            zs = [torch.empty_like(fx["noise"]).normal_(generator=g) for _ in range(S)]
            xs = d.p_sample(fn, fx["noise"].clone(), zs)
            assert (xs - fx[nm]).abs().max().item() < 2e-4, nm


def test_modules(golden):
    mods = golden("modules.pt")
    for nm, fx in mods.items():
        if nm.startswith("res"):
            sd = R.fill_params(R.res_block_shapes(fx["cin"], fx["cout"], 512), fx["seed"])
            y = R.res_block(sd, "blk", fx["x"], fx["temb"])
        else:
            sd = R.fill_params(R.attn_block_shapes(fx["c"]), fx["seed"])
            y = R.attn_block(sd, "blk", fx["x"])
        assert rel_l2(y, fx["y"]) < 1e-5, nm


def test_celebahq_forward(golden):
    fx = golden("unet_celebahq_bs1.pt")
    sd = R.make_state_dict(fx["cfg"], fx["seed"])
    with torch.no_grad():
        eps = R.unet_forward(sd, fx["cfg"], fx["x"], fx["t"])
    assert rel_l2(eps, fx["eps"].float()) < 1e-3   # golden stored as fp16


@pytest.mark.parametrize("name", ["tiny"])
def test_host_mirror_generic_path_matches_reference_goldens(golden, name):
    """The host mirror handed a NON-native denoise_fn must behave exactly like the reference classes ("must fall
    back to the generic PyTorch path"): q_sample / train_losses / p_sample_step / p_sample_progressive / DDIM loops on CPU
    against the goldens written by the unmodified reference."""
    import ddpm_torch_b200 as D
    fx = golden(f"unet_{name}.pt")
    cfg, seed, B = fx["cfg"], fx["seed"], fx["B"]
    sd = R.make_state_dict(cfg, seed)
    fn = lambda x, t: R.unet_forward(sd, cfg, x, t)
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 1000)
    assert torch.equal(betas, R.get_beta_schedule("linear", 1e-4, 0.02, 1000))
    d = D.GaussianDiffusion(betas, "eps", "fixed-large", "mse")
    assert torch.equal(d.q_sample(fx["x0"], fx["t"], fx["noise"]), fx["x_t"])
    with torch.no_grad():
        assert torch.allclose(d.train_losses(fn, fx["x0"], fx["t"], fx["noise"]), fx["losses"], rtol=1e-5, atol=1e-6)
        for vt in ("fixed-large", "fixed-small"):
            dv = D.GaussianDiffusion(betas, "eps", vt, "mse")
            rv = R.RefDiffusion(betas, vt)
            for tv in (0, 1, 500, 999):
                tt = torch.full((B,), tv, dtype=torch.int64)
                xs = dv.p_sample_step(fn, fx["x_t"], tt, generator=torch.Generator().manual_seed(3))
                z = torch.empty_like(fx["x_t"]).normal_(generator=torch.Generator().manual_seed(3))
                assert (xs - rv.p_sample_step(fn, fx["x_t"], tt, z)).abs().max().item() < 1e-5, (vt, tv)
                assert (dv.p_sample_step(fn, fx["x_t"], tt, generator=torch.Generator().manual_seed(3)) - xs).abs().max().item() == 0
        for nm, sched, S, eta in (("ddim5_lin", "linear", 5, 0.0), ("ddim5_quad", "quadratic", 5, 0.0), ("ddim4_eta1", "linear", 4, 1.0)):
            sub = D.get_selection_schedule(sched, S, 1000)
            dd = D.DDIM.from_ddpm(D.GaussianDiffusion(betas, "eps", "fixed-small", "mse"), eta=eta, subsequence=sub)
            out = dd.p_sample(fn, shape=tuple(fx["noise"].shape), device=torch.device("cpu"), noise=fx["noise"], seed=4321)
            assert (out - fx[nm]).abs().max().item() < 2e-4, nm
