"""Bits-per-dim / KL path of the host mirror (generic PyTorch formulas) against goldens written by the
UNMODIFIED reference (tests/golden/bpd_toy.pt, oracle/gen_golden.py::gen_bpd) on a toy non-native denoiser; plus the parts
that raise upstream (``_prior_bpd``: TorchScript type error; ``calc_all_bpd``: shape-tuple unpack; "learned" variance:
KeyError in the constructor) checked against closed forms / compositions of the pinned terms."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ddpm_ref as R  # noqa: E402


def test_loss_terms_and_kl_training_loss_match_reference(golden):
    import ddpm_torch_b200 as D
    fx = golden("bpd_toy.pt")
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, fx["T"])
    fn = R.toy_denoiser(3, 1, seed=5)
    assert len(fx["cases"]) == 6
    for (mt, vt), c in fx["cases"].items():
        d = D.GaussianDiffusion(betas, mt, vt, "kl")
        x_t = d.q_sample(fx["x0"], fx["t"], noise=fx["noise"])
        assert torch.equal(x_t, c["x_t"]), (mt, vt)
        term, pred = d._loss_term_bpd(fn, x_0=fx["x0"], x_t=x_t, t=fx["t"], clip_denoised=True, return_pred=True)
        torch.testing.assert_close(term, c["term"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pred, c["pred_x_0"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(d.train_losses(fn, fx["x0"], fx["t"], noise=fx["noise"]), c["kl_loss"], rtol=1e-5, atol=1e-6)
        mean, var, logvar = d.p_mean_var(fn, x_t, fx["t"], clip_denoised=False, return_pred=False)
        torch.testing.assert_close(mean, c["mean"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(var.expand_as(mean), c["var"], rtol=1e-6, atol=0)
        torch.testing.assert_close(logvar.expand_as(mean), c["logvar"], rtol=1e-6, atol=0)
    # t = 0 is the discretised decoder NLL, independent of the variance parameterisation's KL branch; t > 0 differ
    a, b = fx["cases"][("eps", "fixed-small")]["term"], fx["cases"][("eps", "fixed-large")]["term"]
    assert a[0] == b[0] and a[1] != b[1]


def test_functions_against_closed_forms():
    from ddpm_torch_b200.functions import approx_std_normal_cdf, discretized_gaussian_loglik, flat_mean, normal_kl
    m1, m2 = torch.tensor([0.3, -1.0]), torch.tensor([0.1, 0.5])
    lv1, lv2 = torch.tensor([0.2, -0.7]), torch.tensor([-0.4, 0.3])
    v1, v2 = lv1.exp(), lv2.exp()
    ref = 0.5 * (lv2 - lv1 + (v1 + (m1 - m2) ** 2) / v2 - 1)
    torch.testing.assert_close(normal_kl(m1, lv1, m2, lv2), ref, rtol=1e-6, atol=1e-7)
    assert torch.all(normal_kl(m1, lv1, m1, lv1).abs() < 1e-7)
    x = torch.linspace(-3, 3, 13)
    exact = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    assert (approx_std_normal_cdf(x) - exact).abs().max() < 3e-4            # the tanh approximation's accuracy
    # probabilities of all 256 bins sum to one (edge bins absorb the tails)
    bins = (torch.arange(256, dtype=torch.float32) / 127.5 - 1.0)
    lp = discretized_gaussian_loglik(bins, torch.full((256,), 0.13), log_scale=torch.full((256,), -1.2))
    assert abs(lp.exp().sum().item() - 1.0) < 2e-3
    assert flat_mean(torch.ones(2, 3, 4, 5)).shape == (2,)


def test_prior_and_total_bpd_compose_the_pinned_terms(golden):
    """``_prior_bpd`` and ``calc_all_bpd`` raise upstream; here: prior against the closed-form KL, and calc_all_bpd against a
    manual loop over the reference-pinned ``_loss_term_bpd`` with the same RNG stream."""
    import ddpm_torch_b200 as D
    fx = golden("bpd_toy.pt")
    T = fx["T"]
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, T)
    fn = R.toy_denoiser(3, 1, seed=5)
    d = D.GaussianDiffusion(betas, "eps", "fixed-large", "kl")
    x0 = fx["x0"]
    ab = d.alphas_bar[-1].item()
    mean, var = math.sqrt(ab) * x0, 1 - ab
    closed = (0.5 * (-1 - math.log(var) + mean ** 2 + var)).mean(dim=(1, 2, 3)) / math.log(2.)
    torch.testing.assert_close(d._prior_bpd(x0), closed.to(torch.float32), rtol=1e-5, atol=1e-6)
    torch.manual_seed(123)
    total, losses, prior, mses = d.calc_all_bpd(fn, x0, clip_denoised=True)
    assert total.shape == (4,) and losses.shape == (4, T) and mses.shape == (4, T)
    torch.manual_seed(123)
    t = torch.empty(4, dtype=torch.int64)
    for ti in range(T - 1, -1, -1):
        t.fill_(ti)
        x_t = d.q_sample(x0, t=t)
        term, pred = d._loss_term_bpd(fn, x0, x_t=x_t, t=t, clip_denoised=True, return_pred=True)
        torch.testing.assert_close(losses[:, ti], term, rtol=0, atol=0)
        torch.testing.assert_close(mses[:, ti], (pred - x0).pow(2).mean(dim=(1, 2, 3)), rtol=0, atol=0)
    torch.testing.assert_close(total, losses.sum(1) + prior, rtol=0, atol=0)


def test_learned_variance_generic_path():
    """model_var_type="learned" (cannot be constructed upstream): the denoiser's second half of channels is the log-variance."""
    import ddpm_torch_b200 as D
    betas = D.get_beta_schedule("linear", 1e-4, 0.02, 20)
    d = D.GaussianDiffusion(betas, "eps", "learned", "kl")
    fn = R.toy_denoiser(3, 2, seed=6)
    g = torch.Generator().manual_seed(2)
    x0 = torch.rand(2, 3, 8, 8, generator=g) * 2 - 1
    t = torch.tensor([0, 7])
    x_t = d.q_sample(x0, t, noise=torch.randn(2, 3, 8, 8, generator=g))
    mean, var, logvar, pred = d.p_mean_var(fn, x_t, t, clip_denoised=True, return_pred=True)
    out = fn(x_t, t)
    assert torch.equal(logvar, out[:, 3:]) and torch.equal(var, out[:, 3:].exp()) and mean.shape == x0.shape
    assert torch.isfinite(d.train_losses(fn, x0, t)).all()
    with pytest.raises(AssertionError):
        D.GaussianDiffusion(betas, "eps", "learned", "mse").train_losses(fn, x0, t)      # diffusion.py:230


def test_model_wrapper_is_a_generic_denoise_fn():
    """utils/train.py:349-367 / train.py:70-73: PixelUnshuffle -> model -> PixelShuffle; the diffusion classes treat the
    wrapper as a non-native callable (generic torch path)."""
    import ddpm_torch_b200 as D
    inner_fn = R.toy_denoiser(12, 1, seed=8)

    class Inner(torch.nn.Module):
        def forward(self, x, t):
            return inner_fn(x, t)
    w = D.ModelWrapper(Inner(), torch.nn.PixelUnshuffle(2), torch.nn.PixelShuffle(2))
    g = torch.Generator().manual_seed(4)
    x0 = torch.rand(2, 3, 8, 8, generator=g) * 2 - 1
    t = torch.tensor([3, 11]); noise = torch.randn(2, 3, 8, 8, generator=g)
    ref = torch.nn.functional.pixel_shuffle(inner_fn(torch.nn.functional.pixel_unshuffle(x0, 2), t), 2)
    assert torch.equal(w(x0, t), ref) and torch.equal(w(x0, t=t), ref)
    d = D.GaussianDiffusion(D.get_beta_schedule("linear", 1e-4, 0.02, 20), "eps", "fixed-small", "mse")
    rd = R.RefDiffusion(R.get_beta_schedule("linear", 1e-4, 0.02, 20), "fixed-small")
    torch.testing.assert_close(d.train_losses(w, x0, t, noise), rd.train_losses(lambda x, tt: w(x, tt), x0, t, noise), rtol=1e-6, atol=1e-7)
    assert D.ModelWrapper(Inner())(torch.nn.functional.pixel_unshuffle(x0, 2), t).shape == (2, 12, 4, 4)
