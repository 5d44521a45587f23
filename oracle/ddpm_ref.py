"""ORACLE — test infrastructure, NOT product code.

A plain-PyTorch fp32 functional restatement of the reference's hot path
(tqch/ddpm-torch @ b60eb8d): the UNet forward, the diffusion coefficient
tables, q_sample / MSE loss, the ancestral and DDIM sampler steps.  It exists
only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg can check and time the path without /root/reference
(which does not exist on the GPU box).  Nothing under ddpm_torch_b200/ may
import this module.

The arithmetic lives in third-party PyTorch (torch>=1.12 per the reference's
README.md:20-22; here torch 2.11.0): F.conv2d, F.linear, F.group_norm, F.silu,
softmax, gather.  This file restates HOW the reference composes those calls;
every function cites the reference lines it follows.

Parity pin: the reference ships no tests / golden vectors, so
this restatement is pinned against outputs of the reference itself, imported
in the build container by ``oracle/gen_golden.py`` and committed under
``tests/golden/`` (tests/test_oracle_golden.py re-checks them on every run).

Differences in FORM (not arithmetic) from the reference: parameters come from
a flat ``state_dict`` (the reference's own key names) instead of nn.Modules;
dropout takes an explicit keep-mask list so tests can inject masks.
"""
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

GN_GROUPS = 32      # unet.py:18-20
GN_EPS = 1e-6       # unet.py:20


# --------------------------------------------------------------------------- config
def normalize_cfg(cfg: dict) -> dict:
    """UNet constructor arguments, unet.py:96-121 (defaults and bool broadcast)."""
    c = dict(cfg)
    c.setdefault("out_channels", c["in_channels"])
    c["ch_multipliers"] = tuple(c["ch_multipliers"])
    levels = len(c["ch_multipliers"])
    aa = c["apply_attn"]
    c["apply_attn"] = tuple([aa] * levels if isinstance(aa, bool) else aa)
    c["time_embedding_dim"] = c.get("time_embedding_dim") or 4 * c["hid_channels"]
    c.setdefault("drop_rate", 0.0)
    c.setdefault("resample_with_conv", True)
    assert c["resample_with_conv"], "oracle covers the conv resampler only (all configs use it)"
    return c


def param_shapes(cfg: dict) -> "Dict[str, tuple]":
    """Ordered {state_dict key: shape} exactly as the reference registers them
    (unet.py:122-142 for the top level, :67-81 ResidualBlock, :29-41 AttentionBlock,
    :156-203 level builders).  Verified against the real module in test_oracle_golden."""
    c = normalize_cfg(cfg)
    ch, mult, nrb = c["hid_channels"], c["ch_multipliers"], c["num_res_blocks"]
    E = c["time_embedding_dim"]
    L = len(mult)
    out: "Dict[str, tuple]" = {}

    def lin(p, i, o):
        out[p + ".weight"] = (o, i)
        out[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        out[p + ".weight"] = (o, i, k, k)
        out[p + ".bias"] = (o,)

    def gn(p, n):
        out[p + ".weight"] = (n,)
        out[p + ".bias"] = (n,)

    def res(p, i, o):
        gn(p + ".norm1", i); conv(p + ".conv1", i, o, 3); lin(p + ".fc", E, o)
        gn(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".skip", i, o, 1)

    def attn(p, n):
        gn(p + ".norm", n); conv(p + ".project_in", n, 3 * n, 1); conv(p + ".project_out", n, n, 1)

    def block(p, i, o, with_attn):
        if with_attn:
            res(p + ".0", i, o); attn(p + ".1", o)
        else:
            res(p, i, o)

    lin("embed.0", ch, E); lin("embed.2", E, E)
    conv("in_conv", c["in_channels"], ch, 3)
    chs = [ch * m for m in mult]
    for i in range(L):
        prev = chs[i - 1] if i else ch
        cur = chs[i]
        p = f"downsamples.level_{i}"
        block(f"{p}.0", prev, cur, c["apply_attn"][i])
        for j in range(1, nrb):
            block(f"{p}.{j}", cur, cur, c["apply_attn"][i])
        if i != L - 1:
            conv(f"{p}.{nrb}.1", cur, cur, 3)
    mid = chs[-1]
    res("middle.0", mid, mid); attn("middle.1", mid); res("middle.2", mid, mid)
    for i in range(L):
        nxt = ch if i == 0 else chs[i - 1]
        prev = chs[-1] if i == L - 1 else chs[i + 1]
        cur = chs[i]
        p = f"upsamples.level_{i}"
        block(f"{p}.0", prev + cur, cur, c["apply_attn"][i])
        for j in range(1, nrb):
            block(f"{p}.{j}", 2 * cur, cur, c["apply_attn"][i])
        block(f"{p}.{nrb}", nxt + cur, cur, c["apply_attn"][i])
        if i != 0:
            conv(f"{p}.{nrb + 1}.1", cur, cur, 3)
    gn("out_conv.0", ch); conv("out_conv.2", ch, c["out_channels"], 3)
    return out


def fill_params(shapes: "Dict[str, tuple]", seed: int, scale: float = 1.0) -> "Dict[str, torch.Tensor]":
    """Deterministic non-degenerate values for an ordered {key: shape} dict: every tensor
    is drawn from its own CPU generator seeded by (seed, index).
    conv/linear weights ~ U(-a, a), a = scale*sqrt(3/fan_in) (unit gain);
    biases ~ 0.1*N(0,1); GroupNorm weight ~ 1 + 0.1*N(0,1), bias ~ 0.1*N(0,1)."""
    sd = {}
    for idx, (k, shp) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        is_norm = "norm" in k or k.startswith("out_conv.0")
        if k.endswith(".weight") and not is_norm:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            a = scale * math.sqrt(3.0 / fan_in)
            v = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * a
        elif k.endswith(".weight"):
            v = 1.0 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float32)
        else:
            v = 0.1 * torch.randn(shp, generator=g, dtype=torch.float32)
        sd[k] = v
    return sd


def make_state_dict(cfg: dict, seed: int, scale: float = 1.0) -> "Dict[str, torch.Tensor]":
    """Deterministic NON-degenerate weights for parity tests.

    The reference's own init zeroes every block's last conv (init_scale=0 ->
    gain sqrt(1e-10), modules.py:18; unet.py:37,79,141) so a fresh model outputs
    ~0.  Tests therefore use fill_params() instead, so the weights
    can be regenerated bit-identically on the GPU box without shipping 143 MB."""
    return fill_params(param_shapes(cfg), seed, scale)


def res_block_shapes(cin, cout, E, p="blk"):
    """ResidualBlock parameters in registration order (unet.py:67-81)."""
    o = {}
    o[p + ".norm1.weight"] = (cin,); o[p + ".norm1.bias"] = (cin,)
    o[p + ".conv1.weight"] = (cout, cin, 3, 3); o[p + ".conv1.bias"] = (cout,)
    o[p + ".fc.weight"] = (cout, E); o[p + ".fc.bias"] = (cout,)
    o[p + ".norm2.weight"] = (cout,); o[p + ".norm2.bias"] = (cout,)
    o[p + ".conv2.weight"] = (cout, cout, 3, 3); o[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        o[p + ".skip.weight"] = (cout, cin, 1, 1); o[p + ".skip.bias"] = (cout,)
    return o


def attn_block_shapes(c, p="blk"):
    """AttentionBlock parameters in registration order (unet.py:29-41)."""
    o = {}
    o[p + ".norm.weight"] = (c,); o[p + ".norm.bias"] = (c,)
    o[p + ".project_in.weight"] = (3 * c, c, 1, 1); o[p + ".project_in.bias"] = (3 * c,)
    o[p + ".project_out.weight"] = (c, c, 1, 1); o[p + ".project_out.bias"] = (c,)
    return o


# --------------------------------------------------------------------------- UNet pieces
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """functions.py:10-26: [sin(t*f_i) | cos(t*f_i)], f_i = exp(-i*ln(1e4)/(half-1))."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    f = torch.exp(-torch.arange(half, dtype=torch.float32, device=t.device) * k)
    ang = torch.outer(t.ravel().to(torch.float32), f)
    e = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, [0, 1])
    return e


def _gn(sd, p, x):
    return F.group_norm(x, GN_GROUPS, sd[p + ".weight"], sd[p + ".bias"], GN_EPS)


def _conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def res_block(sd, p, x, temb, keep_mask=None, drop_rate=0.0):
    """unet.py:83-89.  keep_mask (same shape as conv1 output, 0/1) emulates
    nn.Dropout(p, inplace=True) in train mode: y*mask/(1-p)."""
    skip = _conv(sd, p + ".skip", x) if (p + ".skip.weight") in sd else x
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x)), padding=1)
    h = h + F.linear(F.silu(temb), sd[p + ".fc.weight"], sd[p + ".fc.bias"])[:, :, None, None]
    h = F.silu(_gn(sd, p + ".norm2", h))
    if keep_mask is not None:
        h = h * keep_mask / (1.0 - drop_rate)
    h = _conv(sd, p + ".conv2", h, padding=1)
    return h + skip


def attn_block(sd, p, x):
    """unet.py:43-60: single head, d=C, softmax over keys, scale 1/sqrt(C)."""
    B, C, H, W = x.shape
    qkv = _conv(sd, p + ".project_in", _gn(sd, p + ".norm", x))
    q, k, v = qkv.chunk(3, dim=1)
    q, k, v = (z.reshape(B, C, H * W) for z in (q, k, v))
    w = torch.einsum("bci,bcj->bij", q, k) / math.sqrt(C)
    w = torch.softmax(w, dim=-1)
    o = torch.einsum("bij,bcj->bci", w, v).reshape(B, C, H, W)
    return _conv(sd, p + ".project_out", o) + x


def _block(sd, p, x, temb, with_attn, masks, drop_rate):
    km = masks.pop(0) if masks is not None else None
    if with_attn:
        return attn_block(sd, p + ".1", res_block(sd, p + ".0", x, temb, km, drop_rate))
    return res_block(sd, p, x, temb, km, drop_rate)


def unet_forward(sd, cfg, x, t, keep_masks: Optional[List[torch.Tensor]] = None):
    """unet.py:205-233.  keep_masks: one 0/1 tensor per ResidualBlock in call order
    (None → eval mode / drop_rate 0)."""
    c = normalize_cfg(cfg)
    ch, mult, nrb = c["hid_channels"], c["ch_multipliers"], c["num_res_blocks"]
    L, aa, dr = len(mult), c["apply_attn"], c["drop_rate"]
    masks = list(keep_masks) if keep_masks is not None else None
    temb = timestep_embedding(t, ch)
    temb = F.linear(temb, sd["embed.0.weight"], sd["embed.0.bias"])
    temb = F.linear(F.silu(temb), sd["embed.2.weight"], sd["embed.2.bias"])
    hs = [_conv(sd, "in_conv", x, padding=1)]
    for i in range(L):
        p = f"downsamples.level_{i}"
        for j in range(nrb):
            hs.append(_block(sd, f"{p}.{j}", hs[-1], temb, aa[i], masks, dr))
        if i != L - 1:
            # SamePad2d(3,2) on even H: pad bottom/right by one (modules.py:153-160), then stride 2
            hs.append(_conv(sd, f"{p}.{nrb}.1", F.pad(hs[-1], (0, 1, 0, 1)), stride=2))
    h = hs[-1]
    h = res_block(sd, "middle.0", h, temb, masks.pop(0) if masks is not None else None, dr)
    h = attn_block(sd, "middle.1", h)
    h = res_block(sd, "middle.2", h, temb, masks.pop(0) if masks is not None else None, dr)
    for i in range(L - 1, -1, -1):
        p = f"upsamples.level_{i}"
        for j in range(nrb + 1):
            h = _block(sd, f"{p}.{j}", torch.cat([h, hs.pop()], dim=1), temb, aa[i], masks, dr)
        if i != 0:
            h = F.interpolate(h, scale_factor=2, mode="nearest")       # unet.py:199
            h = _conv(sd, f"{p}.{nrb + 1}.1", h, padding=1)
    h = F.silu(_gn(sd, "out_conv.0", h))
    return _conv(sd, "out_conv.2", h, padding=1)


def num_res_blocks_total(cfg) -> int:
    c = normalize_cfg(cfg)
    L, nrb = len(c["ch_multipliers"]), c["num_res_blocks"]
    return L * nrb + 2 + L * (nrb + 1)


def fwd_flops_per_image(cfg, H, W) -> float:
    """2*MAC of convs + linears + the two attention matmuls."""
    c = normalize_cfg(cfg)
    shapes = param_shapes(c)
    ch, mult, nrb = c["hid_channels"], c["ch_multipliers"], c["num_res_blocks"]
    L, aa = len(mult), c["apply_attn"]
    # resolution at which each conv's OUTPUT lives
    res = {}
    r = (H, W)
    res["in_conv"] = r
    for i in range(L):
        for j in range(nrb):
            res[f"downsamples.level_{i}.{j}"] = r
        if i != L - 1:
            r = (r[0] // 2, r[1] // 2)
            res[f"downsamples.level_{i}.{nrb}"] = r
    for m in ("middle.0", "middle.1", "middle.2"):
        res[m] = r
    for i in range(L - 1, -1, -1):
        for j in range(nrb + 1):
            res[f"upsamples.level_{i}.{j}"] = r
        if i != 0:
            r = (r[0] * 2, r[1] * 2)
            res[f"upsamples.level_{i}.{nrb + 1}"] = r
    res["out_conv"] = r
    fl = 0.0
    for k, shp in shapes.items():
        if not k.endswith(".weight"):
            continue
        if len(shp) == 2:
            fl += 2.0 * shp[0] * shp[1]
        elif len(shp) == 4:
            key = max((q for q in res if k.startswith(q + ".")), key=len)
            h, w = res[key]
            fl += 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * h * w
            if k.endswith("project_in.weight"):
                n, cc = h * w, shp[1]
                fl += 2 * (2.0 * n * n * cc)
    return fl


# --------------------------------------------------------------------------- diffusion
def get_beta_schedule(beta_schedule, beta_start, beta_end, timesteps):
    """diffusion.py:13-29 (fp64)."""
    dt = torch.float64
    if beta_schedule == "quad":
        b = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps, dtype=dt) ** 2
    elif beta_schedule == "linear":
        b = torch.linspace(beta_start, beta_end, timesteps, dtype=dt)
    elif beta_schedule in ("warmup10", "warmup50"):
        frac = 0.1 if beta_schedule == "warmup10" else 0.5
        b = beta_end * torch.ones(timesteps, dtype=dt)
        n = int(timesteps * frac)
        b[:n] = torch.linspace(beta_start, beta_end, n, dtype=dt)
    elif beta_schedule == "const":
        b = beta_end * torch.ones(timesteps, dtype=dt)
    elif beta_schedule == "jsd":
        b = 1.0 / torch.linspace(timesteps, 1, timesteps, dtype=dt)
    else:
        raise NotImplementedError(beta_schedule)
    return b


class RefDiffusion:
    """Coefficient tables + eps-prediction / fixed-variance paths of
    GaussianDiffusion (diffusion.py:34-73, 92-105, 107-158, 217-243).
    With ``subsequence``/``eta`` it re-derives the tables as DDIM does (ddim.py:48-94)."""

    def __init__(self, betas, model_var_type="fixed-large", eta=None, subsequence=None):
        assert betas.dtype == torch.float64
        self.model_var_type = model_var_type
        ab = torch.cumprod(1 - betas, dim=0)
        one = torch.ones(1, dtype=torch.float64)
        if subsequence is None:
            alphas = 1 - betas
            ab_prev = torch.cat([one, ab[:-1]])
            post_var = betas * (1 - ab_prev) / (1 - ab)
            logvar_clip = torch.log(torch.cat([post_var[[1]], post_var[1:]]))
            c1 = betas * torch.sqrt(ab_prev) / (1 - ab)
            c2 = torch.sqrt(alphas) * (1 - ab_prev) / (1 - ab)
            large_logvar = torch.log(torch.cat([post_var[[1]], betas[1:]]))
        else:
            eta2 = eta ** 2
            if eta2 != 1.0 and model_var_type != "fixed-small":
                self.model_var_type = "fixed-small"                      # ddim.py:54-59
            ab = ab[subsequence]
            ab_prev = torch.cat([one, ab[:-1]])
            alphas = ab / ab_prev
            betas = 1 - alphas
            post_var = betas * (1 - ab_prev) / (1 - ab) * eta2
            logvar_clip = torch.log(torch.cat([post_var[[1]], post_var[1:]]).clip(min=1e-20))
            c2 = torch.sqrt(1 - ab - eta2 * betas) * torch.sqrt(1 - ab_prev) / (1 - ab)
            c1 = torch.sqrt(ab_prev) * (1 - torch.sqrt(alphas) * c2)
            large_logvar = torch.log(torch.cat([post_var[[1]], betas[1:]]).clip(min=1e-20))
        self.betas = betas
        self.timesteps = len(betas)
        self.alphas_bar = ab
        self.sqrt_alphas_bar = torch.sqrt(ab)
        self.sqrt_one_minus_alphas_bar = torch.sqrt(1 - ab)
        self.sqrt_recip_alphas_bar = torch.sqrt(1 / ab)
        self.sqrt_recip_m1_alphas_bar = torch.sqrt(1 / ab - 1)
        self.posterior_var = post_var
        self.posterior_logvar_clipped = logvar_clip
        self.posterior_mean_coef1 = c1
        self.posterior_mean_coef2 = c2
        self.fixed_model_logvar = large_logvar if self.model_var_type == "fixed-large" else logvar_clip
        self.subsequence = None if subsequence is None else torch.as_tensor(subsequence)

    @staticmethod
    def _extract(arr, t, x):
        """diffusion.py:75-84: fp64 table → x.dtype, gather, reshape [B,1,1,1]."""
        out = torch.as_tensor(arr, dtype=x.dtype, device=x.device).gather(0, t)
        return out.reshape((-1,) + (1,) * (x.ndim - 1))

    def q_sample(self, x0, t, noise):
        """diffusion.py:92-97."""
        return self._extract(self.sqrt_alphas_bar, t, x0) * x0 + \
            self._extract(self.sqrt_one_minus_alphas_bar, t, x0) * noise

    def train_losses(self, denoise_fn, x0, t, noise):
        """diffusion.py:217-243, loss_type mse / mean_type eps; flat_mean functions.py:99-101."""
        x_t = self.q_sample(x0, t, noise)
        out = denoise_fn(x_t, t)
        return ((noise - out) ** 2).mean(dim=[1, 2, 3])

    def p_sample_step(self, denoise_fn, x_t, t, noise):
        """diffusion.py:107-158 (eps, fixed var, clip_denoised=True); ``noise`` is the
        normal_ draw of :155 supplied by the caller."""
        t_model = t if self.subsequence is None else self.subsequence.to(t.device).gather(0, t)  # ddim.py:101
        eps = denoise_fn(x_t, t_model)
        logvar = self._extract(self.fixed_model_logvar, t, x_t)
        x0 = self._extract(self.sqrt_recip_alphas_bar, t, x_t) * x_t - \
            self._extract(self.sqrt_recip_m1_alphas_bar, t, x_t) * eps
        x0 = x0.clamp(-1.0, 1.0)
        mean = self._extract(self.posterior_mean_coef1, t, x0) * x0 + \
            self._extract(self.posterior_mean_coef2, t, x0) * x_t
        nz = (t > 0).reshape((-1,) + (1,) * (x_t.ndim - 1)).to(x_t)
        return mean + nz * torch.exp(0.5 * logvar) * noise

    def p_sample(self, denoise_fn, x_T, noises: Sequence[torch.Tensor]):
        """diffusion.py:160-174 / ddim.py:96-113 with the per-step normal_ draws
        supplied (noises[k] is used at the k-th executed step, i.e. ti = T-1-k)."""
        x = x_T
        B = x.shape[0]
        for k, ti in enumerate(range(self.timesteps - 1, -1, -1)):
            t = torch.full((B,), ti, dtype=torch.int64, device=x.device)
            x = self.p_sample_step(denoise_fn, x, t, noises[k])
        return x

    def coef_table(self) -> torch.Tensor:
        """[T,5] fp32: (sqrt_recip_ab, sqrt_recip_m1_ab, post_c1, post_c2, exp(0.5*logvar) as the
        reference evaluates it: exp(0.5*float32(logvar)))."""
        f = lambda a: torch.as_tensor(a, dtype=torch.float32)
        sig = torch.exp(0.5 * f(self.fixed_model_logvar))
        return torch.stack([f(self.sqrt_recip_alphas_bar), f(self.sqrt_recip_m1_alphas_bar),
                            f(self.posterior_mean_coef1), f(self.posterior_mean_coef2), sig], dim=1)


def get_selection_schedule(schedule, size, timesteps):
    """ddim.py:30-44."""
    assert schedule in {"linear", "quadratic"}
    if schedule == "linear":
        return torch.arange(0, timesteps, timesteps // size)
    return torch.pow(torch.linspace(0, math.sqrt(timesteps * 0.8), size), 2).round().to(torch.int64)


CIFAR10_CFG = dict(in_channels=3, hid_channels=128, out_channels=3, ch_multipliers=(1, 2, 2, 2),
                   num_res_blocks=2, apply_attn=(False, True, False, False), drop_rate=0.1)
CELEBAHQ_CFG = dict(in_channels=3, hid_channels=128, out_channels=3, ch_multipliers=(1, 1, 2, 2, 4, 4),
                    num_res_blocks=2, apply_attn=(False, False, False, False, True, False), drop_rate=0.0)
TINY_CFG = dict(in_channels=3, hid_channels=32, out_channels=3, ch_multipliers=(1, 2),
                num_res_blocks=1, apply_attn=(False, True), drop_rate=0.0)
# smallest config whose every channel count is a multiple of 64 (tensor-core path eligible)
SMALL64_CFG = dict(in_channels=3, hid_channels=64, out_channels=3, ch_multipliers=(1, 2),
                   num_res_blocks=1, apply_attn=(False, True), drop_rate=0.0)


def to_uint8_nhwc(x):
    """generate.py:129, verbatim: fp32 NCHW samples in [-1, 1] -> uint8 NHWC images."""
    return (x * 127.5 + 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)


def toy_denoiser(C, out_mult, seed):
    """A small deterministic NON-native denoise_fn (3x3 conv + timestep-dependent offset) shared by oracle/gen_golden.py and the
    tests of the generic (bits-per-dim / learned-variance / x_0- and mean-prediction) paths; out_mult = 2 for "learned" variance."""
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(out_mult * C, C, 3, 3, generator=g) * 0.2

    def fn(x, t):
        tb = torch.sin(t.to(torch.float32) * 0.37)[:, None, None, None]
        return F.conv2d(x, W.to(x), padding=1) * 0.5 + 0.1 * tb
    return fn
