"""Parameter update of the training step on the flat buffers.

The reference does, after ``loss.backward()`` (ddpm_torch/utils/train.py:159-165)::

    nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.grad_norm)
    self.optimizer.step()                      # torch.optim.Adam(lr, betas)            train.py:128
    self.optimizer.zero_grad(set_to_none=True)
    self.scheduler.step()                      # LambdaLR: min((t + 1) / warmup, 1)     train.py:130-132
    self.ema.update()                          # utils/train.py:300-305

i.e. ~13 tiny kernels for each of the 304 tensors.  Here the same arithmetic is ONE call into the C ABI
(``ddpm_opt_step``: a sum-of-squares launch + one fused clip/Adam/EMA launch over the flat fp32 buffers, no host sync):

    opt = FusedAdam(model, lr=2e-4, betas=(0.9, 0.999), warmup=5000, grad_norm=1.0, ema=EMA(model, 0.9999))
    ...
    loss.backward(); opt.step()

``EMA`` mirrors the reference class (same attributes, ``apply``/``restore``/context manager, ``state_dict`` layout
``{"decay", "shadow", "num_updates"}``) with the shadow tensors being views into one flat buffer; ``FusedAdam.state_dict()``
has the layout of ``torch.optim.Adam.state_dict()`` so checkpoints written by either side load into the other.
There is no CPU / PyTorch fallback: a model that is not the native sm_100a UNet raises.
"""
import ctypes as C
import math
import weakref

import torch

from . import _lib
from .unet import UNet


def _native(model):
    m = getattr(model, "module", model)          # DDP / DataParallel wrapper
    if not isinstance(m, UNet):
        raise RuntimeError("ddpm_torch_b200.optim works on the native ddpm_torch_b200.UNet only (no generic fallback)")
    if not m.flat_params.is_cuda:
        raise RuntimeError("ddpm_torch_b200.optim needs the model on an sm_100a CUDA device (no CPU fallback)")
    return m


class EMA:
    """utils/train.py:280-345.  ``shadow[name]`` are views into ``self.flat`` (same offsets as the model's flat buffer)."""

    def __init__(self, model, decay=0.9999):
        m = _native(model)
        self.flat = m.flat_params.detach().clone()
        names = [k for k, v in m.named_parameters() if v.requires_grad]
        views = {name: self.flat[off:off + math.prod(shape)].view(shape) for name, shape, off in m._meta}
        self.shadow = {k: views[k] for k in names}
        self._refs = {k: weakref.ref(v) for k, v in m.named_parameters() if v.requires_grad}
        self._model = weakref.ref(m)
        self.decay = decay
        self.num_updates = -1
        self.backup = None

    def update(self):
        """Stand-alone update (utils/train.py:300-305).  Not needed with ``FusedAdam(ema=...)``, which folds it into its step."""
        self.num_updates += 1
        decay = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))
        m = self._model()
        assert m is not None, "referenced object no longer exists!"
        self.flat.add_(m.flat_params.detach() - self.flat, alpha=1 - decay)

    def apply(self):
        m = self._model()
        self.backup = m.flat_params.detach().clone()
        with torch.no_grad():
            m.flat_params.copy_(self.flat)
        m.repack()

    def restore(self):
        m = self._model()
        with torch.no_grad():
            m.flat_params.copy_(self.backup)
        m.repack()
        self.backup = None

    def __enter__(self):
        self.apply()

    def __exit__(self, *exc):
        self.restore()

    def state_dict(self):
        return {"decay": self.decay, "shadow": self.shadow, "num_updates": self.num_updates}

    @property
    def extra_states(self):
        return {"decay", "num_updates"}

    def load_state_dict(self, state_dict, strict=True):
        mine = set(self.shadow).union(self.extra_states)
        theirs = set(state_dict["shadow"]).union(self.extra_states)
        bad = set.symmetric_difference(mine, theirs) if strict else set.difference(mine, theirs)
        if bad:
            raise RuntimeError("Key mismatch!\n"
                               f"Missing key(s): {', '.join(set.difference(mine, theirs))}."
                               f"Unexpected key(s): {', '.join(set.difference(theirs, mine))}")
        with torch.no_grad():
            for k, v in state_dict["shadow"].items():
                if k in self.shadow:
                    self.shadow[k].copy_(v)          # keeps the views into the flat buffer
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]


class FusedAdam:
    """clip_grad_norm_ + Adam + LambdaLR warm-up (+ EMA) over the flat buffers in one native call per step."""

    def __init__(self, model, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, warmup=0, grad_norm=0.0, ema=None):
        m = _native(model)
        self._model = m
        self.base_lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.warmup, self.grad_norm, self.ema = int(warmup), float(grad_norm or 0.0), ema
        dev = m.flat_params.device
        n = m.flat_params.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._state = torch.zeros(16, dtype=torch.float32, device=dev)      # 64 B device scratch, zeroed once
        self.norm_out = torch.zeros(2, dtype=torch.float32, device=dev)     # {total_norm, clip_coef} of the last step
        self.steps = 0                                                     # optimizer steps taken (== scheduler steps)

    @property
    def lr(self):
        """learning rate the NEXT step will use: base_lr * min((t + 1) / warmup, 1)   (train.py:130-132)"""
        return self.base_lr * (min((self.steps + 1) / self.warmup, 1.0) if self.warmup > 0 else 1.0)

    @property
    def param_groups(self):
        # the full key set of torch.optim.Adam's param group, so that a stock Adam can load the state and keep stepping
        return [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "initial_lr": self.base_lr, "params": list(range(len(self._model._params)))}]

    def step(self):
        m = self._model
        if m.flat_grads is None:
            raise RuntimeError("FusedAdam.step(): no gradients yet (run a training forward/backward first)")
        if m.flat_params.data_ptr() % 16 or not m._views_ok():
            raise RuntimeError("FusedAdam.step(): parameters are no longer views of the flat buffer")
        if self.exp_avg.device != m.flat_params.device:
            raise RuntimeError("FusedAdam.step(): model moved to another device after the optimizer was built")
        cfg = _lib.OptCfg()
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = self.lr, self.betas[0], self.betas[1], self.eps
        cfg.max_grad_norm = self.grad_norm
        cfg.step = self.steps + 1
        if self.ema is not None:
            self.ema.num_updates += 1
            cfg.ema_decay, cfg.ema_num_updates = float(self.ema.decay), int(self.ema.num_updates)
        else:
            cfg.ema_decay, cfg.ema_num_updates = -1.0, 0
        with torch.cuda.device(m.flat_params.device):
            _lib.check(_lib.lib().ddpm_opt_step(
                m.flat_params.data_ptr(), m.flat_grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.ema.flat.data_ptr() if self.ema is not None else None, m.flat_params.numel(), C.byref(cfg),
                self._state.data_ptr(), self.norm_out.data_ptr(), _lib.stream_ptr()), "opt_step")
        self.steps += 1
        m.repack()                       # packed bf16 weights are stale now
        m.grads_consumed()               # the next backward starts a fresh accumulation
        return self.norm_out

    def zero_grad(self, set_to_none=True):
        """Drops autograd's ``.grad`` views and ends the current accumulation window of the flat gradient buffer (between
        two ``zero_grad``/``step`` calls successive backward passes ADD UP in ``model.flat_grads``, like ``p.grad`` does
        under the reference's ``--num-accum``)."""
        for p in self._model.parameters():
            p.grad = None
        self._model.grads_consumed()

    # ---- torch.optim.Adam-compatible checkpoint layout
    def state_dict(self):
        m = self._model
        st = {}
        if self.steps > 0:
            for i, (a, b) in enumerate(zip(m.grad_views(self.exp_avg), m.grad_views(self.exp_avg_sq))):
                st[i] = {"step": torch.tensor(float(self.steps)), "exp_avg": a, "exp_avg_sq": b}
        return {"state": st, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        m = self._model
        with torch.no_grad():
            self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            steps = 0
            for i, (a, b) in enumerate(zip(m.grad_views(self.exp_avg), m.grad_views(self.exp_avg_sq))):
                if i in sd["state"]:
                    s = sd["state"][i]
                    a.copy_(s["exp_avg"]); b.copy_(s["exp_avg_sq"]); steps = int(float(s["step"]))
        self.steps = steps
        g = sd["param_groups"][0]
        self.base_lr = float(g.get("initial_lr", g["lr"]))
        self.betas, self.eps = (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])


def hbm_bytes_per_step(n_params, ema=True):
    """algorithmic bytes of one fused update: norm pass 4 B + (g, p, m, v [, shadow]) read + (p, m, v [, shadow]) written"""
    return n_params * (4 + (20 if ema else 16) + (16 if ema else 12))
