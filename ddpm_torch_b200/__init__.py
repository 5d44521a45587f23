"""ddpm_torch_b200 — B200-native (sm_100a) drop-in for the hot path of tqch/ddpm-torch:
UNet forward/backward inside GaussianDiffusion.train_losses and the p_sample / DDIM loops."""
from . import _lib  # noqa: F401
