"""Sample post-processing on the device.

``generate.py:128-130`` pulls the fp32 NCHW samples to the host and runs five host-side passes::

    x = diffusion.p_sample(...).cpu()
    x = (x * 127.5 + 127.5).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).numpy()

Here the same expression is one HBM-bound kernel (4 B read + 1 B written per element, bit-exact with the reference
arithmetic) that emits the NHWC uint8 layout ``PIL.Image.fromarray`` consumes, and the host copy is 4x smaller and can
be issued asynchronously into pinned memory so that it overlaps the next batch's sampling loop."""
import torch

from . import _lib


def to_uint8_nhwc(x, out=None):
    """x: float32 [B, C<=4, H, W] on an sm_100a device -> uint8 [B, H, W, C] (same device)."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.ndim == 4):
        raise RuntimeError("to_uint8_nhwc: expected a float32 CUDA tensor [B, C, H, W] (no CPU fallback)")
    B, C, H, W = x.shape
    x = x.contiguous()
    if out is None:
        out = torch.empty((B, H, W, C), dtype=torch.uint8, device=x.device)
    elif out.shape != (B, H, W, C) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("to_uint8_nhwc: `out` must be a contiguous uint8 [B, H, W, C] tensor on the same device")
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ddpm_to_uint8_nhwc(x.data_ptr(), out.data_ptr(), B, C, H, W, _lib.stream_ptr()), "to_uint8_nhwc")
    return out


def to_uint8_host_async(x, pinned=None):
    """Device post-processing + asynchronous copy into pinned host memory.  Returns ``(pinned_uint8_nhwc, event)``; wait on
    the event (``event.synchronize()``) before handing ``pinned.numpy()`` to the PNG writers (generate.py:112-114,130)."""
    dev = to_uint8_nhwc(x)
    if pinned is None or pinned.shape != dev.shape:
        pinned = torch.empty(dev.shape, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return pinned, ev
