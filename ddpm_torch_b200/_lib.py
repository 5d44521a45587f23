"""ctypes binding of libddpm_b200.so (the C ABI declared in include/ddpm_b200.h).

The extension is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.sh``.  There is NO fallback:
if the shared library is missing or the device is not sm_100, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddpm_b200.so")


class GnEpi(C.Structure):
    _fields_ = [("qstats", C.c_void_p)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("block_n", C.c_int), ("M", C.c_int), ("N", C.c_int),
        ("W", C.c_int), ("H", C.c_int), ("NB", C.c_int),
        ("a_ptr", C.c_void_p * 3), ("a_C", C.c_int * 3), ("a_ld", C.c_longlong * 3),
        ("nseg", C.c_int), ("seg_map", C.c_int * 3), ("seg_taps", C.c_int * 3),
        ("seg_kchunks", C.c_int * 3), ("seg_cbase", C.c_int * 3),
        ("b_ptr", C.c_void_p), ("b_K", C.c_int), ("b_rows", C.c_int), ("b_batch", C.c_int),
        ("b_ld", C.c_longlong), ("b_bs", C.c_longlong),
        ("b_k_base", C.c_int), ("a_z_n", C.c_int), ("b_z", C.c_int),
        ("taps", C.c_int), ("splits", C.c_int), ("kblocks", C.c_int),
        ("a_c_base", C.c_int), ("b_c_base", C.c_int), ("grid_z", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_int), ("out_z_stride", C.c_longlong),
        ("out_tap_stride", C.c_longlong), ("flags", C.c_int),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int), ("rows_per_vec", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_int), ("alpha", C.c_float),
        ("a_estride", C.c_int), ("b_estride", C.c_int), ("b_pad", C.c_int),
        ("seg_custom", C.c_int * 3), ("seg_cmul", C.c_int * 3),
        ("seg_dx", (C.c_byte * 9) * 3), ("seg_dy", (C.c_byte * 9) * 3),
        ("o_mul", C.c_int), ("o_py", C.c_int), ("o_px", C.c_int), ("kk_splits", C.c_int),
        ("gn", GnEpi), ("cta_pair", C.c_int),
    ]


class HaloDesc(C.Structure):
    _fields_ = [
        ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cout", C.c_int),
        ("a_ptr", C.c_void_p * 3), ("a_C", C.c_int * 3), ("a_ld", C.c_longlong * 3),
        ("nseg", C.c_int), ("seg_map", C.c_int * 3), ("seg_taps", C.c_int * 3), ("seg_kchunks", C.c_int * 3), ("seg_cbase", C.c_int * 3),
        ("w", C.c_void_p), ("ldw", C.c_longlong), ("Ktot", C.c_int),
        ("out", C.c_void_p), ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int), ("residual", C.c_void_p),
        ("base_offset_mode", C.c_int), ("force_sub", C.c_int),
        ("gn", GnEpi), ("xf_K", C.c_void_p), ("xf_silu", C.c_int),
    ]


class UnetCfg(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("hid_channels", C.c_int), ("out_channels", C.c_int),
                ("levels", C.c_int), ("ch_mult", C.c_int * 8), ("num_res_blocks", C.c_int), ("attn", C.c_int * 8),
                ("temb_dim", C.c_int), ("drop_rate", C.c_float)]


class OptCfg(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("max_grad_norm", C.c_double), ("ema_decay", C.c_double), ("step", C.c_int), ("ema_num_updates", C.c_int)]


_lib = None


def lib():
    """Load the shared library once.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(ddpm_torch_b200 has no CPU/PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        L.ddpm_last_error.restype = C.c_char_p
        L.ddpm_runtime_check.restype = C.c_int
        L.ddpm_device_error_flag.restype = C.c_int
        L.ddpm_gemm_run.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        L.ddpm_gemm_run.restype = C.c_int
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_uint64
        sig = {
            "ddpm_conv_halo_run": ([C.POINTER(HaloDesc), vp], i32),
            "ddpm_attn_fused_run": ([vp, vp, vp, i32, i32, i32, vp], i32),
            "ddpm_attn_fused_bwd_run": ([vp, vp, vp, vp, vp, i32, i32, i32, vp], i32),
            "ddpm_unet_create": ([C.POINTER(UnetCfg), C.POINTER(vp)], i32),
            "ddpm_unet_destroy": ([vp], None),
            "ddpm_unet_num_params": ([vp], i32),
            "ddpm_unet_param_info": ([vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i32 * 4), C.POINTER(i64)], i32),
            "ddpm_unet_flat_elems": ([vp], i64),
            "ddpm_unet_workspace_bytes": ([vp, i32, i32, i32, i32], i64),
            "ddpm_unet_plan": ([vp, i32, i32, i32, i32, vp, vp, vp, i64], i32),
            "ddpm_unet_repack": ([vp, vp], i32),
            "ddpm_unet_forward": ([vp, vp, vp, vp, u64, vp], i32),
            "ddpm_unet_backward": ([vp, vp, vp], i32),
            "ddpm_train_forward": ([vp, vp, vp, vp, vp, vp, vp, u64, vp], i32),
            "ddpm_train_backward": ([vp, vp, vp], i32),
            "ddpm_sampler_setup": ([vp, i32, vp, vp], i32),
            "ddpm_sampler_reset": ([vp, i32, vp], i32),
            "ddpm_sampler_step": ([vp, vp, vp, u64, vp], i32),
            "ddpm_sampler_step_pred": ([vp, vp, vp, u64, vp, vp], i32),
            "ddpm_unet_grad_chunks": ([vp, i32, C.POINTER(i64), C.POINTER(i64)], i32),
            "ddpm_unet_wait_grad_chunk": ([vp, i32, vp], i32),
            "ddpm_unet_plan_stats": ([vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)], i32),
            "ddpm_unet_launches_per_forward": ([vp], i32),
            "ddpm_unet_launch_counts": ([vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
            "ddpm_opt_step": ([vp, vp, vp, vp, vp, i64, C.POINTER(OptCfg), vp, vp, vp], i32),
            "ddpm_to_uint8_nhwc": ([vp, vp, i32, i32, i32, i32, vp], i32),
        }
        for name, (args, res) in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


EXPORTS = ["ddpm_last_error", "ddpm_runtime_check", "ddpm_device_error_flag", "ddpm_gemm_run", "ddpm_conv_halo_run", "ddpm_attn_fused_run", "ddpm_attn_fused_bwd_run",
           "ddpm_unet_create", "ddpm_unet_destroy", "ddpm_unet_num_params", "ddpm_unet_param_info",
           "ddpm_unet_flat_elems", "ddpm_unet_workspace_bytes", "ddpm_unet_plan", "ddpm_unet_repack",
           "ddpm_unet_forward", "ddpm_unet_backward", "ddpm_train_forward", "ddpm_train_backward",
           "ddpm_sampler_setup", "ddpm_sampler_reset", "ddpm_sampler_step", "ddpm_sampler_step_pred", "ddpm_unet_grad_chunks", "ddpm_unet_wait_grad_chunk", "ddpm_unet_plan_stats",
           "ddpm_unet_launches_per_forward", "ddpm_unet_launch_counts", "ddpm_opt_step", "ddpm_to_uint8_nhwc"]


def check(rc, what=""):
    if rc != 0:
        msg = lib().ddpm_last_error().decode(errors="replace")
        raise RuntimeError(f"ddpm_b200 {what} failed (code {rc}): {msg}")


def runtime_check():
    check(lib().ddpm_runtime_check(), "runtime_check")


def stream_ptr(device=None):
    """Raw handle of torch's current stream ON ``device`` (default: the current device).  Engine calls of a model that
    lives on cuda:N must be issued on cuda:N's stream under a ``torch.cuda.device(N)`` guard (generate.py:59 builds
    ``cuda:{rank}`` models without ever calling set_device)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
