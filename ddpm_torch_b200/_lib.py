"""ctypes binding of libddpm_b200.so (the C ABI declared in include/ddpm_b200.h).

The extension is built in-tree by ``__graft_entry__.build()`` / ``csrc/build.sh``.  There is NO fallback:
if the shared library is missing or the device is not sm_100, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddpm_b200.so")


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
    ]


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
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ddpm_last_error().decode(errors="replace")
        raise RuntimeError(f"ddpm_b200 {what} failed (code {rc}): {msg}")


def runtime_check():
    check(lib().ddpm_runtime_check(), "runtime_check")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
