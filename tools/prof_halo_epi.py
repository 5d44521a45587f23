"""Isolated timing of the haloed 3x3 conv with the GroupNorm epilogue fusions (gn_epilogue.cuh), rotating buffers > L2.
usage: python tools/prof_halo_epi.py [variant ...]   variants: plain qstats   (default: both)
Under ncu pass ONE variant and --once (a single launch after warm-up is bracketed by cudaProfilerStart/Stop)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_torch_b200 import _lib

once = "--once" in sys.argv
variants = [a for a in sys.argv[1:] if not a.startswith("--")] or ["plain_pair", "xf_pair", "xfnomath_pair", "xfnofence_pair"]
B, H, W = 128, 32, 32
L = _lib.lib(); st = _lib.stream_ptr()
for (ci, co) in ((128, 128), (128, 384), (384, 128), (256, 256), (512, 256)):
    nbuf = 6
    xs = [torch.randn(B, H, W, ci, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    ys = [torch.empty(B, H, W, co, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(co, 9 * ci, device="cuda") * 0.03).to(torch.bfloat16); bias = torch.zeros(co, device="cuda")
    qs = torch.zeros(B, co // 4, 2, device="cuda", dtype=torch.float64)
    Kx = torch.randn(B, 4, ci, device="cuda")
    for var in variants:
        ds = []
        for i in range(nbuf):
            d = _lib.HaloDesc(); d.NB, d.H, d.W, d.Cout = B, H, W, co
            d.a_ptr[0] = xs[i].data_ptr(); d.a_C[0] = ci; d.a_ld[0] = ci
            d.nseg = 1; d.seg_map[0] = 0; d.seg_taps[0] = 9; d.seg_kchunks[0] = ci // 64
            d.w = w.data_ptr(); d.ldw = 9 * ci; d.Ktot = 9 * ci; d.out = ys[i].data_ptr()
            d.bias = bias.data_ptr()
            if var.startswith("qstats"):
                d.gn.qstats = qs.data_ptr()
            if var.startswith("xf"):
                d.xf_K = Kx.data_ptr(); d.xf_silu = 1 | (4 if "nomath" in var else 0) | (8 if "nofence" in var else 0)
            if var.endswith("_sub2"):
                d.force_sub = 2
            elif var.endswith("_pair"):
                d.force_sub = 3
            else:
                d.force_sub = 1
            ds.append(d)
        for d in ds:
            _lib.check(L.ddpm_conv_halo_run(C.byref(d), st))
        torch.cuda.synchronize()
        if once:
            torch.cuda.profiler.start()
            L.ddpm_conv_halo_run(C.byref(ds[0]), st)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for i in range(30):
            L.ddpm_conv_halo_run(C.byref(ds[i % nbuf]), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print("halo %d->%d %-9s: %.2f us  %.0f TF/s" % (ci, co, var, ms * 1e3, 2.0 * B * H * W * co * 9 * ci / ms / 1e9), flush=True)
assert L.ddpm_device_error_flag() == 0
