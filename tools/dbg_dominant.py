"""Time the dominant conv (3x3, 128->128, 32x32, B=128) through the generic KK engine and through the haloed kernel."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from ddpm_torch_b200 import _lib
r = bench.dominant_kernel_roofline(bench.peaks(), iters=30)
print("generic KK  dbg=%s shallow=%s: ms %.4f TF/s %.0f" % (os.environ.get("DDPM_GEMM_DBG", "0"), os.environ.get("DDPM_GEMM_SHALLOW"), r["ms_per_launch"], r["achieved"]))
if os.environ.get("DDPM_GEMM_DBG", "0") == "0":
    B, H, W, Ci, Co = 128, 32, 32, 128, 128
    for (ci, co) in ((128, 128), (256, 256)):
        for sub in (1, 2):
            nbuf = 8
            xs = [torch.randn(B, H, W, ci, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
            ys = [torch.empty(B, H, W, co, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
            w = (torch.randn(co, 9 * ci, device="cuda") * 0.03).to(torch.bfloat16); bias = torch.zeros(co, device="cuda")
            ds = []
            for i in range(nbuf):
                d = _lib.HaloDesc(); d.NB, d.H, d.W, d.Cout = B, H, W, co
                d.a_ptr[0] = xs[i].data_ptr(); d.a_C[0] = ci; d.a_ld[0] = ci
                d.nseg = 1; d.seg_map[0] = 0; d.seg_taps[0] = 9; d.seg_kchunks[0] = ci // 64
                d.w = w.data_ptr(); d.ldw = 9 * ci; d.Ktot = 9 * ci; d.out = ys[i].data_ptr(); d.bias = bias.data_ptr(); d.force_sub = sub
                ds.append(d)
            L = _lib.lib(); st = _lib.stream_ptr()
            for d in ds: _lib.check(L.ddpm_conv_halo_run(C.byref(d), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
            for i in range(30): L.ddpm_conv_halo_run(C.byref(ds[i % nbuf]), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 30
            print("halo %d->%d sub=%d: ms %.4f TF/s %.0f" % (ci, co, sub, ms, 2.0 * B * H * W * co * 9 * ci / ms / 1e9))
