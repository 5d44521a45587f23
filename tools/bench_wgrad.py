"""Micro-benchmark of the weight-gradient GEMM (MN-major tcgen05 GEMM, split-K fp32 reds) through the C ABI.
usage: python tools/bench_wgrad.py    (prints us per launch for several shapes / split counts)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_torch_b200 import _lib
import ctypes as C

def bf(*shape, seed=0, scale=1.0):
    g = torch.Generator("cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)

def time_desc(d, reps=20, nbuf=1):
    L = _lib.lib(); sp = _lib.stream_ptr()
    for _ in range(3): _lib.check(L.ddpm_gemm_run(C.byref(d), sp), "gemm")
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): L.ddpm_gemm_run(C.byref(d), sp)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def wgrad(B, H, W, Co, Ci, taps, splits, pair=0):
    dy = bf(B, H, W, Co, seed=1, scale=0.1); a = bf(B, H, W, Ci, seed=2)
    out = torch.zeros(taps, Co, Ci, device="cuda")
    d = _lib.GemmDesc()
    d.mode = 1; d.M = Co; d.N = Ci; d.W = W; d.H = H; d.NB = B
    d.a_ptr[0] = dy.data_ptr(); d.a_C[0] = Co; d.a_ld[0] = Co
    d.b_ptr = a.data_ptr(); d.b_K = Ci; d.b_ld = Ci
    d.taps = taps; d.splits = splits; d.kblocks = B * H * W // 64; d.grid_z = taps * splits
    d.out = out.data_ptr(); d.ldo = Ci; d.out_tap_stride = Co * Ci; d.flags = 3; d.alpha = 1.0
    d.cta_pair = pair
    us = time_desc(d)
    fl = 2.0 * B * H * W * Co * Ci * taps
    return us, fl / us * 1e-6

if __name__ == "__main__":
    for (B, H, W, Co, Ci, taps, name) in [(128, 32, 32, 128, 128, 9, "3x3 128->128 @32"), (128, 32, 32, 128, 256, 1, "1x1 256->128 @32"),
                                          (128, 32, 32, 128, 128, 1, "1x1 128->128 @32"), (128, 16, 16, 256, 256, 9, "3x3 256->256 @16"),
                                          (128, 16, 16, 256, 256, 1, "1x1 256->256 @16"), (128, 8, 8, 256, 256, 9, "3x3 256->256 @8")]:
        bn = 256 if Ci % 256 == 0 else 128
        tiles = ((Co + 127) // 128) * (Ci // bn) * taps
        kb = B * H * W // 64
        for splits in sorted(set([max(1, 148 // tiles), max(1, 74 // tiles), max(1, 37 // tiles), max(1, 296 // tiles)])):
            if splits > kb // 4: continue
            us, tf = wgrad(B, H, W, Co, Ci, taps, splits)
            print(f"{name:20s} tiles {tiles:3d} splits {splits:3d} ctas {tiles*splits:4d}  {us:7.1f} us  {tf:7.0f} TF/s")
