"""Time plain K-major GEMMs of the 1x1-conv / attention shapes through the generic tcgen05 engine (DDPM_GEMM_DBG knobs:
1 = no epilogue math/stores, 2 = no MMA issue, 4 = no TMA loads)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_torch_b200 import _lib
L = _lib.lib(); _lib.runtime_check()
shapes = [(32768, 256, 768, "project_in 16x16 B=128"), (32768, 256, 256, "project_out"), (131072, 256, 128, "skip 1x1 L0"),
          (32768, 2304, 256, "3x3 as plain K=2304")]
for (M, K, N, what) in shapes:
    nbuf = 6
    As = [(torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(nbuf)]
    Bm = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    bias = torch.zeros(N, device="cuda")
    ds = []
    for i in range(nbuf):
        d = _lib.GemmDesc()
        d.mode = 0; d.M = M; d.N = N; d.W = M; d.H = 1; d.NB = 1
        d.a_ptr[0] = As[i].data_ptr(); d.a_C[0] = K; d.a_ld[0] = K
        d.nseg = 1; d.seg_map[0] = 0; d.seg_taps[0] = 1; d.seg_kchunks[0] = K // 64; d.seg_cbase[0] = 0
        d.b_ptr = Bm.data_ptr(); d.b_K = K; d.b_rows = N; d.b_batch = 1; d.b_ld = K; d.b_bs = 0
        d.out = outs[i].data_ptr(); d.ldo = N; d.bias = bias.data_ptr(); d.alpha = 1.0
        ds.append(d)
    st = _lib.stream_ptr()
    for d in ds: _lib.check(L.ddpm_gemm_run(C.byref(d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for i in range(30): L.ddpm_gemm_run(C.byref(ds[i % nbuf]), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print("%-26s M=%d K=%d N=%d: %.1f us  %.0f TF/s  (in+out %.0f MB -> %.2f TB/s)" % (what, M, K, N, ms * 1e3, 2.0 * M * K * N / ms / 1e9,
          (M * K + M * N) * 2 / 1e6, (M * K + M * N) * 2 / ms / 1e9))
