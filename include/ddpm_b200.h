/* ddpm_b200 — C ABI of the B200-native DDPM engine (libddpm_b200.so).
 *
 * The reference (tqch/ddpm-torch @ b60eb8d) is pure Python on PyTorch and has NO FFI / plugin interface;
 * its seam for this path is the Python callable protocol  denoise_fn(x_t f32[B,C,H,W], t i64[B]) -> f32[B,C,H,W]
 * (ddpm_torch/diffusion.py:109,238; ddim.py:101) plus the GaussianDiffusion / DDIM methods that call it.
 * This header is what a ctypes binding on the reference side binds instead; every entry
 * point cites the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise; the caller
 *    (PyTorch's caching allocator on the Python side) owns every buffer; the library allocates nothing per call,
 *    never synchronises, and is CUDA-graph capturable; `stream` is a cudaStream_t passed as void*.
 *  - every function returns 0 on success, <0 on error; ddpm_last_error() returns the message (thread local).
 *    No exceptions or aborts cross the ABI (the reference raises Python exceptions: diffusion.py:42-43,119,133).
 *  - activations inside the engine are NHWC bf16; the 3-channel network input/output stay NCHW fp32 exactly
 *    as the reference hands them over (unet.py:205,232).
 */
#ifndef DDPM_B200_H
#define DDPM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* ddpm_last_error(void);
/* 0 when the library was built for sm_100a and a CUDA driver + sm_100 device are present. */
int ddpm_runtime_check(void);
/* Device-side error flag set by a bounded wait that timed out inside a kernel (0 = none). Synchronises. */
int ddpm_device_error_flag(void);

/* ------------------------------------------------------------------------------------------------
 * Low-level operator: one tcgen05 implicit-GEMM launch (used by the parity tests of the engine itself).
 * mode 0 (KK)   A K-major NHWC pixel tiles (+3x3 taps, up to 3 channel segments), B K-major matrix stack
 * mode 1 (MNMN) A,B MN-major, K over pixels (conv wgrad / P^T.dO); split-K with fp32 atomics
 * mode 2 (KMN)  A K-major, B MN-major (P.V)
 * replaces: F.conv2d (modules.py:121-123), F.linear (modules.py:59), torch.einsum x2 (unet.py:46-51) and their
 * autograd backward formulas.
 */
typedef struct ddpm_gemm_desc {
    int mode, block_n;                 /* block_n 0 = auto (64/128/256) */
    int M, N;                          /* output rows / cols per z slice */
    int W, H, NB;                      /* geometry of the NHWC tensors behind the 4-D maps (plain matrix: W=rows,H=1) */
    const void* a_ptr[3]; int a_C[3]; long long a_ld[3];   /* up to 3 A tensors: channels extent, pixel stride (elements) */
    int nseg; int seg_map[3], seg_taps[3], seg_kchunks[3], seg_cbase[3];
    const void* b_ptr; int b_K, b_rows, b_batch; long long b_ld, b_bs;  /* KK: [b_batch][b_rows][b_K]; MN modes: b_K = channel extent, b_ld = pixel stride */
    int b_k_base, a_z_n, b_z;
    int taps, splits, kblocks, a_c_base, b_c_base, grid_z;
    void* out; int ldo; long long out_z_stride, out_tap_stride; int flags;   /* flags: 1 = fp32 out, 2 = atomic add (fp32) */
    const float* bias; const float* rowvec; int rowvec_ld, rows_per_vec;
    const void* residual; int ldr; float alpha;
} ddpm_gemm_desc;
int ddpm_gemm_run(const ddpm_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDPM_B200_H */
