// Host side of the tcgen05 GEMM engine: tensor-map encoding (driver entry point fetched at run time, so the
// library links without libcuda), tile-geometry selection, kernel dispatch.
#pragma once
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstring>
#include <string>
#include "umma_gemm.cuh"

namespace ddpm {

inline std::string& last_error() { static thread_local std::string e; return e; }
inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    last_error() = buf;
    return code;
}
#define DDPM_CUDA_OK(expr)                                                                         \
    do { cudaError_t e_ = (expr); if (e_ != cudaSuccess)                                           \
        return ddpm::fail(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// bf16 tensor viewed as (C, W, H, N) with channels contiguous, pixel stride ld elements; 128-byte swizzle.
// estride = 2 samples every other pixel in W and H (stride-2 convolutions): the box then spans estride*box_{w,h}
// elements and TMA loads ceil(span/estride) of them.
inline int make_tmap_4d(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, long long ld,
                        int box_c, int box_w, int box_h, int box_n, int estride = 1) {
    auto fn = tmap_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) return fail(-4, "tensor map: pointer/stride not 16-byte aligned");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), (cuuint32_t)box_n};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-5, "cuTensorMapEncodeTiled(4d) failed: %d (C=%d W=%d H=%d N=%d ld=%lld box=%d,%d,%d,%d)",
                                       (int)r, C, W, H, N, ld, box_c, box_w, box_h, box_n);
    return 0;
}
// bf16 matrix stack viewed as (K, rows, batch): row stride ld elements, batch stride bs elements.
inline int make_tmap_3d(CUtensorMap* m, const void* ptr, int K, int rows, int batch, long long ld, long long bs,
                        int box_k, int box_rows) {
    auto fn = tmap_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7) || (bs & 7)) return fail(-4, "tensor map: pointer/stride not 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bs : ld * rows) * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-5, "cuTensorMapEncodeTiled(3d) failed: %d (K=%d rows=%d batch=%d ld=%lld)", (int)r, K, rows, batch, ld);
    return 0;
}

// Box over (W,H,N) covering `rows` consecutive pixels of an NHWC tensor (rows = 128 or 64). Needs power-of-two W,H.
inline bool pick_box(int W, int H, int rows, int& w_t, int& h_t, int& n_t) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!pow2(W) || !pow2(H)) return false;
    if (W >= rows) { w_t = rows; h_t = 1; n_t = 1; return W % rows == 0; }
    w_t = W;
    if (W * H >= rows) { h_t = rows / W; n_t = 1; return H % h_t == 0; }
    h_t = H; n_t = rows / (W * H);
    return n_t <= 256;
}

// A fully-resolved GEMM launch (tensor maps are baked for fixed pointers; re-launchable, graph-capturable).
struct GemmLaunch {
    CUtensorMap a[3], b, o;   // o: output map of the TMA-store epilogue (p.tma_store), else a copy of b
    GemmParams p;
    dim3 grid;
    int mode, block_n;
    int pair;               // KK: CTA-pair kernel (cta_group::2); the B map's box then holds block_n / 2 rows
    double flops;
};
template <int BLOCK_N, int MODE, int STAGES, int KSTEPS>
inline int launch_gemm_pair(const GemmLaunch& g, cudaStream_t st) {
    using SM = GemmSmem<BLOCK_N / 2, STAGES, KSTEPS, (MODE == GEMM_MNMN ? 0 : 1)>;
    auto kern = umma_gemm_kernel<BLOCK_N, MODE, STAGES, KSTEPS, 1>;
    static bool attr_done = false;
    if (!attr_done) { DDPM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL)); attr_done = true; }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    int pairs = (int)g.grid.x / 2; if (pairs > num_sms / 2) pairs = num_sms / 2;
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(gemm_threads(MODE)); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    bool pdl = pdl_enabled();
    if (pdl) { cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone; if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone && !getenv("DDPM_PDL_GRAPH")) pdl = false; }
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    DDPM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g.a[0], g.a[1], g.a[2], g.b, g.o, g.p));
    return 0;
}
template <int BLOCK_N, int MODE, int STAGES, int KSTEPS>
inline int launch_gemm_inst2(const GemmLaunch& g, cudaStream_t st) {
    if constexpr (MODE != GEMM_KMN && BLOCK_N >= 128) { if (g.pair) return launch_gemm_pair<BLOCK_N, MODE, STAGES, KSTEPS>(g, st); }
    using SM = GemmSmem<BLOCK_N, STAGES, KSTEPS, (MODE == GEMM_MNMN ? 0 : 1)>;
    auto kern = umma_gemm_kernel<BLOCK_N, MODE, STAGES, KSTEPS>;
    static bool attr_done = false;
    if (!attr_done) {
        DDPM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
        attr_done = true;
    }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    // persistent: one CTA per SM (two when two fit: <= 96 KB of stages and 2 x 2*BLOCK_N <= 512 TMEM columns)
    const int per_sm = (SM::TOTAL <= 110 * 1024 && 4 * BLOCK_N <= 512) ? 2 : 1;
    if (SM::TOTAL > 232448) return fail(-6, "gemm variant needs %d B of shared memory (> 227 KB)", SM::TOTAL);
    int ctas = (int)g.grid.x; if (ctas > num_sms * per_sm) ctas = num_sms * per_sm;
    launch_k(kern, ctas, gemm_threads(MODE), SM::TOTAL, st, g.a[0], g.a[1], g.a[2], g.b, g.o, g.p);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int BLOCK_N, int MODE>
inline int launch_gemm_inst(const GemmLaunch& g, cudaStream_t st) {
    // One persistent CTA per SM, ~192 KB of stages; TMEM double buffering overlaps the epilogue.
    //   N=256: 4 stages x 1 slab (48 KB);  N=128: 3 stages x 2 slabs (64 KB);  N=64: 4 stages x 2 slabs (48 KB)
    if constexpr (BLOCK_N == 256) return launch_gemm_inst2<BLOCK_N, MODE, 4, 1>(g, st);
    else return launch_gemm_inst2<BLOCK_N, MODE, (BLOCK_N == 128 ? 3 : 4), 2>(g, st);
}

inline int launch_gemm(const GemmLaunch& g, cudaStream_t st) {
#define DDPM_GEMM_CASE(BN, MD) if (g.block_n == BN && g.mode == MD) return launch_gemm_inst<BN, MD>(g, st);
    DDPM_GEMM_CASE(64, GEMM_KK) DDPM_GEMM_CASE(128, GEMM_KK) DDPM_GEMM_CASE(256, GEMM_KK)
    DDPM_GEMM_CASE(64, GEMM_MNMN) DDPM_GEMM_CASE(128, GEMM_MNMN) DDPM_GEMM_CASE(256, GEMM_MNMN)
    DDPM_GEMM_CASE(64, GEMM_KMN) DDPM_GEMM_CASE(128, GEMM_KMN) DDPM_GEMM_CASE(256, GEMM_KMN)
#undef DDPM_GEMM_CASE
    return fail(-6, "unsupported gemm variant block_n=%d mode=%d", g.block_n, g.mode);
}

// K-major GEMMs built with cta_pair = 0 use the CTA-pair kernel when this is non-zero (inference plans; see build_gemm)
inline int& gemm_pair_hint() { static thread_local int v = 1; return v; }

inline int pick_block_n(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64); }

}  // namespace ddpm
