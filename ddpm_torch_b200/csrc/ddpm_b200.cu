// libddpm_b200.so — single translation unit (device-side error flag and kernels share one module).
#include <cstdarg>
#include "gemm_build.cuh"
#include "kernels_simt.cuh"

using namespace ddpm;

extern "C" {

const char* ddpm_last_error(void) { return last_error().c_str(); }

int ddpm_runtime_check(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(-1, "no CUDA device / driver"); }
    cudaDeviceProp pr;
    DDPM_CUDA_OK(cudaGetDeviceProperties(&pr, 0));
    if (pr.major != 10) return fail(-1, "device 0 is sm_%d%d, this library is sm_100a only", pr.major, pr.minor);
    if (!tmap_encode_fn()) return fail(-3, "cuTensorMapEncodeTiled unavailable");
    return 0;
}

int ddpm_device_error_flag(void) {
    unsigned v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&v, g_kernel_error, sizeof v) != cudaSuccess) return -1;
    return (int)v;
}

int ddpm_gemm_run(const ddpm_gemm_desc* d, void* stream) {
    GemmLaunch g;
    int rc = build_gemm(*d, g);
    if (rc) return rc;
    return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
