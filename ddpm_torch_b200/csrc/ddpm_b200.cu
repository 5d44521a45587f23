// libddpm_b200.so — single translation unit (device-side error flag and kernels share one module).
#include <cstdarg>
#include <new>
#include <cmath>
#include <initializer_list>
#include "unet_engine_impl.cuh"
#include "optim.cuh"

using namespace ddpm;

struct ddpm_unet {
    UnetEngine e;
    // sampler tables (device, owned by the handle; allocated by ddpm_sampler_setup, never per step)
    long long* d_tmodel = nullptr; float* d_coef = nullptr; int S = 0;
    const float* train_target = nullptr;
};

__global__ void k_set_int(int* p, int v) {
    pdl_entry(); *p = v; }

extern "C" {

const char* ddpm_last_error(void) { return last_error().c_str(); }

int ddpm_runtime_check(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(-1, "no CUDA device / driver"); }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp pr;
    DDPM_CUDA_OK(cudaGetDeviceProperties(&pr, dev));
    if (pr.major != 10) return fail(-1, "device %d is sm_%d%d, this library is sm_100a only", dev, pr.major, pr.minor);
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

int ddpm_conv_halo_run(const ddpm_halo_desc* d, void* stream) {
    HaloLaunch g;
    int rc = build_halo(*d, g);
    if (rc) return rc;
    return launch_halo(g, static_cast<cudaStream_t>(stream));
}

int ddpm_attn_fused_run(const void* qkv, void* out, void* probs, int NB, int T, int C, void* stream) {
    AttnLaunch g;
    int rc = build_attn(qkv, out, NB, T, C, g, probs);
    if (rc) return rc;
    return launch_attn(g, static_cast<cudaStream_t>(stream));
}

int ddpm_attn_fused_bwd_run(const void* qkv, const void* d_out, const void* probs, void* d_scores, void* d_qkv, int NB, int T, int C, void* stream) {
    AttnLaunch g;
    int rc = build_attn_bwd(qkv, d_out, probs, d_scores, d_qkv, NB, T, C, g);
    if (rc) return rc;
    return launch_attn_bwd(g, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------ UNet engine
int ddpm_unet_create(const ddpm_unet_cfg* cfg, ddpm_unet** out) {
    if (!cfg || !out) return fail(-30, "null argument");
    if (cfg->levels < 1 || cfg->levels > 8) return fail(-30, "levels must be 1..8");
    if (cfg->hid_channels <= 0 || cfg->hid_channels % 32) return fail(-30, "hid_channels must be a positive multiple of 32");
    if (cfg->num_res_blocks < 1) return fail(-30, "num_res_blocks must be >= 1");
    ddpm_unet* h = new (std::nothrow) ddpm_unet();
    if (!h) return fail(-30, "out of host memory");
    h->e.cfg = *cfg;
    if (h->e.cfg.temb_dim <= 0) h->e.cfg.temb_dim = 4 * cfg->hid_channels;
    h->e.register_params();
    *out = h;
    return 0;
}
void ddpm_unet_destroy(ddpm_unet* h) {
    if (!h) return;
    if (h->d_tmodel) cudaFree(h->d_tmodel);
    if (h->d_coef) cudaFree(h->d_coef);
    if (h->e.side_stream) { cudaStreamSynchronize(h->e.side_stream);   // a queued weight re-pack still writes into the caller's workspace
        cudaStreamDestroy(h->e.side_stream); cudaEventDestroy(h->e.ev_fork); cudaEventDestroy(h->e.ev_join); }
    for (auto& c : h->e.chunks) if (c.ev) cudaEventDestroy(c.ev);
    if (h->e.ev_pack_all) { cudaEventDestroy(h->e.ev_pack_fork); cudaEventDestroy(h->e.ev_pack_fc); cudaEventDestroy(h->e.ev_pack_all); }
    delete h;
}
int ddpm_unet_num_params(const ddpm_unet* h) { return (int)h->e.params.size(); }
int ddpm_unet_param_info(const ddpm_unet* h, int i, const char** name, int* ndim, int dims[4], long long* offset) {
    if (i < 0 || i >= (int)h->e.params.size()) return fail(-30, "param index out of range");
    const ParamInfo& p = h->e.params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = p.nd;
    if (dims) for (int k = 0; k < 4; ++k) dims[k] = p.dims[k];
    if (offset) *offset = p.off;
    return 0;
}
long long ddpm_unet_flat_elems(const ddpm_unet* h) { return h->e.flat_elems; }

long long ddpm_unet_workspace_bytes(ddpm_unet* h, int B, int H, int W, int training) {
    UnetEngine& e = h->e;
    e.P = nullptr; e.G = nullptr; e.ws = nullptr; e.planned = false;
    e.zf_cursor = e.zb_cursor = 0;
    const int rc = e.plan(B, H, W, training != 0, true);
    if (rc) return rc;
    return (long long)(e.cursor + e.zf_cursor + e.zb_cursor + 8192);
}
int ddpm_unet_plan(ddpm_unet* h, int B, int H, int W, int training, float* params_flat, float* grads_flat, void* workspace, long long workspace_bytes) {
    UnetEngine& e = h->e;
    if (!params_flat || !workspace) return fail(-30, "params / workspace pointer is null");
    if (training && !grads_flat) return fail(-30, "training plan needs a gradient buffer");
    int rc = ddpm_runtime_check(); if (rc) return rc;
    e.planned = false;
    e.P = params_flat; e.G = grads_flat; e.ws = static_cast<uint8_t*>(workspace); e.ws_bytes = (size_t)workspace_bytes;
    e.zf_cursor = e.zb_cursor = 0;
    if ((rc = e.plan(B, H, W, training != 0, true))) return rc;
    if ((long long)(e.cursor + e.zf_cursor + e.zb_cursor + 8192) > workspace_bytes)
        return fail(-32, "workspace too small: need %lld bytes, got %lld", (long long)(e.cursor + e.zf_cursor + e.zb_cursor + 8192), workspace_bytes);
    if ((rc = e.plan(B, H, W, training != 0, false))) return rc;
    if ((long long)e.cursor > workspace_bytes) return fail(-32, "workspace overflow (internal)");
    return e.build();
}
#define NEED_PLAN(h) do { if (!(h) || !(h)->e.planned) return fail(-33, "ddpm_unet_plan has not been called"); } while (0)

int ddpm_unet_repack(ddpm_unet* h, void* stream) { NEED_PLAN(h); return h->e.repack(static_cast<cudaStream_t>(stream)); }

int ddpm_unet_forward(ddpm_unet* h, const float* x, const int64_t* t, float* eps, uint64_t dropout_seed, void* stream) {
    NEED_PLAN(h);
    UnetEngine& e = h->e;
    e.x_in = x; e.t_in = reinterpret_cast<const long long*>(t); e.eps_dst = eps; e.drop_seed = dropout_seed;
    return e.run_list(e.fwd_ops, static_cast<cudaStream_t>(stream));
}
int ddpm_unet_backward(ddpm_unet* h, const float* d_eps, void* stream) {
    NEED_PLAN(h);
    UnetEngine& e = h->e;
    if (!e.train) return fail(-33, "plan was built without training");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    e.deps_src = d_eps;
    return e.run_list(e.bwd_ops, st);
}
int ddpm_train_forward(ddpm_unet* h, const float* x0, const int64_t* t, const float* noise, const float* tab_a, const float* tab_s,
                       float* losses, uint64_t dropout_seed, void* stream) {
    NEED_PLAN(h);
    UnetEngine& e = h->e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int per_img = e.cfg.in_channels * e.H * e.W;
    const long long total = (long long)e.B * per_img;
    float* xt = e.at<float>(e.xt_off); float* eps = e.at<float>(e.eps_off);
    (void)per_img; (void)total;
    // q_sample (diffusion.py:92-97) is in_conv's load prologue: x_t is formed on the fly (and stored once, for the backward)
    e.qs_pro = QsamplePro{noise, reinterpret_cast<const long long*>(t), tab_a, tab_s, xt};
    e.x_in = x0; e.t_in = reinterpret_cast<const long long*>(t); e.eps_dst = eps; e.drop_seed = dropout_seed;
    const int rc = e.run_list(e.fwd_ops, st);
    e.qs_pro = QsamplePro{};
    e.x_in = xt;                                       // what the backward pass reads as the network input
    if (rc) return rc;
    h->train_target = noise;
    launch_k(k_mse, e.B, 256, 0, st, eps, noise, losses, e.cfg.out_channels * e.H * e.W);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
int ddpm_train_backward(ddpm_unet* h, const float* gscale, void* stream) {
    NEED_PLAN(h);
    UnetEngine& e = h->e;
    if (!e.train || !h->train_target) return fail(-33, "ddpm_train_forward must precede ddpm_train_backward on a training plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int per_img = e.cfg.out_channels * e.H * e.W;
    const long long total = (long long)e.B * per_img;
    launch_k(k_mse_grad, grid_for(total), 256, 0, st, e.at<float>(e.eps_off), h->train_target, gscale, e.at<float>(e.deps_off), per_img, total);
    DDPM_CUDA_OK(cudaGetLastError());
    e.deps_src = e.at<float>(e.deps_off);
    return e.run_list(e.bwd_ops, st);
}

int ddpm_sampler_setup(ddpm_unet* h, int S, const int64_t* t_model_host, const float* coef_host) {
    if (!h || S <= 0) return fail(-30, "bad sampler setup");
    if (h->S != S) {           // same length: keep the allocations (a captured sampler graph holds their addresses)
        if (h->d_tmodel) { cudaFree(h->d_tmodel); h->d_tmodel = nullptr; }
        if (h->d_coef) { cudaFree(h->d_coef); h->d_coef = nullptr; }
        DDPM_CUDA_OK(cudaMalloc(&h->d_tmodel, (size_t)S * 8));
        DDPM_CUDA_OK(cudaMalloc(&h->d_coef, (size_t)S * 6 * 4));
    }
    DDPM_CUDA_OK(cudaMemcpy(h->d_tmodel, t_model_host, (size_t)S * 8, cudaMemcpyHostToDevice));
    DDPM_CUDA_OK(cudaMemcpy(h->d_coef, coef_host, (size_t)S * 6 * 4, cudaMemcpyHostToDevice));
    h->S = S;
    return 0;
}
int ddpm_sampler_reset(ddpm_unet* h, int first_step, void* stream) {
    NEED_PLAN(h);
    if (first_step < 0 || first_step >= h->S) return fail(-30, "first_step out of range");
    launch_k(k_set_int, 1, 1, 0, static_cast<cudaStream_t>(stream), h->e.at<int>(h->e.counter_off), first_step);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
int ddpm_sampler_step(ddpm_unet* h, float* x, const float* z, uint64_t seed, void* stream) {
    return ddpm_sampler_step_pred(h, x, z, seed, nullptr, stream);
}
int ddpm_sampler_step_pred(ddpm_unet* h, float* x, const float* z, uint64_t seed, float* pred_x0, void* stream) {
    NEED_PLAN(h);
    if (!h->d_coef) return fail(-33, "ddpm_sampler_setup has not been called");
    UnetEngine& e = h->e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long* tbuf = e.at<long long>(e.tbuf_off); float* cc = e.at<float>(e.coefcur_off); float* eps = e.at<float>(e.eps_off);
    launch_k(k_sampler_prep, 1, 256, 0, st, e.at<int>(e.counter_off), h->d_tmodel, h->d_coef, tbuf, cc, e.B);
    DDPM_CUDA_OK(cudaGetLastError());
    e.x_in = x; e.t_in = tbuf; e.eps_dst = eps; e.drop_seed = 0;
    static const bool no_uni = getenv("DDPM_NO_UNIFORM_T") != nullptr;
    e.uniform_t = !no_uni;                                // the whole batch is at the same timestep (diffusion.py:166)
    // the alpha/beta update of diffusion.py:107-158 rides in the epilogue of the final conv (k_out_gather): eps never reaches
    // memory and the step has no tail launch.  (Plans whose out_conv is not on the tensor-core path keep the separate tail.)
    const bool fused = e.gather_fused_tail;
    if (fused) { e.ps_epi.x = x; e.ps_epi.z = z; e.ps_epi.coef = cc; e.ps_epi.seed = (unsigned long long)seed; e.ps_epi.pred = pred_x0; }
    const int rc = e.run_list(e.fwd_ops, st);
    e.uniform_t = false;
    e.ps_epi = PsampleEpi{nullptr, nullptr, nullptr, 0, nullptr};
    if (rc) return rc;
    if (!fused) {
        const long long total = (long long)e.B * e.cfg.out_channels * e.H * e.W;
        launch_k(k_psample_tail, grid_for(total), 256, 0, st, eps, x, z, cc, (unsigned long long)seed, total, pred_x0);
        DDPM_CUDA_OK(cudaGetLastError());
    }
    return 0;
}
int ddpm_opt_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow, long long n,
                  const ddpm_opt_cfg* c, void* state, float* norm_out, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !c || !state || !norm_out) return fail(-40, "ddpm_opt_step: null buffer");
    if (n <= 0 || (n & 3)) return fail(-40, "ddpm_opt_step: n must be a positive multiple of 4");
    if (c->step < 1) return fail(-40, "ddpm_opt_step: step is 1-based");
    if (c->ema_decay >= 0 && !ema_shadow) return fail(-40, "ddpm_opt_step: EMA enabled but ema_shadow is NULL");
    for (const void* q : {(const void*)params, (const void*)grads, (const void*)exp_avg, (const void*)exp_avg_sq, (const void*)ema_shadow})
        if (reinterpret_cast<uintptr_t>(q) & 15) return fail(-40, "ddpm_opt_step: buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OptScalars o;
    const double bc1 = 1.0 - pow(c->beta1, (double)c->step), bc2 = 1.0 - pow(c->beta2, (double)c->step);
    o.one_minus_b1 = (float)(1.0 - c->beta1); o.b2 = (float)c->beta2; o.one_minus_b2 = (float)(1.0 - c->beta2);
    o.step_size = (float)(c->lr / bc1); o.bc2_sqrt = (float)sqrt(bc2); o.eps = (float)c->eps;
    o.max_norm = (float)c->max_grad_norm;
    if (c->ema_decay >= 0) {
        const double nu = (double)c->ema_num_updates;
        double d = (1.0 + nu) / (10.0 + nu); if (d > c->ema_decay) d = c->ema_decay;
        o.ema_w = (float)(1.0 - d);
    } else o.ema_w = -1.f;
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    const long long n4 = n >> 2;
    long long want = (n4 + 255) / 256; const int cap = num_sms * 8; const int grid = (int)(want < cap ? want : cap);
    double* acc = reinterpret_cast<double*>(state); unsigned int* ticket = reinterpret_cast<unsigned int*>(acc + 1);
    launch_k(k_grad_sumsq, grid, 256, 0, st, reinterpret_cast<const float4*>(grads), n4, acc, ticket, norm_out, o.max_norm);
    DDPM_CUDA_OK(cudaGetLastError());
    launch_k(k_adam_ema, grid, 256, 0, st, reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(exp_avg),
             reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<float4*>(ema_shadow), n4, o, (const float*)norm_out);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
int ddpm_to_uint8_nhwc(const float* x, uint8_t* out, int B, int C, int H, int W, void* stream) {
    if (!x || !out) return fail(-41, "ddpm_to_uint8_nhwc: null buffer");
    if (B <= 0 || H <= 0 || W <= 0 || C < 1 || C > 4) return fail(-41, "ddpm_to_uint8_nhwc: need B,H,W > 0 and 1 <= C <= 4");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long npix = (long long)B * H * W;
    const int grid = (int)((npix + 255) / 256 < 148 * 16 ? (npix + 255) / 256 : 148 * 16);
    switch (C) {
        case 1: launch_k(k_to_uint8_nhwc<1>, grid, 256, 0, st, x, out, npix, H * W); break;
        case 2: launch_k(k_to_uint8_nhwc<2>, grid, 256, 0, st, x, out, npix, H * W); break;
        case 3: launch_k(k_to_uint8_nhwc<3>, grid, 256, 0, st, x, out, npix, H * W); break;
        default: launch_k(k_to_uint8_nhwc<4>, grid, 256, 0, st, x, out, npix, H * W); break;
    }
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
int ddpm_unet_grad_chunks(const ddpm_unet* h, int max_n, long long* lo, long long* hi) {
    if (!h || !h->e.planned || !h->e.train) return fail(-33, "ddpm_unet_grad_chunks: no training plan");
    const int n = (int)h->e.chunks.size();
    for (int i = 0; i < n && i < max_n; ++i) { if (lo) lo[i] = h->e.chunks[i].lo; if (hi) hi[i] = h->e.chunks[i].hi; }
    return n;
}
int ddpm_unet_wait_grad_chunk(ddpm_unet* h, int i, void* stream) {
    NEED_PLAN(h);
    if (i < 0 || i >= (int)h->e.chunks.size() || !h->e.chunks[i].ev) return fail(-30, "gradient chunk index out of range");
    DDPM_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->e.chunks[i].ev, 0));
    return 0;
}
int ddpm_unet_plan_stats(const ddpm_unet* h, int* n_fwd, int* n_bwd, int* n_tc, int* n_gen, double* ff, double* bf) {
    if (n_fwd) *n_fwd = (int)h->e.fwd_ops.size();
    if (n_bwd) *n_bwd = (int)h->e.bwd_ops.size();
    if (n_tc) *n_tc = h->e.n_tc_gemms;
    if (n_gen) *n_gen = h->e.n_generic;
    if (ff) *ff = h->e.fwd_flops;
    if (bf) *bf = h->e.bwd_flops;
    return 0;
}
int ddpm_unet_launches_per_forward(const ddpm_unet* h) { return UnetEngine::count_launches(h->e.fwd_ops); }
int ddpm_unet_launch_counts(const ddpm_unet* h, int* fwd, int* bwd, int* pack) {
    if (fwd) *fwd = UnetEngine::count_launches(h->e.fwd_ops);
    if (bwd) *bwd = UnetEngine::count_launches(h->e.bwd_ops);
    if (pack) *pack = UnetEngine::count_launches(h->e.pack_ops);
    return 0;
}

}  // extern "C"
