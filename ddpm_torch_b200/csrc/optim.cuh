// Fused optimizer step over the flat fp32 buffers: global-norm gradient clipping
// (utils/train.py:159, torch.nn.utils.clip_grad_norm_), Adam (train.py:128, torch.optim.Adam defaults: eps 1e-8, no weight
// decay, no amsgrad) and the EMA shadow update with its warm-up (utils/train.py:300-305) in TWO launches:
//   k_grad_sumsq   : sum of squares of the flat gradient (fp64 accumulate); the last-arriving block turns it into
//                    {total_norm, clip_coef = min(1, max_norm / (total_norm + 1e-6))} on the device - no host sync
//   k_adam_ema     : g' = clip_coef * g ; m += (1-b1)(g'-m) ; v = b2 v + (1-b2) g'^2 ;
//                    p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps) ; shadow += (1-decay)(p - shadow)
// HBM roofline: 20 B read + 16 B written per parameter (16 + 12 without EMA) + 4 B for the norm pass.
#pragma once
#include "ptx.cuh"

namespace ddpm {

struct OptScalars {          // host-computed in double, as torch does, then rounded once to fp32
    float one_minus_b1, b2, one_minus_b2, step_size /* lr / (1 - b1^t) */, bc2_sqrt /* sqrt(1 - b2^t) */, eps;
    float max_norm;          // <= 0: no clipping
    float ema_w;             // 1 - decay_t ; < 0: no EMA
};

__global__ void __launch_bounds__(256) k_grad_sumsq(const float4* __restrict__ g, long long n4, double* __restrict__ acc /*[1], zeroed by the finishing block*/,
                                                    unsigned int* __restrict__ ticket, float* __restrict__ scal /*[2]: total_norm, clip_coef*/, float max_norm) {
    pdl_entry();
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g + i);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    double d = (double)s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(acc, t);
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {      // last block: finish and re-arm for the next step
            __threadfence();
            const double tot = *reinterpret_cast<volatile double*>(acc);
            const float norm = (float)sqrt(tot);
            scal[0] = norm;
            float c = 1.f;
            if (max_norm > 0.f) { c = max_norm / (norm + 1e-6f); if (c > 1.f) c = 1.f; }
            scal[1] = c;
            *acc = 0.0; *ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(256) k_adam_ema(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                                  float4* __restrict__ shadow, long long n4, OptScalars o, const float* __restrict__ scal) {
    pdl_entry();
    const float clip = (o.max_norm > 0.f) ? scal[1] : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g4 = __ldg(g + i); float4 p4 = p[i], m4 = m[i], v4 = v[i];
        float gg[4] = {g4.x, g4.y, g4.z, g4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gc = gg[e] * clip;
            mm[e] = fmaf(o.one_minus_b1, gc - mm[e], mm[e]);                       // exp_avg.lerp_(grad, 1 - beta1)
            vv[e] = fmaf(o.one_minus_b2 * gc, gc, o.b2 * vv[e]);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = __fdiv_rn(sqrtf(vv[e]), o.bc2_sqrt) + o.eps;          // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
            pp[e] = fmaf(-o.step_size, __fdiv_rn(mm[e], denom), pp[e]);              // param.addcdiv_(exp_avg, denom, value=-step_size)
        }
        p[i] = make_float4(pp[0], pp[1], pp[2], pp[3]); m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]); v[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (o.ema_w >= 0.f) {
            float4 s4 = shadow[i];
            s4.x = fmaf(o.ema_w, pp[0] - s4.x, s4.x); s4.y = fmaf(o.ema_w, pp[1] - s4.y, s4.y);
            s4.z = fmaf(o.ema_w, pp[2] - s4.z, s4.z); s4.w = fmaf(o.ema_w, pp[3] - s4.w, s4.w);
            shadow[i] = s4;
        }
    }
}

}  // namespace ddpm
