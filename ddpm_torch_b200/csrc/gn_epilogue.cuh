// GroupNorm work fused into the tcgen05 conv / GEMM epilogues (unet.py:18-20, 83-89).
//
// An epilogue lane owns 32 consecutive output channels of ONE pixel (TMEM lane = accumulator row).  Two fusions:
//
//  forward  (GnEpi::qstats)  the producing conv adds, per (image, 4-channel quad), {sum x, sum x^2} of its OUTPUT to
//           qstats[NB][C/4][2] (fp64).  Quads are the common granularity of every consumer's group structure: GroupNorm(32)
//           over C or over a concat C0+C1 has groups of C/32 channels, a multiple of 4 for C % 128 == 0, and concat boundaries
//           are multiples of 4 too - so a consumer folds whole quads of one or two producers into its groups.  This removes
//           the statistics pass over the tensor (k_gn_stats: 2 B/element of HBM reads per norm).
//
//  backward (GnEpi::K)       the data-gradient conv's accumulator is dy, the gradient at a = mask*silu(gn(x)).  The epilogue
//           loads x, rebuilds n = sc*x + sh, turns dy into dn = dy*keep*silu'(n), stores dn instead of dy and adds the
//           group-level terms of the GroupNorm backward, at quad granularity, to gs[NB][C/4][2] (fp32):
//           {sum gamma*dn, sum gamma*dn*xh}.  This removes the reduce pass (k_gn_bwd_reduce: 6 B/element); the apply pass
//           folds the quads into groups in its prologue and accumulates dgamma / dbeta on the fly.
//
// Both reductions run over the 32 lanes (pixels) of the warp with a halving butterfly: at every stage a lane hands half of
// its partial sums to its partner and keeps the other half, so 16 values cost 16 shuffles in total instead of 80.
#pragma once
#include "../../include/ddpm_b200.h"
#include "ptx.cuh"

namespace ddpm {

struct GnEpi {
    double* qstats;                       // forward: [NB][C/4][2] fp64, zeroed per pass; null = off
    const __nv_bfloat16* x0; const __nv_bfloat16* x1; int C0, C1;   // backward: GroupNorm input (concat of up to two NHWC tensors)
    const float* K;                       // backward: [NB][4][C] {sc, sh, rstd, mean*rstd}; null = off
    const float* gamma; const float* beta;   // backward: the norm's affine parameters [C]
    float* gs;                            // backward: [NB][C/4][2] {sum gamma*dn, sum gamma*dn*xh} per quad, zeroed per pass
    const unsigned char* mask;            // backward: dropout keep bits [pixel][C/8] or null
    float keep_scale; int silu;
};

inline GnEpi gn_epi_from_abi(const ddpm_gn_epi& a) {
    GnEpi g;
    g.qstats = a.qstats; g.x0 = reinterpret_cast<const __nv_bfloat16*>(a.gnb_x0); g.x1 = reinterpret_cast<const __nv_bfloat16*>(a.gnb_x1);
    g.C0 = a.gnb_C0; g.C1 = a.gnb_x1 ? a.gnb_C1 : 0; g.K = a.gnb_K; g.gamma = a.gnb_gamma; g.beta = a.gnb_beta; g.gs = a.gnb_gs; g.mask = a.gnb_mask;
    g.keep_scale = a.gnb_keep_scale > 0.f ? a.gnb_keep_scale : 1.f; g.silu = a.gnb_silu;
    return g;
}

// halving butterfly over the warp: V values per lane in, V/32 out (V = 16: one value on even lanes after the last pair-sum)
template <int HALF>
__device__ __forceinline__ void bfly_stage(float* v, int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float send = up ? v[i] : v[i + HALF];
        const float keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
}

// forward: quad statistics of this lane's 32 channels; dst = &qstats[(image * C/4 + first_channel/4) * 2].
// Must be called by all 32 lanes (invalid rows contribute zeros).
__device__ __forceinline__ void epi_quad_stats(const float (&f)[32], bool valid, double* dst, int lane) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = f[4 * j], b = f[4 * j + 1], c = f[4 * j + 2], d = f[4 * j + 3];
        v[2 * j] = valid ? (a + b) + (c + d) : 0.f;
        v[2 * j + 1] = valid ? fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d))) : 0.f;
    }
    bfly_stage<8>(v, lane, 16);
    bfly_stage<4>(v, lane, 8);
    bfly_stage<2>(v, lane, 4);
    bfly_stage<1>(v, lane, 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    // lane L now holds value index (L >> 1) & 15 = quad*2 + {0: sum, 1: sum of squares}, summed over the warp's 32 pixels
    if (!(lane & 1)) atomicAdd(dst + ((lane >> 1) & 15), (double)v[0]);
}

__device__ __forceinline__ float epi_sigmoid(float y) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
    return fmaf(0.5f, t, 0.5f);
}

// backward: f = dy (32 channels [col, col+32) of one pixel) -> dn in place, plus the GROUP-level terms of the GroupNorm
// backward at quad granularity:  A_q = sum_{c in quad} gamma_c * dn ,  B_q = sum_{c in quad} gamma_c * dn * xh = dn * (y - beta_c)
// (y = sc*x + sh = gamma*xh + beta), reduced over the warp's 32 pixels and added to gs[(image * C/4 + col/4) * 2 ...].
// The per-CHANNEL sums (dgamma, dbeta) are left to the apply pass, whose threads own a channel octet over many pixels and
// accumulate them in registers - a 16-value butterfly here instead of two 32-value ones.
//   rs      : the 32 x values of this lane, packed bf16 pairs (prefetched by the caller)
//   keep    : dropout keep bits of the 32 channels (HAS_MASK)
//   vectors : per-channel {sc, sh, gamma, beta} of the 32 channels - SMEM: shared-memory byte address of sc[32], the other three
//             `vstride` floats apart (broadcast reads); else four global pointers (L1)
// Must be called by all 32 lanes; `valid` = false rows contribute zeros.
template <bool SMEM, bool HAS_MASK>
__device__ __forceinline__ void epi_gn_bwd(float (&f)[32], bool valid, const uint32_t (&rs)[16], uint32_t keep, float keep_scale, int silu,
                                           uint32_t vec_smem, int vstride, const float* sc_g, const float* sh_g, const float* gm_g, const float* be_g,
                                           float* gs_dst, int lane) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {          // one quad (4 channels) per step
        float4 a, b, gm, be;
        if (SMEM) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(vec_smem + 16u * j));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(vec_smem + 4u * vstride + 16u * j));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(gm.x), "=f"(gm.y), "=f"(gm.z), "=f"(gm.w) : "r"(vec_smem + 8u * vstride + 16u * j));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(be.x), "=f"(be.y), "=f"(be.z), "=f"(be.w) : "r"(vec_smem + 12u * vstride + 16u * j));
        } else {
            a = __ldg(reinterpret_cast<const float4*>(sc_g) + j); b = __ldg(reinterpret_cast<const float4*>(sh_g) + j);
            gm = __ldg(reinterpret_cast<const float4*>(gm_g) + j); be = __ldg(reinterpret_cast<const float4*>(be_g) + j);
        }
        const float scv[4] = {a.x, a.y, a.z, a.w}, shv[4] = {b.x, b.y, b.z, b.w}, gv[4] = {gm.x, gm.y, gm.z, gm.w}, bv[4] = {be.x, be.y, be.z, be.w};
        const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rs[2 * j]));
        const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rs[2 * j + 1]));
        const float xv[4] = {x01.x, x01.y, x23.x, x23.y};
        float A = 0.f, Bq = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 4 * j + i;
            float dn = f[e] * keep_scale;
            if (HAS_MASK) dn = ((keep >> e) & 1u) ? dn : 0.f;
            const float y = fmaf(xv[i], scv[i], shv[i]);
            if (silu) {
                const float sg = epi_sigmoid(y);
                dn *= sg * fmaf(y, 1.f - sg, 1.f);
            }
            f[e] = dn;
            A = fmaf(gv[i], dn, A);
            Bq = fmaf(dn, y - bv[i], Bq);
        }
        v[2 * j] = valid ? A : 0.f; v[2 * j + 1] = valid ? Bq : 0.f;
    }
    bfly_stage<8>(v, lane, 16);
    bfly_stage<4>(v, lane, 8);
    bfly_stage<2>(v, lane, 4);
    bfly_stage<1>(v, lane, 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    if (!(lane & 1)) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gs_dst + ((lane >> 1) & 15)), "f"(v[0]) : "memory");
}

}  // namespace ddpm
