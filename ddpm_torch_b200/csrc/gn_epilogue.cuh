// GroupNorm statistics fused into the tcgen05 conv / GEMM epilogues (unet.py:18-20, 83-89).
//
// An epilogue lane owns 32 consecutive output channels of ONE pixel (TMEM lane = accumulator row).  The producing conv adds,
// per (image, 4-channel quad), {sum x, sum x^2} of its OUTPUT to qstats[NB][C/4][2] (fp64).  Quads are the common granularity
// of every consumer's group structure: GroupNorm(32) over C or over a concat C0+C1 has groups of C/32 channels, a multiple of
// 4 for C % 128 == 0, and concat boundaries are multiples of 4 too - so a consumer folds whole quads of one or two producers
// into its groups.  This removes the statistics pass over the tensor (k_gn_stats: 2 B/element of HBM reads per norm) at a cost
// of +2..4 % on the conv (profiles/r02_halo_epilogue_variants_v2.txt).
//
// The reduction over the 32 lanes (pixels) of the warp is a halving butterfly: at every stage a lane hands half of its partial
// sums to its partner and keeps the other half, so 16 values cost 16 shuffles in total instead of 80.
//
// (The matching BACKWARD fusion - dy -> dn and the group sums in the data-gradient conv's epilogue - was built, unit-tested and
// measured in round 2 and removed: with 8 epilogue warps per SM the ~1000-instruction chunk body is latency-bound and the conv
// slows down by more than the reduce pass it replaces; profiles/r02_gn_backward_epilogue_experiment.txt.)
#pragma once
#include "../../include/ddpm_b200.h"
#include "ptx.cuh"

namespace ddpm {

struct GnEpi {
    double* qstats;                       // [NB][C/4][2] fp64, zeroed per pass; null = off
};

inline GnEpi gn_epi_from_abi(const ddpm_gn_epi& a) { GnEpi g; g.qstats = a.qstats; return g; }

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

}  // namespace ddpm
