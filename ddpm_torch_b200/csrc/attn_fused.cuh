// Fused self-attention core of AttentionBlock (unet.py:43-51) for the 16x16 level:  O = softmax(Q K^T / sqrt(C)) V
// with T = 256 tokens and C = 256 channels (single head), ONE launch instead of {Q.K^T GEMM -> fp32 S in HBM -> softmax kernel ->
// bf16 P in HBM -> P.V GEMM}.  S and O live in TMEM, P in shared memory; nothing but Q, K, V in and O out touches HBM.
//
// One CTA per (image, 128-query half); 6 warps:
//   warp 0     TMA producer: Q [128 x 64ch] + K [256 x 64ch] per channel chunk (4 chunks), then V [64 keys x 256 ch] per key chunk (4)
//   warp 1     MMA issuer:   S[128 x 256] = Q.K^T into TMEM columns 0..255 (M=128, N=256, K-major both);
//                            O[128 x 256] = P.V   into TMEM columns 256..511 (A = P from shared memory K-major, B = V MN-major)
//   warps 2-5  softmax + epilogue, one query row per thread: row max over the 256 scores (pass 1), p = exp2((s - max) * log2e/sqrt(C))
//              as bf16 into the swizzled P tile + row sum (pass 2), then O * (1 / sum) -> bf16 -> global.
// The P tile uses exactly the layout TMA would produce for a K-major SWIZZLE_128B operand (rows of 128 B = 64 keys, 16-byte chunk j
// of row r stored at chunk j ^ (r & 7), 16 KB per 64-key slab), so the UMMA descriptors are the ones of the GEMM engine.
// Training plans additionally get the normalised P written to HBM (p.pm) for the backward pass, from the same shared-memory tile.
#pragma once
#include "gemm_host.cuh"

namespace ddpm {

struct AttnParams {
    int NB; __nv_bfloat16* out; float scale_log2e;
    __nv_bfloat16* pm;             // forward: optional [NB][T][T] normalised probabilities out (training).  backward: the saved P (in)
    __nv_bfloat16* ds;             // backward: [NB][T][T] dS out (the dK / dV GEMMs read it)
    int out_ld;                    // row stride of `out` in elements (forward C; backward 3C: dQ lands in the q third of dqkv)
    float scale;                   // backward: 1/sqrt(C)
};

constexpr int ATTN_T = 256, ATTN_D = 256;
constexpr int ATTN_STAGES = 3;
constexpr int ATTN_STAGE_BYTES = 16384 + 32768;           // A (Q chunk) + B (K chunk / V chunk)
constexpr int ATTN_P_BYTES = 128 * ATTN_T * 2;            // 64 KB
constexpr int ATTN_SMEM = ATTN_STAGES * ATTN_STAGE_BYTES + ATTN_P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int ATTN_THREADS = 192;

// BWD = false: the forward described above.
// BWD = true : the query-side half of the backward pass with the same pipeline, operands swapped:
//                dP[128 x 256] = dO . V^T   (A = dO chunks through tmQ, B = V rows through tmK at channel offset 2C)
//                dS = P o (dP - rowsum(P o dP)) / sqrt(C)     (softmax warps; P read from HBM, dS -> shared tile + HBM)
//                dQ[128 x 256] = dS . K     (A = dS tile, B = K rows MN-major through tmV at channel offset C)
//              replacing three launches (dP GEMM -> fp32 in HBM, softmax backward, dQ GEMM); dK = dS^T Q and dV = P^T dO stay
//              GEMM launches that read dS / P from HBM.
template <bool BWD>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smP = smem + ATTN_STAGES * ATTN_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smP + ATTN_P_BYTES);
    uint64_t* empty_bar = full_bar + ATTN_STAGES;
    uint64_t* s_full = empty_bar + ATTN_STAGES;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_full = p_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x >> 1, half = blockIdx.x & 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        for (int s = 0; s < ATTN_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stg = 0;
            bool ok = true;
            for (int kc = 0; kc < ATTN_D / 64 && ok; ++kc, ++stg) {      // Q and K channel chunks
                const int st = stg % ATTN_STAGES; const uint32_t ph = (stg / ATTN_STAGES) & 1;
                if (!mbar_wait(&empty_bar[st], ph ^ 1, 11)) { ok = false; break; }
                uint8_t* sb = smem + st * ATTN_STAGE_BYTES;
                mbar_expect_tx(&full_bar[st], 16384 + 32768);
                tma_load_4d(sb, &tmQ, &full_bar[st], kc * 64, half * 128, 0, img);
                tma_load_4d(sb + 16384, &tmK, &full_bar[st], (BWD ? 2 * ATTN_D : ATTN_D) + kc * 64, 0, 0, img);
            }
            for (int j = 0; j < ATTN_T / 64 && ok; ++j, ++stg) {         // V key chunks (prefetched while the softmax runs)
                const int st = stg % ATTN_STAGES; const uint32_t ph = (stg / ATTN_STAGES) & 1;
                if (!mbar_wait(&empty_bar[st], ph ^ 1, 12)) { ok = false; break; }
                uint8_t* sb = smem + st * ATTN_STAGE_BYTES + 16384;
                mbar_expect_tx(&full_bar[st], 32768);
#pragma unroll
                for (int b = 0; b < ATTN_D / 64; ++b)
                    tma_load_4d(sb + b * 8192, &tmV, &full_bar[st], (BWD ? ATTN_D : 2 * ATTN_D) + b * 64, j * 64, 0, img);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_qk = umma_idesc(128, 256, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc(128, 256, 0, 1);
            int stg = 0;
            bool ok = true;
            for (int kc = 0; kc < ATTN_D / 64 && ok; ++kc, ++stg) {
                const int st = stg % ATTN_STAGES; const uint32_t ph = (stg / ATTN_STAGES) & 1;
                if (!mbar_wait(&full_bar[st], ph, 13)) { ok = false; break; }
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + st * ATTN_STAGE_BYTES), b_addr = a_addr + 16384;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc_qk, (kc | k) != 0);
                umma_commit(&empty_bar[st]);
            }
            if (ok) umma_commit(s_full);
            if (ok && !mbar_wait(p_ready, 0, 14)) ok = false;             // P tile written and fenced by the softmax warps
            tc_fence_after();
            for (int j = 0; j < ATTN_T / 64 && ok; ++j, ++stg) {
                const int st = stg % ATTN_STAGES; const uint32_t ph = (stg / ATTN_STAGES) & 1;
                if (!mbar_wait(&full_bar[st], ph, 15)) { ok = false; break; }
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smP + j * 16384), b_addr = smem_u32(smem + st * ATTN_STAGE_BYTES + 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + 256, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 2048, 8192, 1024), idesc_pv, (j | k) != 0);
                umma_commit(&empty_bar[st]);
            }
            if (ok) umma_commit(o_full);
        }
    } else {
        // ======================= softmax + epilogue: one query row per thread =======================
        const int q = warp & 3;                       // TMEM lane quarter of this warp (warps 2,3,4,5 -> 2,3,0,1)
        const int r = q * 32 + lane;                  // row of the tile
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        if (BWD) {
          if (mbar_wait(s_full, 0, 16)) {
            tc_fence_after();
            const __nv_bfloat16* prow = p.pm + ((long long)img * ATTN_T + half * 128 + r) * ATTN_T;
            __nv_bfloat16* dsrow = p.ds + ((long long)img * ATTN_T + half * 128 + r) * ATTN_T;
            float delta = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t v[32], pr[16];
                tmem_ld32(t_row + (uint32_t)(ch * 32), v);
                ld_global_nc_256(prow + ch * 32, pr); ld_global_nc_256(prow + ch * 32 + 16, pr + 8);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 pp = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pr[e]));
                    delta = fmaf(pp.x, __uint_as_float(v[2 * e]), fmaf(pp.y, __uint_as_float(v[2 * e + 1]), delta));
                }
            }
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t v[32], pr[16], pk[16];
                tmem_ld32(t_row + (uint32_t)(ch * 32), v);
                ld_global_nc_256(prow + ch * 32, pr); ld_global_nc_256(prow + ch * 32 + 16, pr + 8);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 pp = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pr[e]));
                    pk[e] = pack_bf16x2(pp.x * (__uint_as_float(v[2 * e]) - delta) * p.scale, pp.y * (__uint_as_float(v[2 * e + 1]) - delta) * p.scale);
                }
                uint8_t* rowp = smP + (ch >> 1) * 16384 + r * 128;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(rowp + ((((ch & 1) * 4 + i) ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                st_global_256(dsrow + ch * 32, pk); st_global_256(dsrow + ch * 32 + 16, pk + 8);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            if (mbar_wait(o_full, 0, 17)) {
                tc_fence_after();
                __nv_bfloat16* orow = p.out + ((long long)img * ATTN_T + half * 128 + r) * p.out_ld;
#pragma unroll 1
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(t_row + 256u + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
                    uint32_t u[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) u[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                    st_global_256(orow + ch * 32, u); st_global_256(orow + ch * 32 + 16, u + 8);
                }
            }
          }
        } else
        if (mbar_wait(s_full, 0, 16)) {
            tc_fence_after();
            float m = -3.0e38f;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_row + (uint32_t)(ch * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) m = fmaxf(m, __uint_as_float(v[e]));
            }
            const float mc = m * p.scale_log2e;
            float sum = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < 8; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_row + (uint32_t)(ch * 32), v);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0, p1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[2 * e]), p.scale_log2e, -mc)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[2 * e + 1]), p.scale_log2e, -mc)));
                    const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                    const float2 back = __bfloat1622float2(h);            // the sum of what the P.V product will actually see
                    sum += back.x + back.y;
                    pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                uint8_t* rowp = smP + (ch >> 1) * 16384 + r * 128;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(rowp + ((((ch & 1) * 4 + i) ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
            fence_proxy_async_smem();                 // generic-proxy writes of P -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            const float inv = 1.f / sum;
            if (p.pm) {
                // training plans: the backward pass reads the normalised probabilities.  Written from the shared-memory tile
                // while the tensor core runs P.V (these warps would otherwise idle on o_full).
                __nv_bfloat16* prow = p.pm + ((long long)img * ATTN_T + half * 128 + r) * ATTN_T;
#pragma unroll 1
                for (int ch = 0; ch < 8; ++ch) {
                    const uint8_t* rowp = smP + (ch >> 1) * 16384 + r * 128;
                    uint32_t u[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        ld_shared_v4(rowp + ((((ch & 1) * 4 + i) ^ (r & 7)) << 4), u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[e]));
                        u[e] = pack_bf16x2(f.x * inv, f.y * inv);
                    }
                    st_global_256(prow + ch * 32, u); st_global_256(prow + ch * 32 + 16, u + 8);
                }
            }
            if (mbar_wait(o_full, 0, 17)) {
                tc_fence_after();
                __nv_bfloat16* orow = p.out + ((long long)img * ATTN_T + half * 128 + r) * p.out_ld;
#pragma unroll 1
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(t_row + 256u + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
                    uint32_t u[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) u[e] = pack_bf16x2(__uint_as_float(v[2 * e]) * inv, __uint_as_float(v[2 * e + 1]) * inv);
                    st_global_256(orow + ch * 32, u); st_global_256(orow + ch * 32 + 16, u + 8);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct AttnLaunch { CUtensorMap q, k, v; AttnParams p; };

inline bool attn_fused_eligible(int T, int C) { return T == ATTN_T && C == ATTN_D; }

// qkv: bf16 [NB][T][3C] (unet.py:57 chunk order q, k, v); out: bf16 [NB][T][C]
inline int build_attn(const void* qkv, void* out, int NB, int T, int C, AttnLaunch& g, void* pm = nullptr) {
    if (!attn_fused_eligible(T, C)) return fail(-13, "fused attention: T=%d C=%d unsupported (needs T=256, C=256)", T, C);
    memset(&g, 0, sizeof g);
    int rc;
    if ((rc = make_tmap_4d(&g.q, qkv, 3 * C, T, 1, NB, 3 * C, 64, 128, 1, 1))) return rc;
    if ((rc = make_tmap_4d(&g.k, qkv, 3 * C, T, 1, NB, 3 * C, 64, 256, 1, 1))) return rc;
    if ((rc = make_tmap_4d(&g.v, qkv, 3 * C, T, 1, NB, 3 * C, 64, 64, 1, 1))) return rc;
    g.p.NB = NB; g.p.out = reinterpret_cast<__nv_bfloat16*>(out); g.p.pm = reinterpret_cast<__nv_bfloat16*>(pm);
    g.p.scale_log2e = 1.4426950408889634f / sqrtf((float)C); g.p.out_ld = C; g.p.scale = 1.f / sqrtf((float)C);
    return 0;
}
// backward, query side: qkv bf16 [NB][T][3C], dO bf16 [NB][T][C], probs bf16 [NB][T][T] (saved by the forward) ->
// ds bf16 [NB][T][T] and dQ into the q third of dqkv bf16 [NB][T][3C]
inline int build_attn_bwd(const void* qkv, const void* dO, const void* probs, void* ds, void* dqkv, int NB, int T, int C, AttnLaunch& g) {
    if (!attn_fused_eligible(T, C)) return fail(-13, "fused attention backward: T=%d C=%d unsupported (needs T=256, C=256)", T, C);
    memset(&g, 0, sizeof g);
    int rc;
    if ((rc = make_tmap_4d(&g.q, dO, C, T, 1, NB, C, 64, 128, 1, 1))) return rc;
    if ((rc = make_tmap_4d(&g.k, qkv, 3 * C, T, 1, NB, 3 * C, 64, 256, 1, 1))) return rc;
    if ((rc = make_tmap_4d(&g.v, qkv, 3 * C, T, 1, NB, 3 * C, 64, 64, 1, 1))) return rc;
    g.p.NB = NB; g.p.out = reinterpret_cast<__nv_bfloat16*>(dqkv); g.p.out_ld = 3 * C;
    g.p.pm = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(probs)); g.p.ds = reinterpret_cast<__nv_bfloat16*>(ds);
    g.p.scale = 1.f / sqrtf((float)C); g.p.scale_log2e = 0.f;
    return 0;
}
inline int launch_attn(const AttnLaunch& g, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) { DDPM_CUDA_OK(cudaFuncSetAttribute(attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM)); attr_done = true; }
    launch_k(attn_kernel<false>, 2 * g.p.NB, ATTN_THREADS, ATTN_SMEM, st, g.q, g.k, g.v, g.p);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
inline int launch_attn_bwd(const AttnLaunch& g, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) { DDPM_CUDA_OK(cudaFuncSetAttribute(attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM)); attr_done = true; }
    launch_k(attn_kernel<true>, 2 * g.p.NB, ATTN_THREADS, ATTN_SMEM, st, g.q, g.k, g.v, g.p);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ddpm
