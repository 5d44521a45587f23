// CTA-PAIR variant of the haloed 3x3 convolution (conv_halo.cuh): tcgen05.mma.cta_group::2, M = 256 pixels per instruction.
//
// Why: at M = 128 the operands of every MMA are re-read from shared memory at 128 x 16 x 2 B (A) + BLOCK_N x 16 x 2 B (B) per
// 64-cycle (N = 128) instruction = 128 B/clk - the whole shared-memory bandwidth of the SM - while TMA has to WRITE the same
// weight bytes into that shared memory.  The single-CTA kernel therefore sits at ~0.52 of the tensor pipe's nominal rate for
// N = 128 (0.66 of the measured cuBLAS peak) and 0.68 for N = 256, whatever the pipeline depth (round-1 / round-2 measurements,
// profiles/README.md).  With a CTA pair each SM keeps its own 16 x 8 pixel patch (A) but holds only HALF of the weight tile:
// B reads and B TMA writes per SM are halved, and the weight traffic out of L2 halves with them.
//
// Layout of the pair: CTAs 2i and 2i+1 of a 2-CTA cluster take m-tiles 2t and 2t+1 of the same n-tile.  Each loads its own haloed
// A box and rows [rank*BLOCK_N/2, (rank+1)*BLOCK_N/2) of every weight slab to the SAME shared-memory offsets; all TMA bytes
// complete on the leader's (even CTA's) mbarriers; the leader's MMA thread issues for both; tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; both CTAs run their own epilogue on their own TMEM rows and hand the
// accumulator back by arriving on the leader's tmem_empty barrier (remote arrive for the odd CTA).
#pragma once
#include "conv_halo.cuh"

namespace ddpm {

template <int BLOCK_N>
struct Halo2Cfg {
    static constexpr int P = 10;
    static constexpr int A_BYTES = 18 * P * 128;
    static constexpr int A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;
    static constexpr int HALF_N = BLOCK_N / 2;                    // weight rows held by each CTA of the pair
    static constexpr int TPS = (BLOCK_N == 128) ? 3 : 1;           // taps per weight stage
    static constexpr int TAP_BYTES = HALF_N * 128;
    static constexpr int B_BYTES = TAP_BYTES * TPS;                // 24 KB (N=128) / 16 KB (N=256)
    static constexpr int NA = 3;
    static constexpr int NB_ST = (BLOCK_N == 128) ? 4 : 6;         // 96 KB of weight stages
    static constexpr int TMEM_COLS = 2 * BLOCK_N;                  // two accumulators per CTA
    static constexpr int OUT_STAGE_BYTES = 128 * 128;
    static constexpr int TOTAL = NA * A_STRIDE + NB_ST * B_BYTES + 2 * OUT_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 1024 /*bias vector*/;
};

// XF = 1 (inference plans): the 3x3 segment's input is the RAW GroupNorm input x; four extra "transform" warps rewrite every
// haloed A tile in shared memory between its TMA landing and the MMAs: y = silu(sc*x + sh) with the per-(image, channel)
// constants K = {sc, sh} of the norm (out-of-image halo pixels stay zero = the conv's padding of the NORMALISED tensor).
// The elementwise rewrite keeps the 128-byte swizzle (it is in place), so the UMMA descriptors are unchanged.  This removes the
// GroupNorm-apply pass (read x + write a, 4 B/element) and the `a` tensor of unet.py:83-89 from the sampler's forward.
// Protocol: with XF each CTA's A box completes on its OWN full_a barrier; its transform warps wait there, rewrite, fence the
// generic-proxy writes for the async proxy and arrive on the LEADER's ready_a barrier (2 CTAs x 4 warps), which the MMA thread
// waits on instead of full_a.
constexpr int HALO_XF_WARPS = 8;       // 4 warps rewrote a 23 KB tile in ~2600 cycles > the 2304 MMA cycles of an N=128 chunk (measured: conv 2x slower)
template <int BLOCK_N, int XF>
__global__ void __launch_bounds__(HALO_THREADS + 32 * HALO_XF_WARPS * XF, 1)
conv3x3_halo2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const HaloParams p) {
    pdl_trigger();
    using CF = Halo2Cfg<BLOCK_N>;
    constexpr int P = CF::P;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smem + CF::NA * CF::A_STRIDE;
    uint8_t* out_stage = smB + CF::NB_ST * CF::B_BYTES;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(out_stage + 2 * CF::OUT_STAGE_BYTES);
    uint64_t* empty_a = full_a + CF::NA;
    uint64_t* full_b = empty_a + CF::NA;
    uint64_t* empty_b = full_b + CF::NB_ST;
    uint64_t* tmem_full = empty_b + CF::NB_ST;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]  (the LEADER's copy collects the arrivals of both CTAs)
    uint64_t* ready_a = tmem_empty + 2;             // [NA]  XF: "tile rewritten by the transform warps of both CTAs" (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready_a + CF::NA);
    float* s_vec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_a) + 512);
    const uint32_t s_vec_u32 = smem_u32(s_vec);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();        // 0 = leader
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int m_tiles = p.NB * tiles_per_img;       // even (checked on the host)
    const int m_pairs = m_tiles >> 1;
    const int total_pairs = m_pairs * p.n_tiles;
    const int pair_id = blockIdx.x >> 1, n_pairs_grid = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB);
        for (int s = 0; s < CF::NA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); mbar_init(&ready_a[s], 2 * HALO_XF_WARPS); }
        for (int s = 0; s < CF::NB_ST; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * HALO_EPI_WARPS); }
        fence_mbar_init();
    }
    cluster_sync_all();                              // both CTAs synchronised before the collective TMEM allocation; the peer's
                                                     // barriers are initialised before anything arrives on them
    if (warp == 1) tmem_alloc2(tmem_slot, CF::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= TMA producer (both CTAs) =======================
        if (elect_one()) {
            int ia = 0, ib = 0;
            bool ok = true;
            for (int t = pair_id; t < total_pairs && ok; t += n_pairs_grid) {
                const int m_tile = (t % m_pairs) * 2 + (int)rank, n_tile = t / m_pairs;
                const int n = m_tile / tiles_per_img, r = m_tile % tiles_per_img;
                const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
                int kbase = 0;
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const HaloSeg sg = p.seg[s];
                    const CUtensorMap* mA = sg.map == 0 ? &tmA0 : (sg.map == 1 ? &tmA1 : &tmA2);
                    const int Cseg = sg.kchunks * 64;
                    for (int kc = 0; kc < sg.kchunks && ok; ++kc) {
                        {
                            const int sa = ia % CF::NA; const uint32_t ph = (ia / CF::NA) & 1;
                            if (!mbar_wait(&empty_a[sa], ph ^ 1, 5)) { ok = false; break; }
                            const uint32_t bytes = (sg.taps == 9) ? (uint32_t)CF::A_BYTES : (uint32_t)(16 * 8 * 128);
                            if (XF) {            // own barrier: the local transform warps pick the tile up
                                mbar_expect_tx(&full_a[sa], bytes);
                                if (sg.taps == 9) tma_load_4d(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0 - 1, y0 - 1, n);
                                else              tma_load_4d(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0, y0, n);
                            } else {
                                if (rank == 0) mbar_expect_tx(&full_a[sa], 2 * bytes);       // both CTAs' boxes complete here
                                if (sg.taps == 9) tma_load_4d_2cta(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0 - 1, y0 - 1, n);
                                else              tma_load_4d_2cta(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0, y0, n);
                            }
                            ++ia;
                        }
                        for (int tp = 0; tp < sg.taps; tp += CF::TPS) {
                            const int nt = (sg.taps - tp < CF::TPS) ? sg.taps - tp : CF::TPS;
                            const int sb = ib % CF::NB_ST; const uint32_t ph = (ib / CF::NB_ST) & 1;
                            if (!mbar_wait(&empty_b[sb], ph ^ 1, 6)) { ok = false; break; }
                            if (rank == 0) mbar_expect_tx(&full_b[sb], (uint32_t)(2 * nt * CF::TAP_BYTES));
                            for (int j = 0; j < nt; ++j)
                                tma_load_3d_2cta(smB + sb * CF::B_BYTES + j * CF::TAP_BYTES, &tmB, &full_b[sb], kbase + (tp + j) * Cseg + kc * 64,
                                                 n_tile * BLOCK_N + (int)rank * CF::HALF_N, 0);
                            ++ib;
                        }
                    }
                    kbase += sg.taps * Cseg;
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (leader CTA only) =======================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc(256, BLOCK_N, 0, 0);
            int ia = 0, ib = 0, it = 0;
            bool ok = true;
            for (int t = pair_id; t < total_pairs && ok; t += n_pairs_grid, ++it) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                if (!mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 4)) break;       // both CTAs' epilogues have drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                bool first = true;
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const HaloSeg sg = p.seg[s];
                    for (int kc = 0; kc < sg.kchunks && ok; ++kc) {
                        const int sa = ia % CF::NA; const uint32_t pha = (ia / CF::NA) & 1;
                        if (!mbar_wait(XF ? &ready_a[sa] : &full_a[sa], pha, 7)) { ok = false; break; }
                        const uint32_t a_base = smem_u32(smA + sa * CF::A_STRIDE);
                        for (int tp0 = 0; tp0 < sg.taps; tp0 += CF::TPS) {
                            const int nt = (sg.taps - tp0 < CF::TPS) ? sg.taps - tp0 : CF::TPS;
                            const int sb = ib % CF::NB_ST; const uint32_t phb = (ib / CF::NB_ST) & 1;
                            if (!mbar_wait(&full_b[sb], phb, 2)) { ok = false; break; }
                            tc_fence_after();
                            const int pitch = (sg.taps == 9) ? P : 8;
#pragma unroll
                            for (int j = 0; j < CF::TPS; ++j) {
                                if (j >= nt) break;
                                const int tp = tp0 + j;
                                const uint32_t b_addr = smem_u32(smB + sb * CF::B_BYTES + j * CF::TAP_BYTES);
                                const int row0 = (sg.taps == 9) ? (tp / 3) * P + (tp % 3) : 0;
                                const uint32_t a_row = a_base + (uint32_t)row0 * 128u;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t da = umma_smem_desc(a_row + k * 32, 16, (uint32_t)pitch * 128u);
                                    const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024);
                                    umma_bf16_2cta(d_tmem, da, db, idesc, (first && k == 0) ? 0u : 1u);
                                }
                                first = false;
                            }
                            umma_commit_2cta(&empty_b[sb], 3);
                            ++ib;
                        }
                        umma_commit_2cta(&empty_a[sa], 3);
                        ++ia;
                    }
                }
                if (ok) umma_commit_2cta(&tmem_full[acc], 3);
            }
        }
    } else if (XF && warp >= 2 + HALO_EPI_WARPS) {
        // ======================= transform warps (both CTAs): GroupNorm + SiLU on the landed A tile =======================
        const int tid_x = (warp - 2 - HALO_EPI_WARPS) * 32 + lane;      // 0..255
        constexpr int XROWS = 4 * HALO_XF_WARPS;                        // tile rows covered per pass (8 threads per 128-byte row)
        const int c = tid_x & 7, r0 = tid_x >> 3;                      // logical 16-byte chunk (8 channels), first tile row
        int ia = 0;
        bool ok = true;
        for (int t = pair_id; t < total_pairs && ok; t += n_pairs_grid) {
            const int m_tile = (t % m_pairs) * 2 + (int)rank;
            const int n = m_tile / tiles_per_img, r = m_tile % tiles_per_img;
            const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
            for (int s = 0; s < p.nseg && ok; ++s) {
                const HaloSeg sg = p.seg[s];
                const bool do_xf = sg.taps == 9 && sg.map == 0 && p.xfK != nullptr;
                for (int kc = 0; kc < sg.kchunks && ok; ++kc, ++ia) {
                    const int sa = ia % CF::NA; const uint32_t ph = (ia / CF::NA) & 1;
                    float sc[8], sh[8];
                    if (do_xf) {         // the tile's image and this thread's 8 channels: issued before the wait (hides behind the TMA)
                        const float* Kn = p.xfK + (long long)n * 4 * p.xfC + sg.c_base + kc * 64 + c * 8;
                        const float4 a0 = __ldg(reinterpret_cast<const float4*>(Kn)), a1 = __ldg(reinterpret_cast<const float4*>(Kn) + 1);
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(Kn + p.xfC)), b1 = __ldg(reinterpret_cast<const float4*>(Kn + p.xfC) + 1);
                        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
                    }
                    if (!mbar_wait(&full_a[sa], ph, 8)) { ok = false; break; }
                    if (do_xf && !(p.xf_silu & 4)) {          // (bit 2: experiment knob - protocol only, no rewrite)
                        uint8_t* base = smA + sa * CF::A_STRIDE;
#pragma unroll
                        for (int rr = r0; rr < 18 * P; rr += XROWS) {
                            const int yy = rr / P, xx = rr - yy * P;
                            const int gy = y0 - 1 + yy, gx = x0 - 1 + xx;
                            if (gy < 0 || gy >= p.H || gx < 0 || gx >= p.W) continue;      // zero padding of the normalised tensor
                            uint4* ptr = reinterpret_cast<uint4*>(base + rr * 128 + ((c ^ (rr & 7)) << 4));
                            uint4 u = *ptr;
                            __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 v = __bfloat1622float2(hp[e]);
                                float y0f = fmaf(v.x, sc[2 * e], sh[2 * e]), y1f = fmaf(v.y, sc[2 * e + 1], sh[2 * e + 1]);
                                if (p.xf_silu & 1) {
                                    float t0, t1;
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * y0f));
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * y1f));
                                    y0f *= fmaf(0.5f, t0, 0.5f); y1f *= fmaf(0.5f, t1, 0.5f);
                                }
                                hp[e] = __floats2bfloat162_rn(y0f, y1f);
                            }
                            *ptr = u;
                        }
                        if (!(p.xf_silu & 8)) fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&ready_a[sa], 0);
                }
            }
        }
    } else {
        // ======================= epilogue (both CTAs, own TMEM rows) =======================
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const int tid_epi = (int)threadIdx.x - 64;
        constexpr int EPI_T = 32 * HALO_EPI_WARPS;
        constexpr int NSLAB = BLOCK_N / 64;
        int it = 0;
        uint32_t slab_ctr = 0;
        for (int t = pair_id; t < total_pairs; t += n_pairs_grid, ++it) {
            const int m_tile = (t % m_pairs) * 2 + (int)rank, n_tile = t / m_pairs;
            const int n = m_tile / tiles_per_img, rr = m_tile % tiles_per_img;
            const int y0 = (rr / p.tiles_x) * 16, x0 = (rr % p.tiles_x) * 8;
            const int acc = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            if (!mbar_wait(&tmem_full[acc], acc_ph, 3)) break;
            tc_fence_after();
            named_bar_sync(1, EPI_T);
            for (int c = tid_epi; c < BLOCK_N; c += EPI_T) {
                const int col = n_tile * BLOCK_N + c;
                float v = 0.f;
                if (col < p.N) {
                    if (p.bias) v = __ldg(p.bias + col);
                    if (p.rowvec) v += __ldg(p.rowvec + (long long)n * p.rowvec_ld + col);
                }
                s_vec[c] = v;
            }
            named_bar_sync(1, EPI_T);
            const long long pix = ((long long)n * p.H + y0 + (r >> 3)) * p.W + x0 + (r & 7);
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
            for (int i2 = 0; i2 < NSLAB; ++i2) {
                const int s0 = i2 * 64;
                const int c0 = s0 + grp * 32;
                const int col = n_tile * BLOCK_N + c0;
                uint32_t v[32];
                tmem_ld32(t_addr + (uint32_t)c0, v);
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 b4;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(s_vec_u32 + (uint32_t)(c0 + 4 * j) * 4u));
                    f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
                }
                if (p.residual) {
                    uint32_t rr2[16];
                    ld_row64B(p.residual + pix * p.ldr + col, rr2);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr2[e]));
                        f[2 * e] += t2.x; f[2 * e + 1] += t2.y;
                    }
                }
                if (p.gn.qstats) epi_quad_stats(f, true, p.gn.qstats + ((long long)n * (p.N >> 2) + (col >> 2)) * 2, lane);
                uint8_t* buf = out_stage + (slab_ctr & 1) * CF::OUT_STAGE_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(buf + (((grp * 4 + j) ^ (r & 7)) << 4), pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                fence_proxy_async_smem();
                if (tid_epi == 0) bulk_wait_group_read0();
                named_bar_sync(1, EPI_T);
                if (tid_epi == 0) {
                    tma_store_4d(&tmO, out_stage + (slab_ctr & 1) * CF::OUT_STAGE_BYTES, n_tile * BLOCK_N + s0, x0, y0, n);
                    bulk_commit_group();
                }
                ++slab_ctr;
            }
            // hand the accumulator buffer back: every epilogue warp of BOTH CTAs arrives on the leader's barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
        }
        if (threadIdx.x == 64) bulk_wait_group0();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // no arrival / MMA read may still target the peer's shared memory or TMEM
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, CF::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
inline bool halo_pair_eligible(int NB, int H, int W, int Cout) {
    if (!halo_eligible(H, W, Cout)) return false;
    const int bn = pick_block_n(Cout);
    const long long m_tiles = (long long)NB * (H / 16) * (W / 8);
    return bn >= 128 && (m_tiles % 2) == 0;
}

template <int BLOCK_N, int XF>
inline int launch_halo2_inst(const HaloLaunch& g, cudaStream_t st) {
    using CF = Halo2Cfg<BLOCK_N>;
    auto kern = conv3x3_halo2_kernel<BLOCK_N, XF>;
    static_assert(CF::TOTAL <= 232448, "halo pair conv: shared memory budget (227 KB) exceeded");
    static bool attr_done = false;
    if (!attr_done) { DDPM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::TOTAL)); attr_done = true; }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    const int pairs_total = g.tiles / 2;
    int pairs = num_sms / 2; if (pairs > pairs_total) pairs = pairs_total;
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(HALO_THREADS + 32 * HALO_XF_WARPS * XF); cfg.dynamicSmemBytes = CF::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    bool pdl = pdl_enabled();
    if (pdl) { cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone; if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone && !getenv("DDPM_PDL_GRAPH")) pdl = false; }
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    DDPM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g.a[0], g.a[1], g.a[2], g.b, g.o, g.p));
    return 0;
}
inline int launch_halo2(const HaloLaunch& g, cudaStream_t st) {
    if (g.p.xfK) {
        if (g.block_n == 128) return launch_halo2_inst<128, 1>(g, st);
        if (g.block_n == 256) return launch_halo2_inst<256, 1>(g, st);
    } else {
        if (g.block_n == 128) return launch_halo2_inst<128, 0>(g, st);
        if (g.block_n == 256) return launch_halo2_inst<256, 0>(g, st);
    }
    return fail(-6, "unsupported halo pair conv variant");
}

}  // namespace ddpm
