// tcgen05 implicit-GEMM engine: one warp-specialised kernel template that serves
//   KK  : A K-major (NHWC pixel tiles, optional 3x3 taps, up to 3 channel segments), B K-major (packed weights / K^T)
//         -> conv3x3 fwd + dgrad, conv1x1, fused conv2+skip, linear, Q.K^T, dO.V^T
//   MNMN: A MN-major, B MN-major, K runs over pixels -> conv wgrad (per tap, split-K, fp32 red), P^T.dO, dS^T.Q
//   KMN : A K-major, B MN-major, K runs over tokens  -> P.V, dS.K
// Tile: M=128 rows x BLOCK_N columns, BLOCK_K=64 bf16 per stage (one 128-byte swizzle span), fp32 accumulators in TMEM.
// Roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps2-5 = epilogue (TMEM -> regs -> global).
#pragma once
#include "ptx.cuh"
#include "gn_epilogue.cuh"

namespace ddpm {

__device__ __forceinline__ const uint32_t* v_as_u32(const float* f) { return reinterpret_cast<const uint32_t*>(f); }

enum GemmMode { GEMM_KK = 0, GEMM_MNMN = 1, GEMM_KMN = 2 };
enum EpiFlags { EPI_OUT_F32 = 1, EPI_ATOMIC = 2 };
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue.  EIGHT epilogue warps (two per TMEM lane quarter, each
// taking every other 32-column chunk): with one warp per scheduler the ~340-instruction chunk body ran at IPC < 0.5 and the
// epilogue, not the MMA, bounded every short-K GEMM (ncu source view, profiles/README.md).
// The MN-major mode (weight gradients: long K, fp32 atomic epilogue, run on the SIDE stream) keeps 4 epilogue warps: at 320
// threads x 144 registers its CTA would leave no room in the register file for a GroupNorm block of the main chain to be
// co-resident on the SM, which is the whole point of the side stream.
__host__ __device__ constexpr int gemm_epi_warps(int mode) { return mode == 1 /*GEMM_MNMN*/ ? 4 : 8; }
__host__ __device__ constexpr int gemm_threads(int mode) { return 64 + 32 * gemm_epi_warps(mode); }

struct GemmSeg {
    int map;       // which A tensor map (0..2)
    int taps;      // number of spatial taps (1..9)
    int kchunks;   // channel chunks of 64 in this segment
    int c_base;    // first channel coordinate inside the map
    int cmul;      // coordinate multiplier of the tile origin (1; 2 for stride-2 maps built with elementStrides=2)
    signed char dx[9], dy[9];   // per-tap coordinate offsets (3x3 pad 1: dx = t%3-1, dy = t/3-1)
};

struct GemmParams {
    int M, N;                 // valid rows / columns of the output (per z slice)
    // row-space (K-major A) or K-space (MN-major operands) geometry of the NHWC tensor behind the 4D maps (C, W, H, N)
    int W, H;                 // spatial dims (plain matrices: W = rows, H = 1)
    int w_t, h_t, n_t;        // TMA box over (W,H,N): product = 128 (K-major A rows) ; MN-major K-block boxes use wk_t,hk_t,nk_t
    int wk_t, hk_t, nk_t;     // product = 64
    // KK
    int nseg; GemmSeg seg[3];
    int b_k_base;             // first K coordinate in the B map
    int a_z_n, b_z;           // per-blockIdx.z increments: A 'n' coordinate, B batch coordinate
    // MNMN / KMN
    int taps;                 // MNMN: 1 or 9 (shift applied to B)
    int splits;               // MNMN: split-K factor; z = (batch*taps + tap)*splits + split
    int kblocks;              // K extent in blocks of 64 (per batch)
    int a_c_base, b_c_base;   // channel-coordinate bases
    int b_cmul, b_pad;        // MNMN: B coordinate = origin*b_cmul + tap - b_pad (b_cmul = 2 for stride-2 convs, maps with elementStrides 2)
    // optional output-row remap (KK): tile row (n,y,x) on the A grid -> output pixel (n, y*o_mul+o_py, x*o_mul+o_px) on an oW x oH grid
    int o_mul, o_py, o_px, oW, oH;
    int m_tiles, n_tiles, grid_z;   // tile space walked by the persistent CTAs
    int kk_splits;                  // KK: > 1 -> z is a K-split index over the flattened (segment, tap, chunk) slab sequence
    int dbg;                        // bottleneck experiments (DDPM_GEMM_DBG): 1 = epilogue drains TMEM but does no math/stores,
                                    // 2 = MMA warp consumes stages without issuing MMAs, 4 = producer signals stages without TMA loads
    // epilogue
    void* out; int ldo; long long out_z_stride; long long out_tap_stride; int flags;
    int tma_store;            // bf16 output goes registers -> swizzled shared memory -> cp.async.bulk.tensor store (tmO)
    const float* bias;        // [N] or null
    const float* rowvec;      // [M/rows_per_vec][rowvec_ld] or null (timestep-embedding projection per image)
    int rowvec_ld, rows_per_vec;
    const __nv_bfloat16* residual; int ldr;   // [M][ldr] or null
    float alpha;
    GnEpi gn; int gn_hw;      // KK: GroupNorm statistics of the output (gn_epilogue.cuh); rows are NHWC pixels, gn_hw = H*W (multiple of 32)
};

// A pipeline stage holds KSTEPS K-slabs of 64 (KSTEPS x {A 16 KB, B BLOCK_N x 128 B}).  One producer/consumer handshake costs
// ~400 cycles of dependent mbarrier / TMA-issue / commit instructions on the single issuing threads (measured: the kernel with
// loads, MMAs and epilogue math disabled still ran at 52% of its full time), more than the 256-cycle MMA time of one N=128
// slab - so N <= 128 configurations move two slabs per handshake.
template <int BLOCK_N, int STAGES, int KSTEPS = 1, int OUT_STAGING = 1>
struct GemmSmem {
    static constexpr int A_BYTES = 128 * 64 * 2;
    static constexpr int B_BYTES = BLOCK_N * 64 * 2;
    static constexpr int SLAB_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGE_BYTES = KSTEPS * SLAB_BYTES;
    // epilogue staging: 2 x {128 rows x 128 B} output slabs for the TMA store + BLOCK_N floats of (bias + per-image vector)
    // (OUT_STAGING = 0 for the MN-major mode: its CTAs run on the side stream and must leave shared memory for the GroupNorm
    // blocks of the main chain to be co-resident; they use direct stores.)
    static constexpr int OUT_STAGE_BYTES = OUT_STAGING ? 128 * 128 : 0;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 2 * OUT_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*epilogue vector*/;
};

__device__ __forceinline__ void pix_decompose(int p, int W, int H, int& n, int& y, int& x) {
    const int hw = W * H;
    n = p / hw;
    const int r = p - n * hw;
    y = r / W;
    x = r - y * W;
}

// Persistent kernel: gridDim.x CTAs walk the tile list t = blockIdx.x, blockIdx.x + gridDim.x, ...
//   t -> (m_tile fastest, n_tile, z) so that CTAs running concurrently share the same weight (B) tile in L2.
// TMEM holds TWO accumulator buffers (2 x BLOCK_N columns): the epilogue warps drain buffer i while the MMA warp already
// accumulates tile i+1 into the other one; the TMA producer runs ahead across tile boundaries.
// (TMA multicast of the weight tile over 2/4-CTA clusters was built and measured no faster in round 1 - L2 traffic is not the
// limiter - and has been removed.)
// PAIR (KK mode): two CTAs of a cluster (tcgen05 cta_group::2) take m_tiles 2j and 2j+1 of the same (n_tile, z): each loads its
// own A rows and HALF of the B tile, the even CTA issues M = 256 MMAs for both (see conv_halo2.cuh for the protocol and why:
// at M = 128 the MMA operand reads saturate the SM's shared-memory bandwidth).  B_ROWS = weight rows held per CTA.
template <int BLOCK_N, int MODE, int STAGES, int KSTEPS, int PAIR = 0>
__global__ void __launch_bounds__(gemm_threads(MODE), 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
    pdl_trigger();                                  // dependents may be scheduled; they block in their own pdl_wait()
    static_assert(!PAIR || MODE != GEMM_KMN, "CTA pairs are implemented for the K-major and the MN-major (weight-gradient) modes");
    constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;
    using SM = GemmSmem<B_ROWS, STAGES, KSTEPS, (MODE == GEMM_MNMN ? 0 : 1)>;
    constexpr int A_MN = (MODE == GEMM_MNMN) ? 1 : 0;
    constexpr int B_MN = (MODE == GEMM_KK) ? 0 : 1;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;     // 128 / 256 / 512: powers of two

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* out_stage = smem + STAGES * SM::STAGE_BYTES;                  // [2][128 rows][128 B], 1024-aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 2 * SM::OUT_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_vec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);   // [BLOCK_N] bias (+ per-image vector) of the current tile
    const uint32_t s_vec_u32 = smem_u32(s_vec);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader of the pair
    // tile walk: single CTA: t -> (m_tile fastest, n_tile, z); pair: t indexes PAIRS of m_tiles (m_tile = 2*(t % m_pairs) + rank)
    const int m_tiles = PAIR ? (p.m_tiles >> 1) : p.m_tiles, n_tiles = p.n_tiles;
    const int total_tiles = m_tiles * n_tiles * p.grid_z;
    const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, cta_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    // K slabs of a tile (uniform across roles)
    int kk_slabs = 0;
    if (MODE == GEMM_KK) for (int s = 0; s < p.nseg; ++s) kk_slabs += p.seg[s].taps * p.seg[s].kchunks;
    const int per_split = (MODE == GEMM_MNMN) ? (p.kblocks + p.splits - 1) / p.splits : 0;
    const int kk_per = (MODE == GEMM_KK && p.kk_splits > 1) ? (kk_slabs + p.kk_splits - 1) / p.kk_splits : kk_slabs;
    auto slabs_of = [&](int z) -> int {
        if (MODE == GEMM_KK) {
            if (p.kk_splits <= 1) return kk_slabs;
            const int lo = z * kk_per; int hi = lo + kk_per; if (hi > kk_slabs) hi = kk_slabs;
            return hi > lo ? hi - lo : 0;
        }
        if (MODE == GEMM_KMN) return p.kblocks;
        const int split = z % p.splits;
        const int kb0 = split * per_split;
        int kb1 = kb0 + per_split; if (kb1 > p.kblocks) kb1 = p.kblocks;
        return kb1 > kb0 ? kb1 - kb0 : 0;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], (PAIR ? 2 : 1) * gemm_epi_warps(MODE)); }
        fence_mbar_init();
    }
    // pair: both CTAs are synchronised BEFORE the collective TMEM allocation (and the peer's barriers are initialised before
    // anything arrives on them)
    if (PAIR) cluster_sync_all();
    if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                     // prologue above overlapped the previous kernel's tail; its data is visible from here
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= TMA producer =======================
        // elect.sync from the converged warp (not `lane == 0`): the compiler then knows a single thread runs the
        // uniform-datapath instructions (UTMALDG / UTCHMMA) and emits them straight-line; under a plain lane test it
        // wraps every one of them in an ELECT/branch loop (~60-100 cycles per MMA, measured as a 2x loss at N=128).
        if (elect_one()) {
            int stg = 0;                              // global stage counter
            int sub = 0;                              // slabs already issued into the current stage
            int left = 0;                             // slabs of the current tile not yet issued
            bool ok = true;
            uint64_t* fb = nullptr;
            // acquire the slot for the next slab (waits for the stage when it starts one); returns the slab's smem base
            auto acquire = [&](int) -> uint8_t* {
                const int st = stg % STAGES;
                if (sub == 0) {
                    const uint32_t ph = (stg / STAGES) & 1;
                    if (!mbar_wait(&empty_bar[st], ph ^ 1, 1)) { ok = false; return nullptr; }
                }
                fb = &full_bar[st];
                return smem + st * SM::STAGE_BYTES + sub * SM::SLAB_BYTES;
            };
            // after the loads of a slab were issued: arm the barrier when the stage is complete (or the tile ends)
            auto commit_slab = [&]() {
                ++sub; --left;
                if (sub == KSTEPS || left == 0) { if (crank == 0) mbar_expect_tx(fb, (uint32_t)((PAIR ? 2 : 1) * sub * SM::SLAB_BYTES)); sub = 0; ++stg; }
            };
            for (int t = cta_id; t < total_tiles && ok; t += cta_step) {
                const int m_tile = PAIR ? (t % m_tiles) * 2 + (int)crank : t % m_tiles, n_tile = (t / m_tiles) % n_tiles, z = t / (m_tiles * n_tiles);
                left = slabs_of(z); sub = 0;
                if (MODE == GEMM_KK) {
                    int n0, y0, x0;
                    pix_decompose(m_tile * 128, p.W, p.H, n0, y0, x0);
                    n0 += z * p.a_z_n;
                    const int k_lo = p.kk_splits > 1 ? z * kk_per : 0;
                    const int k_hi = p.kk_splits > 1 ? (k_lo + kk_per < kk_slabs ? k_lo + kk_per : kk_slabs) : kk_slabs;
                    int kcount = 0;
                    for (int s = 0; s < p.nseg && ok; ++s) {
                        const GemmSeg sg = p.seg[s];
                        const CUtensorMap* mA = sg.map == 0 ? &tmA0 : (sg.map == 1 ? &tmA1 : &tmA2);
                        for (int tp = 0; tp < sg.taps && ok; ++tp) {
                            const int xc = x0 * sg.cmul + sg.dx[tp];
                            const int yc = y0 * sg.cmul + sg.dy[tp];
                            for (int kc = 0; kc < sg.kchunks; ++kc, ++kcount) {
                                if (kcount < k_lo || kcount >= k_hi) continue;
                                uint8_t* st = acquire(0);
                                if (!st) break;
                                if (p.dbg & 4) { asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(fb)), "r"((uint32_t)SM::SLAB_BYTES) : "memory"); commit_slab(); continue; }
                                if (PAIR) {
                                    tma_load_4d_2cta(st, mA, fb, sg.c_base + kc * 64, xc, yc, n0);
                                    tma_load_3d_2cta(st + SM::A_BYTES, &tmB, fb, p.b_k_base + kcount * 64, n_tile * BLOCK_N + (int)crank * B_ROWS, z * p.b_z);
                                } else {
                                    tma_load_4d(st, mA, fb, sg.c_base + kc * 64, xc, yc, n0);
                                    tma_load_3d(st + SM::A_BYTES, &tmB, fb, p.b_k_base + kcount * 64, n_tile * BLOCK_N, z * p.b_z);
                                }
                                commit_slab();
                            }
                        }
                    }
                } else if (MODE == GEMM_MNMN) {
                    const int split = z % p.splits, tap = (z / p.splits) % p.taps, batch = z / (p.splits * p.taps);
                    const int kb0 = split * per_split;
                    const int ns = slabs_of(z);
                    const int dx = p.taps == 9 ? (tap % 3) - p.b_pad : 0;
                    const int dy = p.taps == 9 ? (tap / 3) - p.b_pad : 0;
                    for (int i = 0; i < ns; ++i) {
                        uint8_t* st = acquire(0);
                        if (!st) break;
                        int n0, y0, x0;
                        pix_decompose((kb0 + i) * 64, p.W, p.H, n0, y0, x0);
                        n0 += batch;
                        if (PAIR) {      // own 128 rows of M, own half of the N columns; bytes complete on the leader's barrier
#pragma unroll
                            for (int b = 0; b < 2; ++b)
                                tma_load_4d_2cta(st + b * 8192, &tmA0, fb, p.a_c_base + m_tile * 128 + b * 64, x0, y0, n0);
#pragma unroll
                            for (int b = 0; b < B_ROWS / 64; ++b)
                                tma_load_4d_2cta(st + SM::A_BYTES + b * 8192, &tmB, fb, p.b_c_base + n_tile * BLOCK_N + (int)crank * B_ROWS + b * 64,
                                                 x0 * p.b_cmul + dx, y0 * p.b_cmul + dy, n0);
                        } else {
#pragma unroll
                            for (int b = 0; b < 2; ++b)
                                tma_load_4d(st + b * 8192, &tmA0, fb, p.a_c_base + m_tile * 128 + b * 64, x0, y0, n0);
#pragma unroll
                            for (int b = 0; b < BLOCK_N / 64; ++b)
                                tma_load_4d(st + SM::A_BYTES + b * 8192, &tmB, fb, p.b_c_base + n_tile * BLOCK_N + b * 64,
                                            x0 * p.b_cmul + dx, y0 * p.b_cmul + dy, n0);
                        }
                        commit_slab();
                    }
                } else {  // GEMM_KMN
                    int n0, y0, x0;
                    pix_decompose(m_tile * 128, p.W, p.H, n0, y0, x0);
                    n0 += z * p.a_z_n;
                    for (int i = 0; i < p.kblocks; ++i) {
                        uint8_t* st = acquire(0);
                        if (!st) break;
                        tma_load_4d(st, &tmA0, fb, p.a_c_base + i * 64, x0, y0, n0);
#pragma unroll
                        for (int b = 0; b < BLOCK_N / 64; ++b)
                            tma_load_4d(st + SM::A_BYTES + b * 8192, &tmB, fb, p.b_c_base + n_tile * BLOCK_N + b * 64, i * 64, 0, z);
                        commit_slab();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (crank == 0 && elect_one()) {             // pair: the leader issues for both CTAs
            constexpr uint32_t idesc = umma_idesc(PAIR ? 256 : 128, BLOCK_N, A_MN, B_MN);
            int stg = 0, it = 0;
            bool ok = true;
            for (int t = cta_id; t < total_tiles && ok; t += cta_step, ++it) {
                const int z = t / (m_tiles * n_tiles);
                const int ns = slabs_of(z);
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                if (!mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 4)) break;      // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int i = 0; i < ns; i += KSTEPS, ++stg) {
                    const int st = stg % STAGES;
                    const uint32_t ph = (stg / STAGES) & 1;
                    if (!mbar_wait(&full_bar[st], ph, 2)) { ok = false; break; }
                    tc_fence_after();
                    const int nsub = (ns - i) < KSTEPS ? (ns - i) : KSTEPS;
#pragma unroll
                    for (int sb = 0; sb < KSTEPS; ++sb) {
                        if (sb >= nsub) break;
                        const uint32_t a_addr = smem_u32(smem + st * SM::STAGE_BYTES + sb * SM::SLAB_BYTES);
                        const uint32_t b_addr = a_addr + SM::A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // K-major: 16 elements = 32 B inside the swizzle span; MN-major: 16 K-rows of 128 B
                            const uint64_t da = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                                     : umma_smem_desc(a_addr + k * 32, 16, 1024);
                            const uint64_t db = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                                     : umma_smem_desc(b_addr + k * 32, 16, 1024);
                            if (!(p.dbg & 2)) { if (PAIR) umma_bf16_2cta(d_tmem, da, db, idesc, (i | sb | k) != 0); else umma_bf16(d_tmem, da, db, idesc, (i | sb | k) != 0); }
                        }
                    }
                    if (PAIR) umma_commit_2cta(&empty_bar[st], 3); else umma_commit(&empty_bar[st]);
                }
                if (ok) { if (PAIR) umma_commit_2cta(&tmem_full[acc], 3); else umma_commit(&tmem_full[acc]); }
            }
        }
    } else {
        // ======================= epilogue: 8 warps, 2 per TMEM lane quarter =======================
        const int q = warp & 3;
        constexpr int NGRP = gemm_epi_warps(MODE) / 4;   // warp groups per TMEM lane quarter (1 or 2)
        const int grp = (warp - 2) >> 2;             // which 32-column part of every (32*NGRP)-column slab this warp converts
        const int r = q * 32 + lane;                 // accumulator row
        int it = 0;
        uint32_t slab_ctr = 0;                       // staging-buffer parity of the TMA-store path, continues across tiles
        for (int t = cta_id; t < total_tiles; t += cta_step, ++it) {
            const int m_tile = PAIR ? (t % m_tiles) * 2 + (int)crank : t % m_tiles, n_tile = (t / m_tiles) % n_tiles, z = t / (m_tiles * n_tiles);
            const int ns = slabs_of(z);
            const int acc = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            if (!mbar_wait(&tmem_full[acc], acc_ph, 3)) break;
            tc_fence_after();
            int tap = 0, batch = z;
            if (MODE == GEMM_MNMN) { tap = (z / p.splits) % p.taps; batch = z / (p.splits * p.taps); }
            const int row = m_tile * 128 + r;
            const bool row_ok = row < p.M && ns > 0;
            long long orow = row;                         // output row (pixel) index
            if (MODE == GEMM_KK && p.o_mul > 1) {
                int n_, y_, x_;
                pix_decompose(row, p.W, p.H, n_, y_, x_);
                orow = ((long long)n_ * p.oH + y_ * p.o_mul + p.o_py) * p.oW + x_ * p.o_mul + p.o_px;
            }
            const long long zoff = (MODE == GEMM_MNMN) ? (long long)batch * p.out_z_stride + (long long)tap * p.out_tap_stride
                                                      : ((MODE == GEMM_KK && p.kk_splits > 1) ? 0 : (long long)z * p.out_z_stride);
            // Bias and (when every row of the tile belongs to the same image) the per-image vector are staged ONCE per tile in
            // shared memory: fetched per chunk with global loads they cost as much as the stores (measured, K=256 GEMMs:
            // 35.0 us -> 24.0 us without the bias loads, 23.3 us without the stores, 13.8 us without either).
            const bool vec_uniform = p.rowvec != nullptr && (p.rows_per_vec % 128) == 0;
            const float* rv = (p.rowvec && !vec_uniform && row_ok) ? p.rowvec + (long long)(row / p.rows_per_vec) * p.rowvec_ld : nullptr;
            const int tid_epi = (int)threadIdx.x - 64;
            constexpr int EPI_T = 32 * gemm_epi_warps(MODE);
            named_bar_sync(1, EPI_T);                                 // the previous tile's readers are done with s_vec
            for (int c = tid_epi; c < BLOCK_N; c += EPI_T) {
                const int col = n_tile * BLOCK_N + c;
                float v = 0.f;
                if (col < p.N && ns > 0) {
                    if (p.bias) v = __ldg(p.bias + col);
                    if (vec_uniform) v += __ldg(p.rowvec + (long long)((m_tile * 128) / p.rows_per_vec) * p.rowvec_ld + col);
                }
                s_vec[c] = v;
            }
            named_bar_sync(1, EPI_T);
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            const bool tma_out = p.tma_store != 0 && NGRP == 2;
#pragma unroll 1
            for (int s0 = 0; s0 < BLOCK_N; s0 += 32 * NGRP) {         // one slab per iteration: 64 columns (2 groups) or 32 (1 group)
                if (n_tile * BLOCK_N + s0 >= p.N || ns == 0) break;   // uniform across the CTA
                const int c0 = s0 + grp * 32;
                const int col = n_tile * BLOCK_N + c0;
                const bool active = col < p.N;                        // N % 64 == 32: the last slab has one chunk only
                uint32_t v[32];
                if (active) { tmem_ld32(t_addr + (uint32_t)c0, v); tmem_ld_wait(); }
                if (p.dbg & 1) continue;
                if (active && (row_ok || tma_out)) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 b4;                                          // broadcast read (explicit ld.shared: the carved pointer is generic)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(s_vec_u32 + (uint32_t)(c0 + 4 * j) * 4u));
                    f[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.alpha, b4.x); f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.alpha, b4.y);
                    f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.alpha, b4.z); f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.alpha, b4.w);
                }
                if (rv) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + col) + j);
                        f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w; }
                }
                if (p.residual && row_ok) {
                    uint32_t rs[16];
                    ld_row64B(p.residual + orow * p.ldr + col, rs);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rs[e]));
                        f[2 * e] += t2.x; f[2 * e + 1] += t2.y;
                    }
                }
                if (MODE == GEMM_KK && p.gn.qstats) {
                    // GroupNorm statistics of the output (warp-uniform: the 32 rows of a warp lie in one image, M % 32 == 0;
                    // with the output-row scatter of the sub-pixel convs the image index is still that of the A-grid row)
                    const int n_img = row_ok ? row / p.gn_hw : 0;
                    epi_quad_stats(f, row_ok, p.gn.qstats + ((long long)n_img * (p.N >> 2) + (col >> 2)) * 2, lane);
                }
                if (tma_out && NGRP == 2) {
                    // registers -> 128B-swizzled staging slab (row r, 16-byte chunk j at physical chunk j ^ (r & 7))
                    uint8_t* buf = out_stage + (slab_ctr & 1) * SM::OUT_STAGE_BYTES + r * 128;
                    const int jb = grp * 4;                          // which half of the 64-column slab
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        st_shared_v4(buf + (((jb + j) ^ (r & 7)) << 4), pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                     pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                } else if (p.flags & EPI_ATOMIC) {
                    float* o = reinterpret_cast<float*>(p.out) + zoff + orow * p.ldo + col;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(f[j]), "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3]) : "memory");
                } else if (p.flags & EPI_OUT_F32) {
                    float* o = reinterpret_cast<float*>(p.out) + zoff + orow * p.ldo + col;
                    if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) st_global_256(o + j * 8, v_as_u32(f + j * 8));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + zoff + orow * p.ldo + col;
                    uint32_t u[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) u[e] = pack_bf16x2(f[e * 2], f[e * 2 + 1]);
                    if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) { st_global_256(o, u); st_global_256(o + 16, u + 8); }   // two full sectors per store
                    else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(o)[j] = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                    }
                }
                }
                if (tma_out && NGRP == 2) {
                    // one barrier per slab: before it, the issuing thread has waited until the PREVIOUS store finished reading
                    // its buffer (the one the next slab will overwrite); after it, every lane's rows are in the staging slab
                    fence_proxy_async_smem();
                    if (tid_epi == 0) bulk_wait_group_read0();
                    named_bar_sync(1, EPI_T);
                    if (tid_epi == 0) {
                        tma_store_3d(&tmO, out_stage + (slab_ctr & 1) * SM::OUT_STAGE_BYTES, n_tile * BLOCK_N + s0, m_tile * 128, z);
                        bulk_commit_group();
                    }
                    ++slab_ctr;
                }
            }
            // hand the accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }   // pair: on the leader's barrier
        }
        if (threadIdx.x == 64) bulk_wait_group0();  // outstanding TMA stores read this CTA's shared memory
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                   // no arrival / MMA read may still target the peer's shared memory or TMEM
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace ddpm
