// 3x3 stride-1 convolution with a HALOED input tile (tcgen05 + TMA), the kernel behind ResidualBlock conv1 / conv2(+skip)
// and their data gradients at resolutions >= 16x16.
//
// The generic engine (umma_gemm.cuh) re-loads the 128-pixel A tile once per tap: 9x the activation traffic out of L2, which
// (with the weight tile) saturates the L2->SM path at ~35% tensor-pipe utilisation.  Here a CTA owns a 16 x (8*SUB) pixel
// patch of one image.  Per 64-channel chunk ONE 4-D TMA box {64 ch, P px, 18 rows} lands the patch plus its 1-pixel halo in
// shared memory (rows = pixels, 128 B each, SWIZZLE_128B, row pitch P = patch width + 2 = 10 or 18 pixels; out-of-image pixels are zero-filled
// by TMA = the conv's padding).  The A operand of tap (ky,kx) for sub-tile s is then just a UMMA descriptor into that
// buffer: start row ky*P + kx + 8s, 8-row core matrices (8 consecutive x) with stride P*128 B between y rows.  The start is
// then no longer 1024-B aligned; measured on B200 (tests/test_gemm_gpu.py::test_conv_halo_base_offset_probe): the
// descriptor's base_offset must stay 0 — the 128-B swizzle is applied on absolute shared-memory address bits, which is
// exactly how TMA wrote the rows (setting base_offset = (addr>>7)&7 gives garbage).  The 9 taps x SUB sub-tiles of a chunk therefore cost one A load; every
// weight slab (one tap, 64 channels) is shared by the SUB sub-tiles.
// Optional 1x1 segments (the ResidualBlock skip conv over the raw, possibly concatenated, input) ride in the same accumulator.
#pragma once
#include "../../include/ddpm_b200.h"
#include "gemm_host.cuh"
#include "gn_epilogue.cuh"

namespace ddpm {

struct HaloSeg { int map; int taps; int kchunks; int c_base; };   // taps = 9 (haloed 3x3) or 1 (1x1, plain patch)

struct HaloParams {
    int H, W, NB;              // image geometry (H % 16 == 0, W % (8*SUB) == 0)
    int tiles_x, tiles_y;      // patches per image
    int n_tiles;               // Cout / BLOCK_N
    int nseg; HaloSeg seg[3];
    int N;                     // Cout
    void* out; int ldo;        // bf16 NHWC
    const float* bias; const float* rowvec; int rowvec_ld;   // rowvec: per-image vector [NB][rowvec_ld]
    const __nv_bfloat16* residual; int ldr;
    int desc_base_offset_mode; // 0 (default, correct): base_offset field = 0 ; 1: (start>>7)&7 (probe knob, wrong on B200)
    GnEpi gn;                  // GroupNorm statistics of the output (gn_epilogue.cuh); null = off
    const float* xfK; int xfC, xf_silu;   // CTA-pair kernel only (conv_halo2.cuh): GroupNorm(+SiLU) of the 3x3 input applied in shared memory;
                                          // xfK [NB][4][xfC] = {sc, sh, ..} per (image, input channel), null = off
};

template <int BLOCK_N, int SUB>
struct HaloCfg {
    static constexpr int P = 8 * SUB + 2;                       // halo row pitch in pixels = patch width + 2 (10 or 18)
    static constexpr int A_ROWS = 18 * P;
    static constexpr int A_BYTES = A_ROWS * 128;                // 22.5 KB / 40.5 KB actually filled by TMA
    static constexpr int A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;   // stage pitch (swizzle atoms need 1024-B aligned bases)
    // Weight stage = TPS taps of one 64-channel chunk.  Measured (DDPM_HALO_DBG experiments, profiles/README.md): every
    // producer -> issuer -> commit round trip costs ~300 cycles of the single issuing thread, whatever the stage holds; with
    // one tap per stage (256 MMA cycles at N=128) the tensor pipe starves at 65 %.  Three taps (one ky row) per stage = 768
    // MMA cycles per round trip.
    static constexpr int TPS = (BLOCK_N == 256) ? 1 : 3;
    static constexpr int TAP_BYTES = BLOCK_N * 128;
    static constexpr int B_BYTES = TAP_BYTES * TPS;
    static constexpr int NA = 2;
    // SUB = 2: every weight slab feeds two sub-tiles (256 pixels), i.e. half the weight traffic out of L2 per FLOP and 1536 MMA
    // cycles per stage; two 48 KB stages (3072 cycles in flight) then cover the TMA round trip.
    static constexpr int NB_ST = (TPS == 3) ? ((BLOCK_N == 128) ? ((SUB == 1) ? 3 : 2) : ((SUB == 1) ? 6 : 4)) : ((SUB == 1) ? 4 : 3);
    static constexpr int NACC = (2 * SUB * BLOCK_N <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (NACC * SUB * BLOCK_N <= 128) ? 128 : ((NACC * SUB * BLOCK_N <= 256) ? 256 : 512);
    static constexpr int OUT_STAGE_BYTES = 128 * 128;           // one 64-channel x 128-pixel output slab (TMA store source)
    static constexpr int TOTAL = NA * A_STRIDE + NB_ST * B_BYTES + 2 * OUT_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 1024 /*bias vector*/;
};
constexpr int HALO_EPI_WARPS = 8;                               // two per TMEM lane quarter (see umma_gemm.cuh)
constexpr int HALO_THREADS = 64 + 32 * HALO_EPI_WARPS;

// descriptor with explicit base offset (bits 49..51)
__device__ __forceinline__ uint64_t umma_smem_desc_bo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off) {
    return umma_smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)(base_off & 7) << 49);
}

template <int BLOCK_N, int SUB>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const HaloParams p) {
    pdl_trigger();
    using CF = HaloCfg<BLOCK_N, SUB>;
    constexpr int P = CF::P;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smem + CF::NA * CF::A_STRIDE;
    uint8_t* out_stage = smB + CF::NB_ST * CF::B_BYTES;            // [2][128 px][128 B]
    uint64_t* full_a = reinterpret_cast<uint64_t*>(out_stage + 2 * CF::OUT_STAGE_BYTES);
    uint64_t* empty_a = full_a + CF::NA;
    uint64_t* full_b = empty_a + CF::NA;
    uint64_t* empty_b = full_b + CF::NB_ST;
    uint64_t* tmem_full = empty_b + CF::NB_ST;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_vec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_a) + 512);     // [BLOCK_N] bias + per-image vector
    const uint32_t s_vec_u32 = smem_u32(s_vec);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int m_tiles = p.NB * tiles_per_img;
    const int total_tiles = m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB);
        for (int s = 0; s < CF::NA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < CF::NB_ST; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], HALO_EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, CF::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= TMA producer =======================
        // elect.sync from the converged warp (not `lane == 0`): the compiler then knows a single thread runs the
        // uniform-datapath instructions (UTMALDG / UTCHMMA) and emits them straight-line; under a plain lane test it
        // wraps every one of them in an ELECT/branch loop (~60-100 cycles per MMA, measured as a 2x loss at N=128).
        if (elect_one()) {
            int ia = 0, ib = 0;
            bool ok = true;
            for (int t = blockIdx.x; t < total_tiles && ok; t += gridDim.x) {
                const int m_tile = t % m_tiles, n_tile = t / m_tiles;
                const int n = m_tile / tiles_per_img, r = m_tile % tiles_per_img;
                const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8 * SUB;
                int kbase = 0;                       // running K coordinate of the packed weights
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const HaloSeg sg = p.seg[s];
                    const CUtensorMap* mA = sg.map == 0 ? &tmA0 : (sg.map == 1 ? &tmA1 : &tmA2);
                    const int Cseg = sg.kchunks * 64;
                    for (int kc = 0; kc < sg.kchunks && ok; ++kc) {
                        {   // activation patch (+halo for the 3x3 segment)
                            const int sa = ia % CF::NA; const uint32_t ph = (ia / CF::NA) & 1;
                            if (!mbar_wait(&empty_a[sa], ph ^ 1, 5)) { ok = false; break; }
                            const uint32_t bytes = (sg.taps == 9) ? (uint32_t)CF::A_BYTES : (uint32_t)(16 * 8 * SUB * 128);
                            mbar_expect_tx(&full_a[sa], bytes);
                            if (sg.taps == 9) tma_load_4d(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0 - 1, y0 - 1, n);
                            else              tma_load_4d(smA + sa * CF::A_STRIDE, mA, &full_a[sa], sg.c_base + kc * 64, x0, y0, n);
                            ++ia;
                        }
                        for (int tp = 0; tp < sg.taps; tp += CF::TPS) {
                            const int nt = (sg.taps - tp < CF::TPS) ? sg.taps - tp : CF::TPS;
                            const int sb = ib % CF::NB_ST; const uint32_t ph = (ib / CF::NB_ST) & 1;
                            if (!mbar_wait(&empty_b[sb], ph ^ 1, 6)) { ok = false; break; }
                            mbar_expect_tx(&full_b[sb], (uint32_t)(nt * CF::TAP_BYTES));
                            for (int j = 0; j < nt; ++j)
                                tma_load_3d(smB + sb * CF::B_BYTES + j * CF::TAP_BYTES, &tmB, &full_b[sb], kbase + (tp + j) * Cseg + kc * 64, n_tile * BLOCK_N, 0);
                            ++ib;
                        }
                    }
                    kbase += sg.taps * Cseg;
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc(128, BLOCK_N, 0, 0);
            int ia = 0, ib = 0, it = 0;
            bool ok = true;
            for (int t = blockIdx.x; t < total_tiles && ok; t += gridDim.x, ++it) {
                const int acc = (CF::NACC == 2) ? (it & 1) : 0;
                const uint32_t acc_ph = (CF::NACC == 2) ? ((it >> 1) & 1) : (it & 1);
                if (!mbar_wait(&tmem_empty[acc], acc_ph ^ 1, 4)) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SUB * BLOCK_N);
                bool first = true;
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const HaloSeg sg = p.seg[s];
                    for (int kc = 0; kc < sg.kchunks && ok; ++kc) {
                        const int sa = ia % CF::NA; const uint32_t pha = (ia / CF::NA) & 1;
                        if (!mbar_wait(&full_a[sa], pha, 7)) { ok = false; break; }
                        const uint32_t a_base = smem_u32(smA + sa * CF::A_STRIDE);
                        for (int tp0 = 0; tp0 < sg.taps; tp0 += CF::TPS) {
                            const int nt = (sg.taps - tp0 < CF::TPS) ? sg.taps - tp0 : CF::TPS;
                            const int sb = ib % CF::NB_ST; const uint32_t phb = (ib / CF::NB_ST) & 1;
                            if (!mbar_wait(&full_b[sb], phb, 2)) { ok = false; break; }
                            tc_fence_after();
                            // haloed patch: row (y+ky)*P + (x+kx+8s) ; plain 1x1 patch: row y*(8*SUB) + x + 8s
                            const int pitch = (sg.taps == 9) ? P : 8 * SUB;
#pragma unroll
                            for (int j = 0; j < CF::TPS; ++j) {
                                if (j >= nt) break;
                                const int tp = tp0 + j;
                                const uint32_t b_addr = smem_u32(smB + sb * CF::B_BYTES + j * CF::TAP_BYTES);
                                const int row0 = (sg.taps == 9) ? (tp / 3) * P + (tp % 3) : 0;
#pragma unroll
                                for (int sub = 0; sub < SUB; ++sub) {
                                    const uint32_t a_row = a_base + (uint32_t)(row0 + 8 * sub) * 128u;
                                    const uint32_t bo = p.desc_base_offset_mode ? ((a_row >> 7) & 7u) : 0u;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint64_t da = umma_smem_desc_bo(a_row + k * 32, 16, (uint32_t)pitch * 128u, bo);
                                        const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024);
                                        umma_bf16(d_tmem + (uint32_t)(sub * BLOCK_N), da, db, idesc, (first && k == 0) ? 0u : 1u);
                                    }
                                }
                                first = false;
                            }
                            umma_commit(&empty_b[sb]);
                            ++ib;
                        }
                        umma_commit(&empty_a[sa]);
                        ++ia;
                    }
                }
                if (ok) umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ======================= epilogue =======================
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;             // which 32-channel half of every 64-channel slab this warp converts
        const int r = q * 32 + lane;                 // accumulator row = pixel (y = r/8, x = r%8) of the sub-tile
        const int tid_epi = (int)threadIdx.x - 64;
        constexpr int EPI_T = 32 * HALO_EPI_WARPS;
        constexpr int NSLAB = BLOCK_N / 64;          // Cout % BLOCK_N == 0 (pick_block_n): every slab of every tile is full
        constexpr int NIT = SUB * NSLAB;
        int it = 0;
        uint32_t slab_ctr = 0;                       // staging-buffer parity, continues across tiles
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int m_tile = t % m_tiles, n_tile = t / m_tiles;
            const int n = m_tile / tiles_per_img, rr = m_tile % tiles_per_img;
            const int y0 = (rr / p.tiles_x) * 16, x0 = (rr % p.tiles_x) * 8 * SUB;
            const int acc = (CF::NACC == 2) ? (it & 1) : 0;
            const uint32_t acc_ph = (CF::NACC == 2) ? ((it >> 1) & 1) : (it & 1);
            if (!mbar_wait(&tmem_full[acc], acc_ph, 3)) break;
            tc_fence_after();
            // bias + per-image timestep vector of this tile, staged once (a patch lies inside one image)
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
#pragma unroll 1
            for (int i2 = 0; i2 < NIT; ++i2) {
                const int sub = i2 / NSLAB, s0 = (i2 % NSLAB) * 64;
                const long long pix = ((long long)n * p.H + y0 + (r >> 3)) * p.W + x0 + 8 * sub + (r & 7);
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * SUB * BLOCK_N + sub * BLOCK_N);
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
                // GroupNorm statistics of the output for its consumers (warp-uniform branch; gn_epilogue.cuh)
                if (p.gn.qstats) epi_quad_stats(f, true, p.gn.qstats + ((long long)n * (p.N >> 2) + (col >> 2)) * 2, lane);
                // registers -> 128B-swizzled staging slab (row r = pixel, 16-byte chunk j at physical chunk j ^ (r & 7))
                uint8_t* buf = out_stage + (slab_ctr & 1) * CF::OUT_STAGE_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    st_shared_v4(buf + (((grp * 4 + j) ^ (r & 7)) << 4), pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                // one barrier per slab: before it the issuing thread has waited until the previous store finished reading its
                // buffer (the one the next slab overwrites); after it all 128 pixel rows of this slab are staged
                fence_proxy_async_smem();
                if (tid_epi == 0) bulk_wait_group_read0();
                named_bar_sync(1, EPI_T);
                if (tid_epi == 0) {
                    tma_store_4d(&tmO, out_stage + (slab_ctr & 1) * CF::OUT_STAGE_BYTES, n_tile * BLOCK_N + s0, x0 + 8 * sub, y0, n);
                    bulk_commit_group();
                }
                ++slab_ctr;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (threadIdx.x == 64) bulk_wait_group0();   // outstanding TMA stores read this CTA's shared memory
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, CF::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct HaloLaunch { CUtensorMap a[3], b, o; HaloParams p; int block_n, sub; int tiles; double flops; int pair; /* CTA-pair kernel (conv_halo2.cuh) */ };

typedef ddpm_halo_desc HaloDesc;   // the C-ABI struct doubles as the internal description

inline bool halo_eligible(int H, int W, int Cout) { return H % 16 == 0 && W % 8 == 0 && Cout % 64 == 0 && H >= 16 && W >= 8; }

inline int build_halo(const HaloDesc& d, HaloLaunch& g) {
    memset(&g, 0, sizeof g);
    if (!halo_eligible(d.H, d.W, d.Cout)) return fail(-12, "halo conv: unsupported geometry %dx%d Cout=%d", d.H, d.W, d.Cout);
    g.block_n = pick_block_n(d.Cout);
    int sub = 1;   // measured on B200 (tools/dbg_dominant.py): one 16x8 sub-tile per CTA beats two (1089 vs 1012 TFLOP/s at N=128)
    if (d.force_sub == 1 || d.force_sub == 2) sub = d.force_sub;
    // CTA-pair kernel (cta_group::2): force_sub = 3 forces it, 1 / 2 force the single-CTA kernel, 0 = policy below
    static const int pair_policy = getenv("DDPM_HALO_PAIR") ? atoi(getenv("DDPM_HALO_PAIR")) : 1;
    const bool pair_ok = g.block_n >= 128 && (((long long)d.NB * (d.H / 16) * (d.W / 8)) % 2) == 0;
    g.pair = (d.force_sub == 3 || (d.force_sub == 0 && pair_policy)) && pair_ok ? 1 : 0;
    if (d.force_sub == 3 && !pair_ok) return fail(-12, "halo conv: the CTA-pair kernel needs Cout %% 128 == 0 and an even number of 16x8 tiles");
    if (sub == 2 && d.W % 16) return fail(-12, "halo conv: SUB=2 needs W %% 16 == 0");
    g.sub = sub;
    HaloParams& p = g.p;
    p.H = d.H; p.W = d.W; p.NB = d.NB; p.tiles_x = d.W / (8 * sub); p.tiles_y = d.H / 16; p.n_tiles = (d.Cout + g.block_n - 1) / g.block_n;
    p.N = d.Cout; p.out = d.out; p.ldo = d.Cout; p.bias = d.bias; p.rowvec = d.rowvec; p.rowvec_ld = d.rowvec_ld;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual); p.ldr = d.Cout; p.desc_base_offset_mode = d.base_offset_mode;
    p.gn = gn_epi_from_abi(d.gn);
    p.xfK = d.xf_K; p.xfC = d.a_C[d.seg_map[0]]; p.xf_silu = d.xf_silu;
    if (p.xfK && !(g.pair && d.seg_taps[0] == 9 && d.seg_map[0] == 0)) return fail(-12, "halo conv: the fused GroupNorm input transform needs the CTA-pair kernel and a 3x3 first segment on map 0");
    p.nseg = d.nseg;
    const int P = 8 * sub + 2;
    int rc;
    for (int s = 0; s < d.nseg; ++s) {
        p.seg[s] = HaloSeg{d.seg_map[s], d.seg_taps[s], d.seg_kchunks[s], d.seg_cbase[s]};
        const int m = d.seg_map[s];
        if (d.seg_taps[s] == 9) { if ((rc = make_tmap_4d(&g.a[m], d.a_ptr[m], d.a_C[m], d.W, d.H, d.NB, d.a_ld[m], 64, P, 18, 1))) return rc; }
        else                    { if ((rc = make_tmap_4d(&g.a[m], d.a_ptr[m], d.a_C[m], d.W, d.H, d.NB, d.a_ld[m], 64, 8 * sub, 16, 1))) return rc; }
    }
    for (int i = 0; i < 3; ++i) { bool used = false; for (int s = 0; s < d.nseg; ++s) used |= d.seg_map[s] == i; if (!used) g.a[i] = g.a[d.seg_map[0]]; }
    if ((rc = make_tmap_3d(&g.b, d.w, d.Ktot, d.Cout, 1, d.ldw, 0, 64, g.pair ? g.block_n / 2 : g.block_n))) return rc;   // pair: each CTA loads half of the weight rows
    if ((rc = make_tmap_4d(&g.o, d.out, d.Cout, d.W, d.H, d.NB, d.Cout, 64, 8, 16, 1))) return rc;      // output slab = 64 ch x (8 x 16) px
    g.tiles = d.NB * p.tiles_x * p.tiles_y * p.n_tiles;
    g.flops = 2.0 * d.NB * d.H * d.W * (double)d.Cout * d.Ktot;
    return 0;
}

template <int BLOCK_N, int SUB>
inline int launch_halo_inst(const HaloLaunch& g, cudaStream_t st) {
    using CF = HaloCfg<BLOCK_N, SUB>;
    auto kern = conv3x3_halo_kernel<BLOCK_N, SUB>;
    static_assert(CF::TOTAL <= 232448, "halo conv: shared memory budget (227 KB) exceeded");
    static bool attr_done = false;
    if (!attr_done) { DDPM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::TOTAL)); attr_done = true; }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    const int ctas = g.tiles < num_sms ? g.tiles : num_sms;
    launch_k(kern, ctas, HALO_THREADS, CF::TOTAL, st, g.a[0], g.a[1], g.a[2], g.b, g.o, g.p);
    DDPM_CUDA_OK(cudaGetLastError());
    return 0;
}
inline int launch_halo2(const HaloLaunch& g, cudaStream_t st);      // conv_halo2.cuh
inline int launch_halo(const HaloLaunch& g, cudaStream_t st) {
    if (g.pair) return launch_halo2(g, st);
    if (g.block_n == 64 && g.sub == 1) return launch_halo_inst<64, 1>(g, st);
    if (g.block_n == 64 && g.sub == 2) return launch_halo_inst<64, 2>(g, st);
    if (g.block_n == 128 && g.sub == 1) return launch_halo_inst<128, 1>(g, st);
    if (g.block_n == 128 && g.sub == 2) return launch_halo_inst<128, 2>(g, st);
    if (g.block_n == 256 && g.sub == 1) return launch_halo_inst<256, 1>(g, st);
    if (g.block_n == 256 && g.sub == 2) return launch_halo_inst<256, 2>(g, st);
    return fail(-6, "unsupported halo conv variant");
}

}  // namespace ddpm
