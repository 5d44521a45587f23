// ddpm_gemm_desc (C ABI) -> GemmLaunch (tensor maps + kernel params)
#pragma once
#include "../../include/ddpm_b200.h"
#include "gemm_host.cuh"

namespace ddpm {

// standard segment: 1 tap (1x1) or 9 taps (3x3, stride 1, pad 1)
inline GemmSeg make_seg(int map, int taps, int kchunks, int c_base) {
    GemmSeg g; memset(&g, 0, sizeof g);
    g.map = map; g.taps = taps; g.kchunks = kchunks; g.c_base = c_base; g.cmul = 1;
    if (taps == 9) for (int t = 0; t < 9; ++t) { g.dx[t] = (signed char)(t % 3 - 1); g.dy[t] = (signed char)(t / 3 - 1); }
    return g;
}

inline int build_gemm(const ddpm_gemm_desc& d, GemmLaunch& g) {
    memset(&g, 0, sizeof g);
    g.mode = d.mode;
    g.block_n = d.block_n ? d.block_n : pick_block_n(d.N);
    if (g.block_n != 64 && g.block_n != 128 && g.block_n != 256) return fail(-10, "block_n must be 64/128/256");
    if (d.N % 32) return fail(-10, "N=%d must be a multiple of 32", d.N);
    GemmParams& p = g.p;
    p.M = d.M; p.N = d.N; p.W = d.W; p.H = d.H;
    p.out = d.out; p.ldo = d.ldo; p.out_z_stride = d.out_z_stride; p.out_tap_stride = d.out_tap_stride; p.flags = d.flags;
    p.bias = d.bias; p.rowvec = d.rowvec; p.rowvec_ld = d.rowvec_ld; p.rows_per_vec = d.rows_per_vec > 0 ? d.rows_per_vec : 1;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual); p.ldr = d.ldr; p.alpha = d.alpha;
    p.b_k_base = d.b_k_base; p.a_z_n = d.a_z_n; p.b_z = d.b_z;
    p.taps = d.taps > 0 ? d.taps : 1; p.splits = d.splits > 0 ? d.splits : 1; p.kblocks = d.kblocks;
    p.a_c_base = d.a_c_base; p.b_c_base = d.b_c_base;
    p.b_cmul = d.b_estride > 1 ? d.b_estride : 1; p.b_pad = d.b_estride > 1 ? d.b_pad : 1;
    p.o_mul = d.o_mul > 1 ? d.o_mul : 1; p.o_py = d.o_py; p.o_px = d.o_px; p.oW = d.W * p.o_mul; p.oH = d.H * p.o_mul;
    const int aes = d.a_estride > 1 ? d.a_estride : 1, bes = d.b_estride > 1 ? d.b_estride : 1;
    const int gz = d.grid_z > 0 ? d.grid_z : 1;
    const int n_tiles = (d.N + g.block_n - 1) / g.block_n;
    const int m_tiles = (d.M + 127) / 128;
    { static const int dbg = getenv("DDPM_GEMM_DBG") ? atoi(getenv("DDPM_GEMM_DBG")) : 0; p.dbg = dbg; }
    p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.grid_z = gz; p.kk_splits = d.kk_splits > 1 ? d.kk_splits : 1;
    if (p.kk_splits > 1 && (d.mode != GEMM_KK || gz != p.kk_splits || !(d.flags & EPI_ATOMIC))) return fail(-10, "kk_splits needs mode 0, grid_z == kk_splits and the atomic fp32 epilogue");
    g.grid = dim3(m_tiles * n_tiles * gz, 1, 1);   // clipped to the SM count at launch (persistent CTAs)
    int rc;
    if (d.mode == GEMM_KK) {
        if (!pick_box(d.W, d.H, 128, p.w_t, p.h_t, p.n_t)) return fail(-11, "KK: unsupported geometry W=%d H=%d", d.W, d.H);
        p.nseg = d.nseg;
        if (d.nseg < 1 || d.nseg > 3) return fail(-11, "KK: nseg must be 1..3");
        int slabs = 0;
        for (int s = 0; s < d.nseg; ++s) {
            if (d.seg_custom[s]) {
                if (d.seg_taps[s] < 1 || d.seg_taps[s] > 9) return fail(-11, "KK: custom taps must be 1..9");
                p.seg[s] = make_seg(d.seg_map[s], 1, d.seg_kchunks[s], d.seg_cbase[s]);
                p.seg[s].taps = d.seg_taps[s]; p.seg[s].cmul = d.seg_cmul[s] > 0 ? d.seg_cmul[s] : 1;
                for (int t = 0; t < d.seg_taps[s]; ++t) { p.seg[s].dx[t] = d.seg_dx[s][t]; p.seg[s].dy[t] = d.seg_dy[s][t]; }
            } else {
                if (d.seg_taps[s] != 1 && d.seg_taps[s] != 9) return fail(-11, "KK: taps must be 1 or 9");
                p.seg[s] = make_seg(d.seg_map[s], d.seg_taps[s], d.seg_kchunks[s], d.seg_cbase[s]);
            }
            slabs += d.seg_taps[s] * d.seg_kchunks[s];
        }
        for (int i = 0; i < 3; ++i) {
            if (!d.a_ptr[i]) { g.a[i] = g.a[0]; continue; }
            if ((rc = make_tmap_4d(&g.a[i], d.a_ptr[i], d.a_C[i], d.W * aes, d.H * aes, d.NB, d.a_ld[i], 64, p.w_t, p.h_t, p.n_t, aes))) return rc;
        }
        // CTA-pair kernel (cta_group::2): two consecutive m_tiles share one M = 256 MMA, each CTA loads half of the weight rows
        // Policy (cta_pair = 0), measured on B200 (profiles/r02_pair_ab.txt): in the single-stream sampler the pair kernel wins
        // (4.74 -> 4.65 ms per step); in the training step, where the weight-gradient GEMMs of the side stream run next to the
        // main chain, 2-CTA clusters of 227 KB CTAs schedule worse than independent CTAs (9.47 -> 9.68 ms) - so pairs are used by
        // inference plans only (gemm_pair_hint(), set by UnetEngine::plan).  DDPM_GEMM_PAIR overrides (bit 0 KK, bit 1 MN-major).
        static const int pair_env = getenv("DDPM_GEMM_PAIR") ? atoi(getenv("DDPM_GEMM_PAIR")) : -1;
        const int pair_policy = pair_env >= 0 ? pair_env : (gemm_pair_hint() ? 1 : 0);
        const bool pair_ok = g.block_n >= 128 && (m_tiles % 2) == 0;
        g.pair = (d.cta_pair >= 2 || (d.cta_pair == 0 && (pair_policy & 1))) && pair_ok ? 1 : 0;
        if (d.cta_pair == 2 && !pair_ok) return fail(-11, "KK: the CTA-pair kernel needs N %% 128 == 0 and an even number of 128-row tiles");
        if ((rc = make_tmap_3d(&g.b, d.b_ptr, d.b_K, d.b_rows, d.b_batch > 0 ? d.b_batch : 1, d.b_ld, d.b_bs, 64, g.pair ? g.block_n / 2 : g.block_n))) return rc;
        g.flops = 2.0 * d.M * d.N * 64.0 * slabs * gz;
    } else if (d.mode == GEMM_MNMN) {
        if (!pick_box(d.W, d.H, 64, p.wk_t, p.hk_t, p.nk_t)) return fail(-11, "MNMN: unsupported geometry W=%d H=%d", d.W, d.H);
        if (d.M % 64) return fail(-11, "MNMN: M must be a multiple of 64");
        {   // CTA pairs along M (two 128-row tiles of output channels per M = 256 MMA), each CTA holding half of the N columns
            // measured slower inside the training step (9.47 -> 9.67 ms, see the K-major policy below): off unless asked for
            static const int pair_env = getenv("DDPM_GEMM_PAIR") ? atoi(getenv("DDPM_GEMM_PAIR")) : 0;
            const bool pair_ok = g.block_n >= 128 && (m_tiles % 2) == 0 && (d.M % 128) == 0;
            g.pair = (d.cta_pair >= 2 || (d.cta_pair == 0 && (pair_env & 2))) && pair_ok ? 1 : 0;
            if (d.cta_pair == 2 && !pair_ok) return fail(-11, "MNMN: the CTA-pair kernel needs M %% 256 == 0 and N %% 128 == 0");
        }
        if ((rc = make_tmap_4d(&g.a[0], d.a_ptr[0], d.a_C[0], d.W, d.H, d.NB, d.a_ld[0], 64, p.wk_t, p.hk_t, p.nk_t))) return rc;
        g.a[1] = g.a[2] = g.a[0];
        if ((rc = make_tmap_4d(&g.b, d.b_ptr, d.b_K, d.W * bes, d.H * bes, d.NB, d.b_ld, 64, p.wk_t, p.hk_t, p.nk_t, bes))) return rc;
        g.flops = 2.0 * d.M * d.N * 64.0 * d.kblocks * p.taps * (gz / (p.taps * p.splits));
    } else if (d.mode == GEMM_KMN) {
        if (!pick_box(d.W, d.H, 128, p.w_t, p.h_t, p.n_t)) return fail(-11, "KMN: unsupported geometry W=%d H=%d", d.W, d.H);
        if (d.H != 1 || d.W % 64) return fail(-11, "KMN: B operand needs a (C, tokens%%64==0, 1, batch) tensor");
        if ((rc = make_tmap_4d(&g.a[0], d.a_ptr[0], d.a_C[0], d.W, d.H, d.NB, d.a_ld[0], 64, p.w_t, p.h_t, p.n_t))) return rc;
        g.a[1] = g.a[2] = g.a[0];
        if ((rc = make_tmap_4d(&g.b, d.b_ptr, d.b_K, d.W, d.H, d.NB, d.b_ld, 64, 64, 1, 1))) return rc;
        g.flops = 2.0 * d.M * d.N * 64.0 * d.kblocks * gz;
    } else {
        return fail(-10, "unknown gemm mode %d", d.mode);
    }
    p.gn = gn_epi_from_abi(d.gn); p.gn_hw = d.W * d.H;
    if (p.gn.qstats && (d.mode != GEMM_KK || p.kk_splits > 1 || gz != 1 || (d.flags & (EPI_OUT_F32 | EPI_ATOMIC)) || (p.gn_hw % 32) || (d.M % 32)))
        return fail(-10, "GroupNorm statistics fusion needs a plain bf16 KK conv over NHWC pixels with H*W %% 32 == 0");
    // TMA-store epilogue: plain bf16 outputs with identity row mapping (everything but fp32 / atomic / scattered-row outputs)
    g.o = g.b; p.tma_store = 0;
    static const bool no_tma_store = getenv("DDPM_NO_TMA_STORE") != nullptr;
    const bool plain_rows = (d.mode != GEMM_KK || p.o_mul == 1) && d.mode != GEMM_MNMN;   // MN-major mode keeps 4 epilogue warps + direct stores
    if (!no_tma_store && !(d.flags & (EPI_OUT_F32 | EPI_ATOMIC)) && plain_rows && d.out && (d.ldo % 8) == 0 &&
        (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && (d.out_z_stride % 8) == 0) {
        if ((rc = make_tmap_3d(&g.o, d.out, d.N, d.M, gz, d.ldo, d.out_z_stride, 64, 128))) return rc;
        p.tma_store = 1;
    }
    return 0;
}

}  // namespace ddpm
