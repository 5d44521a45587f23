// UNet plan builder + executor.  The network of ddpm_torch/models/unet.py:92-233 is compiled once per
// (batch, resolution) into flat lists of kernel launches (pack / forward / backward) over a caller-owned workspace.
// The backward list is produced tape-style: every forward block registers a closure that appends its adjoint ops.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "gemm_build.cuh"
#include "conv_halo.cuh"
#include "conv_halo2.cuh"
#include "attn_fused.cuh"
#include "kernels_simt.cuh"

namespace ddpm {

struct ParamInfo { std::string name; int nd; int dims[4]; long long numel; long long off; };
struct T4 { long long off = -1; int B = 0, H = 0, W = 0, C = 0;
            long long qs = -1;    // workspace offset of the producer-written quad statistics [B][C/4][2] fp64 (-1: none)
            long long pix() const { return (long long)B * H * W; } long long numel() const { return pix() * C; } };
struct Src { T4 t0, t1; bool two = false; int C() const { return t0.C + (two ? t1.C : 0); } };
enum { OP_TEMB = 1 };
struct Op { std::string name; double flops; std::function<int(cudaStream_t)> run; int launches = 1; bool side = false; int tag = 0;
            int wait_pack = 0;   // forward list: 1 = needs the packed timestep-MLP weights, 2 = first consumer of the packed conv weights
            int chunk = -1;   // >= 0: gradient-chunk boundary of the backward list (see UnetEngine::chunks)
};
// A contiguous range [lo, hi) of the flat gradient buffer that is FINAL once `ev` has fired (recorded by the backward pass on the
// side stream, after the last weight-gradient GEMM / gradient unpack / main-chain reduction that writes into it).  Chunks are
// listed in the order the backward completes them, so a data-parallel caller can start the all-reduce of a chunk while the
// rest of the backward is still running (utils/train.py:149-153 does this with DDP's buckets).
struct GradChunk { long long lo = 0, hi = 0; int unpack_lo = 0, unpack_hi = 0; cudaEvent_t ev = nullptr; };
struct GnSaved { Src in; float* K; const float* gamma; const float* beta; float* dgamma; float* dbeta; int silu; float drop_p; uint32_t layer; unsigned char* mask; };

static inline int grid_for(long long n, int threads = 256) { long long g = (n + threads - 1) / threads; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1; return (int)g; }
static inline int oct_threads(int C) { const int oct = C / 8; return oct <= 256 ? (256 / oct) * oct : 0; }

struct UnetEngine {
    ddpm_unet_cfg cfg;
    std::vector<ParamInfo> params;
    std::map<std::string, int> pidx;
    long long flat_elems = 0;
    // bound buffers
    float* P = nullptr; float* G = nullptr;
    uint8_t* ws = nullptr; size_t ws_bytes = 0;
    // plan state
    int B = 0, H = 0, W = 0; bool train = false; bool dry = true; size_t cursor = 0; bool planned = false;
    std::vector<Op> pack_ops, fwd_ops, bwd_ops;
    size_t first_packed_op = 0;
    std::vector<std::function<void()>> tape;
    std::vector<std::string> tape_tag; std::string cur_tag = "late";       // gradient-chunk group of every tape entry
    void tape_push(std::function<void()> f) { tape.push_back(std::move(f)); tape_tag.push_back(cur_tag); }
    std::vector<GradChunk> chunks;
    std::map<long long, std::pair<T4, bool>> grads;     // fwd tensor offset -> (grad tensor, written?)
    size_t zero_fwd_off = 0, zero_fwd_bytes = 0, zero_bwd_off = 0, zero_bwd_bytes = 0;   // per-pass zeroed regions
    size_t once_zero_off = 0, once_zero_bytes = 0;                                       // zeroed at plan time only
    std::vector<SgemmParams> tp_table_host; size_t tp_table_off = 0; int tp_max_c = 0;
    std::vector<SgemmParams> tp_uni_table_host; size_t tp_uni_table_off = 0;
    std::vector<Op> temb_uni_ops;          // one-row timestep path + broadcast (sampler: the whole batch shares t)
    std::vector<FcEnt> fc_table_host; size_t fc_table_off = 0;   // tensor-core timestep projections (concatenated fc weights)
    bool uniform_t = false;                // set by ddpm_sampler_step around its forward
    std::vector<SgemmParams> tpw_table_host, tpd_table_host; size_t tpw_table_off = 0, tpd_table_off = 0;
    uint32_t layer_counter = 0;
    double fwd_flops = 0, bwd_flops = 0;
    int n_tc_gemms = 0, n_generic = 0;
    // per-call IO (read by ops at run time)
    const float* x_in = nullptr; const long long* t_in = nullptr; float* eps_out = nullptr;
    unsigned long long drop_seed = 0;
    // diffusion-side internal buffers
    size_t xt_off = 0, eps_off = 0, tbuf_off = 0, coefcur_off = 0, counter_off = 0, deps_off = 0;

    // ------------------------------------------------------------------ parameters (reference registration order)
    void add_param(const std::string& n, std::initializer_list<int> d) {
        ParamInfo p; p.name = n; p.nd = (int)d.size(); p.numel = 1; int i = 0;
        for (int v : d) { p.dims[i++] = v; p.numel *= v; }
        for (; i < 4; ++i) p.dims[i] = 1;
        p.off = flat_elems; flat_elems += (p.numel + 63) / 64 * 64;
        pidx[n] = (int)params.size(); params.push_back(p);
    }
    void reg_lin(const std::string& p, int i, int o) { add_param(p + ".weight", {o, i}); add_param(p + ".bias", {o}); }
    void reg_conv(const std::string& p, int i, int o, int k) { add_param(p + ".weight", {o, i, k, k}); add_param(p + ".bias", {o}); }
    void reg_gn(const std::string& p, int n) { add_param(p + ".weight", {n}); add_param(p + ".bias", {n}); }
    void reg_res(const std::string& p, int i, int o) {
        reg_gn(p + ".norm1", i); reg_conv(p + ".conv1", i, o, 3); reg_lin(p + ".fc", cfg.temb_dim, o);
        reg_gn(p + ".norm2", o); reg_conv(p + ".conv2", o, o, 3);
        if (i != o) reg_conv(p + ".skip", i, o, 1);
    }
    void reg_attn(const std::string& p, int n) { reg_gn(p + ".norm", n); reg_conv(p + ".project_in", n, 3 * n, 1); reg_conv(p + ".project_out", n, n, 1); }
    void reg_block(const std::string& p, int i, int o, bool at) { if (at) { reg_res(p + ".0", i, o); reg_attn(p + ".1", o); } else reg_res(p, i, o); }
    int chs(int l) const { return cfg.hid_channels * cfg.ch_mult[l]; }
    void register_params() {
        const int ch = cfg.hid_channels, L = cfg.levels, nrb = cfg.num_res_blocks, E = cfg.temb_dim;
        reg_lin("embed.0", ch, E); reg_lin("embed.2", E, E);
        reg_conv("in_conv", cfg.in_channels, ch, 3);
        for (int i = 0; i < L; ++i) {
            const int prev = i ? chs(i - 1) : ch, cur = chs(i);
            const std::string p = "downsamples.level_" + std::to_string(i);
            reg_block(p + ".0", prev, cur, cfg.attn[i]);
            for (int j = 1; j < nrb; ++j) reg_block(p + "." + std::to_string(j), cur, cur, cfg.attn[i]);
            if (i != L - 1) reg_conv(p + "." + std::to_string(nrb) + ".1", cur, cur, 3);
        }
        const int mid = chs(L - 1);
        reg_res("middle.0", mid, mid); reg_attn("middle.1", mid); reg_res("middle.2", mid, mid);
        for (int i = 0; i < L; ++i) {
            const int nxt = i == 0 ? ch : chs(i - 1), prev = i == L - 1 ? chs(L - 1) : chs(i + 1), cur = chs(i);
            const std::string p = "upsamples.level_" + std::to_string(i);
            reg_block(p + ".0", prev + cur, cur, cfg.attn[i]);
            for (int j = 1; j < nrb; ++j) reg_block(p + "." + std::to_string(j), 2 * cur, cur, cfg.attn[i]);
            reg_block(p + "." + std::to_string(nrb), nxt + cur, cur, cfg.attn[i]);
            if (i != 0) reg_conv(p + "." + std::to_string(nrb + 1) + ".1", cur, cur, 3);
        }
        reg_gn("out_conv.0", ch); reg_conv("out_conv.2", ch, cfg.out_channels, 3);
        layout_params();
    }
    // Memory layout of the flat buffers: registration order, EXCEPT that the tensors whose gradients only exist at the very end
    // of the backward pass (the embedding MLP and the per-block timestep projections fc.weight, unet.py:122-126,77) are moved
    // behind everything else ("late region"), so that the rest forms contiguous per-level ranges that complete one after the
    // other.  The parameter LIST (state_dict order) is unchanged - only the offsets differ.
    static bool is_late(const std::string& n) { return n.rfind("embed.", 0) == 0 || (n.size() > 10 && n.compare(n.size() - 10, 10, ".fc.weight") == 0); }
    long long main_elems = 0;
    void layout_params() {
        long long off = 0;
        for (auto& p : params) if (!is_late(p.name)) { p.off = off; off += (p.numel + 63) / 64 * 64; }
        main_elems = off;
        for (auto& p : params) if (is_late(p.name)) { p.off = off; off += (p.numel + 63) / 64 * 64; }
        flat_elems = off;
    }
    // group of a parameter / tape entry for the gradient chunks: d<i> (in_conv + down level i), mid, u<i> (+ out_conv), late
    std::string group_of(const std::string& n) const {
        if (is_late(n)) return "late";
        if (n.rfind("in_conv", 0) == 0) return "d0";
        if (n.rfind("out_conv", 0) == 0) return "u" + std::to_string(cfg.levels - 1);
        if (n.rfind("middle", 0) == 0) return "mid";
        const size_t k = n.find("level_");
        if (k != std::string::npos) return std::string(n[0] == 'd' ? "d" : "u") + std::to_string(atoi(n.c_str() + k + 6));
        return "late";
    }
    const ParamInfo& pinfo(const std::string& n) const { return params[pidx.at(n)]; }
    float* PP(const std::string& n) const { return P + pinfo(n).off; }
    float* GP(const std::string& n) const { return G ? G + pinfo(n).off : nullptr; }

    // ------------------------------------------------------------------ workspace
    size_t alloc(size_t bytes) { const size_t o = (cursor + 1023) & ~size_t(1023); cursor = o + bytes; return o; }
    template <class T> T* at(size_t off) const { return reinterpret_cast<T*>(ws + off); }
    T4 newT(int b, int h, int w, int c) { T4 t; t.B = b; t.H = h; t.W = w; t.C = c; t.off = (long long)alloc((size_t)t.numel() * 2); return t; }
    bf16* bp(const T4& t) const { return at<bf16>((size_t)t.off); }
    GnSrc gsrc(const Src& s) const { GnSrc g; g.x0 = bp(s.t0); g.C0 = s.t0.C; g.x1 = s.two ? bp(s.t1) : nullptr; g.C1 = s.two ? s.t1.C : 0; return g; }
    static Src one(const T4& t) { Src s; s.t0 = t; s.two = false; return s; }

    void push(std::vector<Op>& L, const std::string& name, double flops, std::function<int(cudaStream_t)> f, int launches = 1, bool side = false) {
        Op o; o.name = name; o.flops = flops; o.run = std::move(f); o.launches = launches; o.side = side; L.push_back(std::move(o));
    }
    static int count_launches(const std::vector<Op>& L) { int n = 0; for (auto& o : L) n += o.launches; return n; }
    // grad tensor of a forward tensor; `first` tells the producer whether to overwrite (=) or accumulate (+=)
    T4 grad_of(const T4& t, bool* first) {
        auto it = grads.find(t.off);
        if (it == grads.end()) { T4 g = newT(t.B, t.H, t.W, t.C); it = grads.emplace(t.off, std::make_pair(g, false)).first; }
        if (first) { *first = !it->second.second; it->second.second = true; }
        return it->second.first;
    }
    bool has_grad(const T4& t) const { auto it = grads.find(t.off); return it != grads.end() && it->second.second; }

    // ------------------------------------------------------------------ op builders
    bool tc_ok_geom(int h, int w) const { int a, b, c; return pick_box(w, h, 128, a, b, c) && pick_box(w, h, 64, a, b, c); }

    // conv: out = conv(in) [+ 1x1 skip(skip_in)] + bias + rowvec[b] + residual ; packed weights [Co][ldw]
    struct ConvSpec {
        std::string name; Src in; int ksize = 3, stride = 1, map = MAP_NORMAL;
        const bf16* wp = nullptr; long long ldw = 0; bool has_skip = false; Src skip_in;
        bool skip_identity = false;        // the skip chunks carry identity weights (the block's residual): no FLOPs of the reference's count
        const float* bias = nullptr; const float* rowvec = nullptr; int rowvec_ld = 0; const bf16* residual = nullptr;
        T4 out; float* out_nchw = nullptr; int Co = 0; int Ho = 0, Wo = 0; bool accumulate = false;
        bool want_qstats = false;          // the output feeds a GroupNorm -> statistics in the epilogue when the kernel can
        const float* xfK = nullptr; int xf_silu = 0;   // inference: `in` is the RAW GroupNorm input, the conv applies act(sc*x + sh) in its
                                                       // operand path (CTA-pair haloed kernel only; see xf_ok / gn_prep_xf)
    };
    // GroupNorm(+SiLU) folded into the consumer 3x3 conv's operand path (conv_halo2.cuh XF): inference plans, single-source input with
    // producer statistics, and a conv that takes the CTA-pair haloed kernel.
    // OPT-IN (DDPM_XF=1, read at plan time): measured on B200 (profiles/r02_halo_xf_experiment.txt) the rewrite of the landed tile
    // costs the conv +46 % in isolation (+17 % for the extra barrier hop alone, the rest for the rewrite competing for the shared-memory
    // bandwidth the MMAs already saturate) against the ~40 % of a conv that the removed apply pass takes - break-even at best, and
    // the whole sampler step got slower (4.68 -> 5.09 ms at bs=256).  Kept, with unit and network parity tests, for hardware /
    // configurations where the balance differs.
    bool xf_ok(const Src& x, int Co, int h, int w) const {
        const char* on = getenv("DDPM_XF");
        const bool off = !on || atoi(on) == 0 || getenv("DDPM_NO_GN_EPI") != nullptr || getenv("DDPM_NO_HALO") != nullptr ||
                         (getenv("DDPM_HALO_PAIR") && atoi(getenv("DDPM_HALO_PAIR")) == 0);
        if (off || train || x.two || x.t0.qs < 0 || x.t0.C % 128 || Co % 128) return false;
        if (!halo_eligible(h, w, Co) || !tc_ok_geom(h, w)) return false;
        return (((long long)x.t0.B * (h / 16) * (w / 8)) % 2) == 0;
    }
    float* gn_prep_xf(std::vector<Op>& L, const std::string& name, const Src& in, const std::string& pname) {
        const int C = in.C(), Bn = in.t0.B, HW = in.t0.H * in.t0.W;
        float* K = at<float>(alloc((size_t)Bn * 4 * C * 4));
        GnApply a; memset(&a, 0, sizeof a);
        a.s = gsrc(in); a.HW = HW; a.qs0 = at<double>((size_t)in.t0.qs); a.qs1 = nullptr;
        a.gamma = PP(pname + ".weight"); a.beta = PP(pname + ".bias"); a.eps = 1e-6f; a.Kout = K;
        const int thr = C < 256 ? C : 256;
        push(L, name + ".prep", 0, [a, Bn, thr](cudaStream_t st) { launch_k(k_gn_prep, Bn, thr, 0, st, a); return (int)cudaGetLastError(); });
        return K;
    }
    // epilogue fusion of a conv: fills `gn` and returns the workspace offset of the quad statistics (-1: none)
    long long conv_gn_epi(const ConvSpec& c, ddpm_gn_epi& gn) {
        memset(&gn, 0, sizeof gn);
        long long qs = -1;
        static const bool no_fuse = getenv("DDPM_NO_GN_EPI") != nullptr;
        if (no_fuse) return qs;
        const int Bn = c.in.t0.B, HW = c.Ho * c.Wo;
        if (c.want_qstats && c.Co % 32 == 0 && HW % 32 == 0) {
            qs = (long long)zero_fwd((size_t)Bn * (c.Co / 4) * 2 * 8);
            gn.qstats = at<double>((size_t)qs);
        }
        return qs;
    }
    long long conv_op(std::vector<Op>& L, const ConvSpec& c, double* flops_acc) {
        const int Cin = c.in.C(), taps = c.ksize * c.ksize;
        const T4& i0 = c.in.t0;
        const int Bn = i0.B;
        const long long Pout = (long long)Bn * c.Ho * c.Wo;
        double fl = 2.0 * Pout * c.Co * (double)(taps * Cin + ((c.has_skip && !c.skip_identity) ? c.skip_in.C() : 0));
        if (flops_acc) *flops_acc += fl;
        const bool s2 = c.stride == 2 && c.map == MAP_NORMAL && c.ksize == 3 && !c.has_skip && c.Ho * 2 == i0.H && c.Wo * 2 == i0.W;
        bool tc = (c.stride == 1 || s2) && c.map == MAP_NORMAL && !c.out_nchw && !c.in.two && (i0.C % 64 == 0) && (c.Co % 64 == 0) &&
                  (s2 || (c.Ho == i0.H && c.Wo == i0.W)) && tc_ok_geom(c.Ho, c.Wo) && !(c.accumulate && c.residual);
        if (c.has_skip) tc = tc && (c.skip_in.t0.C % 64 == 0) && (!c.skip_in.two || c.skip_in.t1.C % 64 == 0);
        static const bool no_halo = getenv("DDPM_NO_HALO") != nullptr;
        if (tc && !no_halo && c.ksize == 3 && c.stride == 1 && halo_eligible(c.Ho, c.Wo, c.Co)) {
            ddpm_halo_desc h; memset(&h, 0, sizeof h);
            h.NB = Bn; h.H = c.Ho; h.W = c.Wo; h.Cout = c.Co;
            h.a_ptr[0] = bp(i0); h.a_C[0] = i0.C; h.a_ld[0] = i0.C;
            h.nseg = 1; h.seg_map[0] = 0; h.seg_taps[0] = 9; h.seg_kchunks[0] = i0.C / 64; h.seg_cbase[0] = 0;
            long long K = 9LL * Cin;
            if (c.has_skip) {
                h.a_ptr[1] = bp(c.skip_in.t0); h.a_C[1] = c.skip_in.t0.C; h.a_ld[1] = c.skip_in.t0.C;
                h.seg_map[1] = 1; h.seg_taps[1] = 1; h.seg_kchunks[1] = c.skip_in.t0.C / 64; h.seg_cbase[1] = 0; h.nseg = 2; K += c.skip_in.t0.C;
                if (c.skip_in.two) {
                    h.a_ptr[2] = bp(c.skip_in.t1); h.a_C[2] = c.skip_in.t1.C; h.a_ld[2] = c.skip_in.t1.C;
                    h.seg_map[2] = 2; h.seg_taps[2] = 1; h.seg_kchunks[2] = c.skip_in.t1.C / 64; h.seg_cbase[2] = 0; h.nseg = 3; K += c.skip_in.t1.C;
                }
            }
            h.w = c.wp; h.ldw = c.ldw; h.Ktot = (int)K; h.out = bp(c.out); h.bias = c.bias; h.rowvec = c.rowvec; h.rowvec_ld = c.rowvec_ld;
            h.residual = c.accumulate ? (const void*)bp(c.out) : (const void*)c.residual; h.base_offset_mode = 0;
            const long long qs = conv_gn_epi(c, h.gn);
            if (c.xfK) { h.xf_K = c.xfK; h.xf_silu = c.xf_silu; h.force_sub = 3; }
            ++n_tc_gemms;
            if (dry) { push(L, c.name, fl, [](cudaStream_t) { return 0; }); return qs; }
            HaloLaunch g; int rc = build_halo(h, g);
            if (rc) { plan_error = rc; return qs; }
            push(L, c.name + "[halo]", fl, [g](cudaStream_t st) { return launch_halo(g, st); });
            return qs;
        }
        if (c.xfK) { plan_error = fail(-31, "internal: conv '%s' was planned with a fused GroupNorm input but did not take the CTA-pair haloed kernel", c.name.c_str()); return -1; }
        if (tc) {
            ddpm_gemm_desc d; memset(&d, 0, sizeof d);
            d.mode = GEMM_KK; d.M = (int)Pout; d.N = c.Co; d.W = c.Wo; d.H = c.Ho; d.NB = Bn;
            d.a_ptr[0] = bp(i0); d.a_C[0] = i0.C; d.a_ld[0] = i0.C;
            d.nseg = 1; d.seg_map[0] = 0; d.seg_taps[0] = taps; d.seg_kchunks[0] = i0.C / 64; d.seg_cbase[0] = 0;
            if (s2) {   // SamePad2d(3,2): pad bottom/right only -> tap offsets 0..2 on the stride-2 sampled map, OOB = zero fill
                d.a_estride = 2; d.seg_custom[0] = 1; d.seg_cmul[0] = 2;
                for (int t = 0; t < 9; ++t) { d.seg_dx[0][t] = (signed char)(t % 3); d.seg_dy[0][t] = (signed char)(t / 3); }
            }
            long long K = (long long)taps * Cin;
            if (c.has_skip) {
                d.a_ptr[1] = bp(c.skip_in.t0); d.a_C[1] = c.skip_in.t0.C; d.a_ld[1] = c.skip_in.t0.C;
                d.seg_map[1] = 1; d.seg_taps[1] = 1; d.seg_kchunks[1] = c.skip_in.t0.C / 64; d.seg_cbase[1] = 0; d.nseg = 2;
                K += c.skip_in.t0.C;
                if (c.skip_in.two) {
                    d.a_ptr[2] = bp(c.skip_in.t1); d.a_C[2] = c.skip_in.t1.C; d.a_ld[2] = c.skip_in.t1.C;
                    d.seg_map[2] = 2; d.seg_taps[2] = 1; d.seg_kchunks[2] = c.skip_in.t1.C / 64; d.seg_cbase[2] = 0; d.nseg = 3;
                    K += c.skip_in.t1.C;
                }
            }
            d.b_ptr = c.wp; d.b_K = (int)K; d.b_rows = c.Co; d.b_batch = 1; d.b_ld = c.ldw; d.b_bs = 0;
            d.out = bp(c.out); d.ldo = c.Co; d.alpha = 1.f; d.grid_z = 1;
            d.bias = c.bias; d.rowvec = c.rowvec; d.rowvec_ld = c.rowvec_ld; d.rows_per_vec = c.Ho * c.Wo;
            d.residual = c.accumulate ? (const void*)bp(c.out) : (const void*)c.residual; d.ldr = c.Co;
            ++n_tc_gemms;
            // small-M problems (8x8 / 4x4 levels) do not fill 148 SMs with output tiles: split the K loop across CTAs
            // (fp32 atomics into a zeroed scratch) and finish bias / timestep vector / residual / bf16 in a tiny second kernel
            int bn = pick_block_n(c.Co);
            int tiles = (int)((Pout + 127) / 128) * ((c.Co + bn - 1) / bn);
            // 8x8 levels: 64 tiles of 128x256 would need split-K (+ a finalize launch and fp32 atomics); 128 tiles of 128x128
            // fill the machine in one wave with the same MMA time per CTA and no second pass
            static const bool no_n128 = getenv("DDPM_NO_N128_FILL") != nullptr;
            if (!no_n128 && bn == 256 && tiles <= 74 && tiles * 2 <= 148) { bn = 128; tiles *= 2; d.block_n = 128; }
            const int slabs = (int)(K / 64);
            int splits = 1;
            // (a finalize launch + fp32 atomics cost more than a half-empty single wave: split only when under ~1/3 of the SMs have a tile)
            if (tiles <= 48 && slabs >= 16) { splits = 148 / tiles; if (splits > slabs / 8) splits = slabs / 8; if (splits > 16) splits = 16; if (splits < 1) splits = 1; }
            if (splits > 1) {
                float* scratch = at<float>(alloc_once_zero((size_t)Pout * c.Co * 4));
                ddpm_gemm_desc ds = d;
                ds.kk_splits = splits; ds.grid_z = splits; ds.out = scratch; ds.flags = EPI_OUT_F32 | EPI_ATOMIC;
                ds.bias = nullptr; ds.rowvec = nullptr; ds.residual = nullptr;
                const float* bias = c.bias; const float* rowvec = c.rowvec; const int rvld = c.rowvec_ld, rpv = c.Ho * c.Wo;
                const bf16* resid = reinterpret_cast<const bf16*>(d.residual); bf16* outp = bp(c.out);
                const long long M = Pout; const int N = c.Co; const int nfin = grid_for(M * (N / 8));
                if (dry) { push(L, c.name, fl, [](cudaStream_t) { return 0; }, 2); return -1; }
                GemmLaunch g; int rc = build_gemm(ds, g);
                if (rc) { plan_error = rc; return -1; }
                push(L, c.name + "[splitk]", fl, [=](cudaStream_t st) {
                    const int r2 = launch_gemm(g, st); if (r2) return r2;
                    launch_k(k_splitk_finalize, nfin, 256, 0, st, scratch, bias, rowvec, rvld, rpv, resid, outp, M, N);
                    return (int)cudaGetLastError(); }, 2);
                return -1;
            }
            const long long qs = conv_gn_epi(c, d.gn);
            if (dry) { push(L, c.name, fl, [](cudaStream_t) { return 0; }); return qs; }
            GemmLaunch g; int rc = build_gemm(d, g);
            if (rc) { plan_error = rc; return qs; }
            push(L, c.name, fl, [g](cudaStream_t st) { return launch_gemm(g, st); });
            return qs;
        }
        ++n_generic;
        ConvG g; memset(&g, 0, sizeof g);
        g.in = gsrc(c.in); g.wp = c.wp; g.ldw = c.ldw; g.bias = c.bias; g.rowvec = c.rowvec; g.rowvec_ld = c.rowvec_ld;
        g.residual = c.residual; g.out = c.out_nchw ? (void*)c.out_nchw : (void*)bp(c.out); g.out_nchw_f32 = c.out_nchw ? 1 : 0;
        g.B = Bn; g.Hi = i0.H; g.Wi = i0.W; g.Ho = c.Ho; g.Wo = c.Wo; g.Co = c.Co; g.ksize = c.ksize; g.stride = c.stride;
        g.pad = (c.map == MAP_NORMAL && c.stride == 1) ? c.ksize / 2 : 0; g.map = c.map; g.accumulate = c.accumulate ? 1 : 0;
        const dim3 grid((unsigned)((Pout + 63) / 64), (unsigned)((c.Co + 63) / 64));
        const bool nchw_dyn = c.out_nchw != nullptr;
        UnetEngine* self = this;
        push(L, c.name + "[simt]", fl, [g, grid, nchw_dyn, self](cudaStream_t st) {
            ConvG gg = g; if (nchw_dyn && self->eps_dst) gg.out = self->eps_dst;
            launch_k(k_conv_generic, grid, 256, 0, st, gg); return (int)cudaGetLastError(); });
        if (c.has_skip) {
            ConvG s = g; s.in = gsrc(c.skip_in); s.wp = c.wp + (long long)taps * Cin; s.ksize = 1; s.pad = 0; s.map = MAP_NORMAL; s.stride = 1;
            s.bias = nullptr; s.rowvec = nullptr; s.residual = nullptr; s.accumulate = 1;
            push(L, c.name + ".skip[simt]", 0, [s, grid](cudaStream_t st) { launch_k(k_conv_generic, grid, 256, 0, st, s); return (int)cudaGetLastError(); });
        }
        return -1;
    }
    float* eps_dst = nullptr;   // run-time destination of the final conv (caller buffer or internal eps buffer)
    PsampleEpi ps_epi = {nullptr, nullptr, nullptr, 0, nullptr};   // run-time: sampler update fused into the final gather (x != null)
    QsamplePro qs_pro = {};           // run-time: q_sample fused into in_conv's loads (ddpm_train_forward)
    bool gather_fused_tail = false;   // plan property: the final conv ends in k_out_gather (tensor-core out_conv path)
    const float* deps_src = nullptr;   // run-time d(loss)/d(eps), fp32 NCHW (caller buffer or internal buffer)
    int plan_error = 0;

    // weight gradient of a conv: dW (OIHW fp32, flat grads) from dy [B,Ho,Wo,Co] and the conv input
    void wgrad_op(const std::string& name, const T4& dy, const Src& in, int ksize, int stride, int map, float* dw, int Co_valid) {
        const int Cin = in.C(), taps = ksize * ksize, Co = dy.C;
        const long long P = dy.pix();
        const double fl = 2.0 * P * Co_valid * (double)taps * Cin;
        bwd_flops += fl;
        const bool s2 = stride == 2 && ksize == 3 && !in.two && in.t0.H == 2 * dy.H && in.t0.W == 2 * dy.W;
        const bool tc = (stride == 1 || s2) && map == MAP_NORMAL && Co % 64 == 0 && Co == Co_valid && in.t0.C % 64 == 0 && (!in.two || in.t1.C % 64 == 0) &&
                        tc_ok_geom(dy.H, dy.W) && (s2 || in.t0.H == dy.H);
        if (tc) {
            float* scratch = nullptr; size_t sc_off = 0;
            if (taps == 9) { sc_off = alloc_once_zero((size_t)9 * Co * Cin * 4); scratch = at<float>(sc_off); }
            for (int s = 0; s < (in.two ? 2 : 1); ++s) {
                const T4& a = s ? in.t1 : in.t0; const int coff = s ? in.t0.C : 0;
                ddpm_gemm_desc d; memset(&d, 0, sizeof d);
                d.mode = GEMM_MNMN; d.M = Co; d.N = a.C; d.W = dy.W; d.H = dy.H; d.NB = dy.B;
                d.a_ptr[0] = bp(dy); d.a_C[0] = Co; d.a_ld[0] = Co;
                d.b_ptr = bp(a); d.b_K = a.C; d.b_ld = a.C;
                if (s2) { d.b_estride = 2; d.b_pad = 0; }
                d.taps = taps; d.kblocks = (int)((P + 63) / 64);
                const int bn = pick_block_n(a.C);
                const int tiles = ((Co + 127) / 128) * (a.C / bn) * taps;
                // split-K so that the launch is ONE full wave (<= 148 CTAs): more splits only add epilogue atomics,
                // and 148*k + 1 CTAs would cost a whole extra wave
                // (1x1 weight gradients are a few output tiles and HBM-bound: half a wave is faster in isolation - 23.3 vs 27.0 us for
                // 256->128 at 32x32, profiles/r02_wgrad_microbench.txt - and leaves the other SMs to the main chain: step 9.12 -> 9.07 ms for the 1-2 tile
                // ones, 9.047 -> 9.028 ms more with the 6-12 tile ones (project_in / project_out) included)
                int splits = ((taps == 1 && tiles <= 12) ? 74 : 148) / tiles; if (splits > d.kblocks / 4) splits = d.kblocks / 4; if (splits < 1) splits = 1;
                d.splits = splits; d.grid_z = taps * splits;
                d.flags = EPI_OUT_F32 | EPI_ATOMIC; d.alpha = 1.f;
                if (taps == 9) { d.out = scratch + coff; d.ldo = Cin; d.out_tap_stride = (long long)Co * Cin; }
                else { d.out = dw + coff; d.ldo = Cin; }
                ++n_tc_gemms;
                // weight gradients are leaves of the backward graph: run them on the side stream, where the tensor-bound GEMMs
                // overlap the memory-bound GroupNorm kernels of the main chain (DDPM_WGRAD_MAIN=1 keeps them in line)
                static const bool wg_side = getenv("DDPM_WGRAD_MAIN") == nullptr;
                if (dry) { push(bwd_ops, name, s ? 0 : fl, [](cudaStream_t) { return 0; }, 1, wg_side); continue; }
                GemmLaunch g; int rc = build_gemm(d, g);
                if (rc) { plan_error = rc; return; }
                push(bwd_ops, name, s ? 0 : fl, [g](cudaStream_t st) { return launch_gemm(g, st); }, 1, wg_side);
            }
            if (taps == 9) {   // packed [tap][Co][Ci] scratch -> OIHW flat gradient: batched into ONE table-driven launch at the end of backward
                PackEntry e; memset(&e, 0, sizeof e);
                e.kind = PK_UNPACK_GRAD; e.Co = Co; e.Ci = Cin; e.taps = 9; e.fout = dw; e.scratch = scratch;
                unpack_table_host.push_back(e);
            }
            return;
        }
        ++n_generic;
        WgradG w; memset(&w, 0, sizeof w);
        w.dy = bp(dy); w.in = gsrc(in); w.dw = dw; w.s_co = (long long)Cin * taps; w.s_ci = taps; w.s_tap = 1;
        w.B = dy.B; w.Hi = in.t0.H; w.Wi = in.t0.W; w.Ho = dy.H; w.Wo = dy.W; w.Co = Co; w.ksize = ksize; w.stride = stride;
        w.pad = (map == MAP_NORMAL && stride == 1) ? ksize / 2 : 0; w.map = map; w.Co_valid = Co_valid;
        const int tiles = ((Cin + 63) / 64) * ((Co + 63) / 64) * taps;
        int splits = (592 + tiles - 1) / tiles; const int maxs = (int)((P + 255) / 256); if (splits > maxs) splits = maxs; if (splits < 1) splits = 1;
        w.pix_per_split = (int)(((P + splits - 1) / splits + 15) / 16 * 16);
        splits = (int)((P + w.pix_per_split - 1) / w.pix_per_split);
        const dim3 grid((Cin + 63) / 64, (Co + 63) / 64, taps * splits);
        push(bwd_ops, name + "[simt]", fl, [w, grid](cudaStream_t st) { launch_k(k_wgrad_generic, grid, 256, 0, st, w); return (int)cudaGetLastError(); }, 1, true);
    }
    size_t alloc_once_zero(size_t bytes) {   // carve from the plan-time-zeroed arena (contiguous region grown on demand)
        const size_t o = alloc(bytes);
        once_list.push_back({o, bytes});
        return o;
    }
    std::vector<std::pair<size_t, size_t>> once_list;

    // per-pass zeroed arenas (GroupNorm statistic accumulators etc.): bump-allocated inside a region zeroed by ONE memset
    size_t zf_cursor = 0, zb_cursor = 0;
    size_t zero_fwd(size_t bytes) { const size_t o = zf_cursor; zf_cursor += (bytes + 255) & ~size_t(255); return zero_fwd_off + o; }
    size_t zero_bwd(size_t bytes) { const size_t o = zb_cursor; zb_cursor += (bytes + 255) & ~size_t(255); return zero_bwd_off + o; }

    // blocks over the pixels of one image: keep the whole launch within one wave of (148 SMs x occ) CTAs
    static void gn_grid(int Bn, int HW, int occ, int& nblk, int& ppb) {
        nblk = (148 * occ) / Bn; if (nblk > HW / 16) nblk = HW / 16; if (nblk < 1) nblk = 1;
        ppb = (HW + nblk - 1) / nblk; nblk = (HW + ppb - 1) / ppb;
    }
    GnSaved gn_fwd(std::vector<Op>& L, const std::string& name, const Src& in, const std::string& pname, const T4& out, int silu, float drop_p) {
        const int C = in.C(), Bn = in.t0.B, HW = in.t0.H * in.t0.W;
        GnSaved sv; sv.in = in; sv.silu = silu; sv.drop_p = drop_p; sv.layer = ++layer_counter;   // dropout is armed per call by a non-zero seed
        sv.gamma = PP(pname + ".weight"); sv.beta = PP(pname + ".bias"); sv.dgamma = GP(pname + ".weight"); sv.dbeta = GP(pname + ".bias");
        double* stats = at<double>(zero_fwd((size_t)Bn * 64 * 8));
        sv.K = at<float>(alloc((size_t)Bn * 4 * C * 4));
        const GnSrc gs = gsrc(in);
        const int thr = oct_threads(C);
        float* K = sv.K; const float* ga = sv.gamma; const float* be = sv.beta;
        sv.mask = (drop_p > 0.f && train) ? at<unsigned char>(alloc((size_t)Bn * HW * (C / 8))) : nullptr;
        // statistics delivered by the producers' epilogues (conv_gn_epi): one launch, no pass over the tensor for the stats
        const bool have_qs = in.t0.qs >= 0 && (!in.two || in.t1.qs >= 0) && (C % 128) == 0;
        if (have_qs) {
            static const int ap_occ = getenv("DDPM_GN_APPLY_OCC") ? atoi(getenv("DDPM_GN_APPLY_OCC")) : 4;
            GnApply a; memset(&a, 0, sizeof a);
            a.s = gs; a.K = K; a.y = bp(out); a.HW = HW; a.silu = silu; a.drop_p = sv.drop_p; a.seed = 0; a.layer = sv.layer; a.mask = sv.mask;
            a.qs0 = at<double>((size_t)in.t0.qs); a.qs1 = in.two ? at<double>((size_t)in.t1.qs) : nullptr;
            a.gamma = ga; a.beta = be; a.eps = 1e-6f; a.Kout = K;
            int nb2, ppb2; gn_grid(Bn, HW, ap_occ, nb2, ppb2);
            const dim3 g2(nb2, Bn);
            UnetEngine* self = this;
            push(L, name + ".apply[qs]", 0, [a, g2, thr, ppb2, self](cudaStream_t st) { GnApply aa = a; aa.seed = self->drop_seed; if (!aa.seed) aa.drop_p = 0.f;
                launch_k(k_gn_apply, g2, thr, 0, st, aa, ppb2); return (int)cudaGetLastError(); });
            return sv;
        }
        if (HW <= 64 && !getenv("DDPM_NO_GN_EPI")) {       // tiny maps: statistics + apply in ONE launch, one block per image
            GnApply a; memset(&a, 0, sizeof a);
            a.s = gs; a.K = K; a.y = bp(out); a.HW = HW; a.silu = silu; a.drop_p = sv.drop_p; a.seed = 0; a.layer = sv.layer; a.mask = sv.mask;
            a.gamma = ga; a.beta = be; a.eps = 1e-6f; a.Kout = K;
            UnetEngine* self = this;
            push(L, name + ".small", 0, [a, Bn, thr, self](cudaStream_t st) { GnApply aa = a; aa.seed = self->drop_seed; if (!aa.seed) aa.drop_p = 0.f;
                launch_k(k_gn_small, Bn, thr, 0, st, aa); return (int)cudaGetLastError(); });
            return sv;
        }
        static const int st_occ = getenv("DDPM_GN_STATS_OCC") ? atoi(getenv("DDPM_GN_STATS_OCC")) : 4;
        static const int ap_occ = getenv("DDPM_GN_APPLY_OCC") ? atoi(getenv("DDPM_GN_APPLY_OCC")) : 4;
        int nblk, ppb; gn_grid(Bn, HW, st_occ, nblk, ppb);
        const dim3 g1(nblk, Bn);
        GnFin fin; fin.gamma = ga; fin.beta = be; fin.K = K; fin.eps = 1e-6f; fin.ticket = at<int>(zero_fwd((size_t)Bn * 4));
        push(L, name + ".stats", 0, [=](cudaStream_t st) {
            launch_k(k_gn_stats, g1, thr, 0, st, gs, stats, HW, ppb, fin);
            return (int)cudaGetLastError(); });
        GnApply a; memset(&a, 0, sizeof a);
        a.s = gs; a.K = K; a.y = bp(out); a.HW = HW; a.silu = silu; a.drop_p = sv.drop_p; a.seed = 0; a.layer = sv.layer; a.mask = sv.mask;
        int nb2, ppb2; gn_grid(Bn, HW, ap_occ, nb2, ppb2);
        const dim3 g2(nb2, Bn);
        UnetEngine* self = this;
        push(L, name + ".apply", 0, [a, g2, thr, ppb2, self](cudaStream_t st) { GnApply aa = a; aa.seed = self->drop_seed; if (!aa.seed) aa.drop_p = 0.f;
            launch_k(k_gn_apply, g2, thr, 0, st, aa, ppb2); return (int)cudaGetLastError(); });
        return sv;
    }
    // dx(in) (=|+=) gn_bwd(dy) + addend ; dgamma/dbeta accumulate into the flat grads
    void gn_bwd(const std::string& name, const GnSaved& sv, const T4& dy, const bf16* addend,
                float* cs_per_img = nullptr, int cs_ld = 0, float* cs_total = nullptr, float* cs_total2 = nullptr) {
        const Src& in = sv.in;
        const int C = in.C(), Bn = in.t0.B, HW = in.t0.H * in.t0.W;
        GnBwd a; memset(&a, 0, sizeof a);
        a.s = gsrc(in); a.dy = bp(dy); a.K = sv.K; a.gamma = sv.gamma;
        a.cs = at<float>(zero_bwd((size_t)Bn * 2 * C * 4)); a.PQ = at<float>(alloc((size_t)Bn * 2 * C * 4));
        a.dgamma = sv.dgamma; a.dbeta = sv.dbeta;
        bool f0 = true, f1 = true;
        const T4 g0 = grad_of(in.t0, &f0); a.dx0 = bp(g0); a.acc0 = f0 ? 0 : 1;
        if (in.two) { const T4 g1 = grad_of(in.t1, &f1); a.dx1 = bp(g1); a.acc1 = f1 ? 0 : 1; }
        a.addend = addend; a.B = Bn; a.HW = HW; a.silu = sv.silu; a.drop_p = sv.drop_p; a.layer = sv.layer; a.mask = sv.mask;
        a.ticket = at<int>(zero_bwd((size_t)Bn * 4));
        static const bool no_dn = getenv("DDPM_GN_NO_DN") != nullptr;
        a.dn_inplace = (sv.silu && !no_dn) ? 1 : 0;
        const int thr = oct_threads(C);
        // Grid granularity measured on the whole step (bench.py, DDPM_GN_BWD_OCC / _STATS_OCC / _APPLY_OCC sweeps): blocks per
        // image = 148*occ/B; occ 4 is best for all three passes (10.49 ms/step; occ 2: 10.69, 8: 11.04, 16: 11.71) - per-block
        // prologue / atomics cost more than the wave-quantisation they would save.
        static const int gn_bwd_occ = getenv("DDPM_GN_BWD_OCC") ? atoi(getenv("DDPM_GN_BWD_OCC")) : 4;
        int nblk, ppb; gn_grid(Bn, HW, gn_bwd_occ, nblk, ppb);
        a.pix_per_block = ppb;
        const dim3 g1(nblk, Bn);
        const size_t shm = (size_t)2 * C * 4;
        const size_t shm2 = (cs_per_img || cs_total || cs_total2) ? (size_t)C * 4 : 0;
        UnetEngine* self = this;
        push(bwd_ops, name + ".gn_bwd", 0, [=](cudaStream_t st) {
            GnBwd aa = a; aa.seed = self->drop_seed; if (!aa.seed) aa.drop_p = 0.f;
            launch_k(k_gn_bwd_reduce, g1, thr, shm, st, aa);
            if (shm2) launch_k(k_gn_bwd_apply<true>, g1, thr, shm2, st, aa, cs_per_img, cs_ld, cs_total, cs_total2);
            else      launch_k(k_gn_bwd_apply<false>, g1, thr, 0, st, aa, cs_per_img, cs_ld, cs_total, cs_total2);
            return (int)cudaGetLastError(); }, 2);
    }
    void colsum_op(const std::string& name, const T4& dy, float* per_img, int ld, float* total, float* total2, int C_valid) {
        const int HW = dy.H * dy.W, C = dy.C, Bn = dy.B;
        const int thr = oct_threads(C);
        int nblk = (592 + Bn - 1) / Bn; if (nblk > HW / 8) nblk = HW / 8; if (nblk < 1) nblk = 1;
        const int ppb = (HW + nblk - 1) / nblk; nblk = (HW + ppb - 1) / ppb;
        const dim3 g(nblk, Bn); const bf16* p = bp(dy);
        const size_t shm = (size_t)C * 4;
        push(bwd_ops, name + ".colsum", 0, [=](cudaStream_t st) { launch_k(k_colsum, g, thr, shm, st, p, per_img, ld, total, total2, HW, C, C_valid, ppb); return (int)cudaGetLastError(); }, 1, /*side=*/true);
    }

    // ------------------------------------------------------------------ packed weights (table-driven: one launch re-packs everything)
    struct Packed { bf16* fwd = nullptr; long long ld_f = 0; bf16* dgr = nullptr; long long ld_d = 0; };
    std::vector<PackEntry> pack_table_host, unpack_table_host; size_t pack_table_off = 0, unpack_table_off = 0;
    // fwd pack [Co][taps*Ci (+extra)], dgrad pack [Ci][taps*Co] (dkind 1 plain / 2 flipped taps / 3 stride-2 parity blocks)
    Packed pack_conv(const std::string& pname, int Co, int Ci, int ksize, int extra_k, bool flip, bool want_dgrad, int dkind_override = 0) {
        const int taps = ksize * ksize;
        Packed pk; pk.ld_f = (long long)taps * Ci + extra_k;
        pk.fwd = at<bf16>(alloc((size_t)Co * pk.ld_f * 2));
        if (want_dgrad && train) { pk.ld_d = (long long)taps * Co; pk.dgr = at<bf16>(alloc((size_t)Ci * pk.ld_d * 2)); }
        PackEntry e; memset(&e, 0, sizeof e);
        e.kind = PK_CONV; e.Co = Co; e.Ci = Ci; e.taps = taps; e.k_off = 0; e.w = PP(pname + ".weight");
        e.fwd = pk.fwd; e.ld_f = pk.ld_f; e.dgr = pk.dgr; e.ld_d = pk.ld_d;
        e.dkind = pk.dgr ? (dkind_override ? dkind_override : (flip ? 2 : 1)) : 0;
        pack_table_host.push_back(e);
        return pk;
    }
    void pack_extra(const std::string& pname, const Packed& into, int k_off, int Co, int Ci, bf16* dgr, long long ld_d) {
        PackEntry e; memset(&e, 0, sizeof e);
        e.kind = PK_CONV; e.Co = Co; e.Ci = Ci; e.taps = 1; e.k_off = k_off; e.w = PP(pname + ".weight");
        e.fwd = into.fwd; e.ld_f = into.ld_f; e.dgr = dgr; e.ld_d = ld_d; e.dkind = dgr ? 1 : 0;
        pack_table_host.push_back(e);
    }
    void pack_identity(const Packed& into, int k_off, int C) {
        PackEntry e; memset(&e, 0, sizeof e);
        e.kind = PK_IDENTITY; e.Co = C; e.Ci = C; e.taps = 1; e.k_off = k_off; e.fwd = into.fwd; e.ld_f = into.ld_f;
        pack_table_host.push_back(e);
    }
    void pack_bias_add(float* dst, const float* a, const float* b, int n) {
        PackEntry e; memset(&e, 0, sizeof e);
        e.kind = PK_BIAS_ADD; e.Co = n; e.Ci = 1; e.taps = 1; e.w = a; e.w2 = b; e.fout = dst;
        pack_table_host.push_back(e);
    }

    // ------------------------------------------------------------------ blocks
    T4 res_block(const std::string& p, const Src& x, int cout, int tp_off, int tp_ld, float* TP, float* dTP);
    T4 attn_block(const std::string& p, const T4& x);
    T4 down_conv(const std::string& p, const T4& x);
    T4 up_conv(const std::string& p, const T4& x);
    void bmm(std::vector<Op>& L, const std::string& name, int form, const bf16* A, long long lda, long long sa, const bf16* Bp, long long ldb, long long sb,
             void* C, long long ldc, long long sc, bool c_f32, int nb, int T, int Cc, float alpha, double* fl_acc);
    int plan(int B_, int H_, int W_, bool train_, bool dry_);
    int build();
    // Ops flagged `side` (bias-gradient column sums, gradient unpacks: small, memory-bound, nothing downstream needs them
    // before the end of the pass) are forked onto a second stream so they overlap the tensor-core kernels, and joined at
    // the end of the list.  Fork/join are event edges, so the pattern is CUDA-graph capturable.
    // (Stream priorities were measured both ways on the whole step - main chain high: 10.17 ms, side stream high: 9.46 ms, equal: 9.06 ms -
    // and are not used; profiles/r02_scheduling_ab.txt.)
    cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Weight repack of a TRAINING plan (every step, after the optimizer): issued on the side stream so that it overlaps the ops
    // of the next forward that do not read packed weights (timestep embedding MLP on the fp32 masters, q_sample + in_conv);
    // the forward list waits on ev_pack_fc / ev_pack_all right before the first consumer (Op::wait_pack).
    cudaEvent_t ev_pack_fork = nullptr, ev_pack_fc = nullptr, ev_pack_all = nullptr; bool pack_pending = false;
    int repack(cudaStream_t st) {
        static const bool inline_pack = getenv("DDPM_PACK_INLINE") != nullptr;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone; cudaStreamIsCapturing(st, &cs);
        if (!train || !side_stream || !ev_pack_all || inline_pack || cs != cudaStreamCaptureStatusNone || getenv("DDPM_DEBUG_SYNC") || getenv("DDPM_NO_SIDE_STREAM"))
            return run_list(pack_ops, st);
        if (cudaEventRecord(ev_pack_fork, st) || cudaStreamWaitEvent(side_stream, ev_pack_fork, 0)) return fail(-20, "repack fork failed");
        for (int pass = 0; pass < 2; ++pass) {        // early packs (in_conv, timestep MLP: Op::wait_pack == 1) first, then the rest
            for (auto& o : pack_ops) {
                if ((o.wait_pack == 1) != (pass == 0)) continue;
                const int rc = o.run(side_stream);
                if (rc) return fail(-20, "op '%s' failed: %s", o.name.c_str(), cudaGetErrorString((cudaError_t)rc));
            }
            if (pass == 0 && cudaEventRecord(ev_pack_fc, side_stream)) return fail(-20, "repack event failed");
        }
        if (cudaEventRecord(ev_pack_all, side_stream)) return fail(-20, "repack event failed");
        pack_pending = true;
        return 0;
    }
    int run_list(std::vector<Op>& L, cudaStream_t caller_st) {
        cudaStream_t st = caller_st;
        static const bool dbg = getenv("DDPM_DEBUG_SYNC") != nullptr;   // serialise + attribute faults to an op (never under graph capture)
        static const bool no_side = getenv("DDPM_NO_SIDE_STREAM") != nullptr;
        bool forked = false;
        // DDPM_OP_TIMING=<file>: CUDA events after every main-stream op (side stream stays concurrent), appended as "name ms" lines
        static const char* timing_path = getenv("DDPM_OP_TIMING");
        std::vector<cudaEvent_t> tev;
        if (timing_path && !dbg) { tev.resize(L.size() + 1); for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], st); }
        size_t oi = 0; int main_since_fork = 0;
        const bool skip_temb = uniform_t && !temb_uni_ops.empty() && &L == &fwd_ops;
        for (auto& o : L) {
            int rc;
            ++oi;
            if (skip_temb && o.tag == OP_TEMB) {
                if (o.name == "temb.sin") for (auto& u : temb_uni_ops) { rc = u.run(st); if (rc) return fail(-20, "op '%s' failed: %s", u.name.c_str(), cudaGetErrorString((cudaError_t)rc)); }
                if (!tev.empty()) cudaEventRecord(tev[oi], st);
                continue;
            }
            if (o.wait_pack && pack_pending) {
                if (cudaStreamWaitEvent(st, o.wait_pack == 1 ? ev_pack_fc : ev_pack_all, 0)) return fail(-20, "repack join failed");
                if (o.wait_pack == 2) pack_pending = false;
            }
            if (o.chunk >= 0) {
                // gradient-chunk boundary: everything the main chain has written so far is joined into the side stream (which
                // carries the weight-gradient GEMMs), the chunk's gradient unpack runs there, then the chunk's event is recorded
                cudaStream_t cs_ = (side_stream && !dbg && !no_side) ? side_stream : st;
                rc = 0;
                if (cs_ != st) { rc = (int)cudaEventRecord(ev_fork, st); if (!rc) rc = (int)cudaStreamWaitEvent(cs_, ev_fork, 0); main_since_fork = 0; forked = true; }
                if (!rc) rc = o.run(cs_);
                if (!rc && chunks[o.chunk].ev) rc = (int)cudaEventRecord(chunks[o.chunk].ev, cs_);
                if (rc) return fail(-20, "op '%s' failed: %s", o.name.c_str(), cudaGetErrorString((cudaError_t)rc));
                if (!tev.empty()) cudaEventRecord(tev[oi], st);
                continue;
            }
            if (o.side && side_stream && !dbg && !no_side) {
                rc = 0;
                if (main_since_fork || !forked) {     // consecutive side ops share one fork edge
                    rc = (int)cudaEventRecord(ev_fork, st);
                    if (!rc) rc = (int)cudaStreamWaitEvent(side_stream, ev_fork, 0);
                    main_since_fork = 0;
                }
                if (!rc) rc = o.run(side_stream);
                forked = true;
            } else {
                rc = o.run(st);
                ++main_since_fork;
            }
            if (!rc && dbg) rc = (int)cudaStreamSynchronize(st);
            if (rc) return fail(-20, "op '%s' failed: %s", o.name.c_str(), cudaGetErrorString((cudaError_t)rc));
            if (!tev.empty()) cudaEventRecord(tev[oi], st);
        }
        if (forked) {
            int rc = (int)cudaEventRecord(ev_join, side_stream);
            if (!rc) rc = (int)cudaStreamWaitEvent(st, ev_join, 0);
            if (rc) return fail(-20, "side-stream join failed: %s", cudaGetErrorString((cudaError_t)rc));
        }
        if (!tev.empty()) {
            cudaEvent_t fin; cudaEventCreate(&fin); cudaEventRecord(fin, st); cudaStreamSynchronize(st);
            if (FILE* f = fopen(timing_path, "a")) {
                for (size_t i = 0; i < L.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, tev[i], tev[i + 1]); fprintf(f, "%s%s %.5f\n", L[i].side ? "[side]" : "", L[i].name.c_str(), ms); }
                float ms = 0; cudaEventElapsedTime(&ms, tev[L.size()], fin); fprintf(f, "[join] %.5f\n", ms);
                fclose(f);
            }
            for (auto& e : tev) cudaEventDestroy(e);
            cudaEventDestroy(fin);
        }
        return 0;
    }
};

}  // namespace ddpm
