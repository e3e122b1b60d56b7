// C ABI of libdgprf (see include/dgprf.h): argument validation, workspace layout and the
// per-layer launch sequences.  No model state is kept between calls; the only things remembered are plans that are pure
// functions of the arguments (the last workspace layout / step geometry of this thread, re-derived whenever the model
// description, the batch size, the mode or the process environment differs).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "kernels.cuh"

static thread_local char g_err[512] = "";

void dgprf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* dgprf_last_error(void) { return g_err; }

// Signature of the process environment: the debug / A-B switches are read with getenv, and a sampler calls the step tens of
// thousands of times per second, so plans are cached per thread and keyed (among others) by this value.  setenv / putenv /
// unsetenv replace or move the pointer of the entry they touch (glibc never edits a "NAME=value" string in place), so the
// pointers alone identify the contents; ~100 pointer reads instead of ~15 getenv string scans per step.
extern char** environ;
uint64_t dgprf_env_signature(void) {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    if (environ)
        for (char** e = environ; *e; ++e) h = (h ^ (uint64_t)(uintptr_t)*e) * 0x100000001b3ull;
    return h;
}

bool dgprf_pdl_enabled() {
    static thread_local uint64_t env = 0;
    static thread_local bool on = true, known = false;
    const uint64_t e = dgprf_env_signature();
    if (!known || e != env) { on = getenv("DGPRF_NO_PDL") == nullptr; env = e; known = true; }
    return on;
}

// ---- per-(kernel, device) shared-memory opt-in -------------------------------------------------
#include <mutex>
#include <vector>
struct SmemRec { const void* k; int dev; size_t smem; };
int dgprf_ensure_smem(const void* kernel, size_t smem) {
    static std::mutex mu;
    static std::vector<SmemRec> recs;
    int dev = 0;
    DGPRF_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (auto& r : recs)
        if (r.k == kernel && r.dev == dev) {
            if (smem > r.smem) {
                DGPRF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                r.smem = smem;
            }
            return DGPRF_OK;
        }
    DGPRF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    recs.push_back({kernel, dev, smem});
    return DGPRF_OK;
}

// ---- measurement hook ---------------------------------------------------------------------------
#include <vector>
#include <string>
struct ProfRec { std::string name; cudaEvent_t a, b; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;

void dgprf_prof_begin(const char* name, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r; r.name = name;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof->push_back(r);
}
void dgprf_prof_end(cudaStream_t st) {
    if (!g_prof_on || g_prof->empty()) return;
    cudaEventRecord(g_prof->back().b, st);
}
extern "C" int dgprf_profile_start(void) {
    if (!g_prof) g_prof = new std::vector<ProfRec>();
    for (auto& r : *g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof->clear();
    g_prof_on = true;
    return DGPRF_OK;
}
extern "C" int dgprf_profile_stop(int max_records, char* names, float* ms, int* n_out) {
    g_prof_on = false;
    DGPRF_REQUIRE(names && ms && n_out && max_records >= 0, "profile_stop: bad arguments");
    int n = 0;
    if (g_prof) {
        for (auto& r : *g_prof) {
            if (n < max_records) {
                DGPRF_CHECK_CUDA(cudaEventSynchronize(r.b));
                float t = 0.f;
                DGPRF_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
                ms[n] = t;
                snprintf(names + 32 * n, 32, "%s", r.name.c_str());
                ++n;
            }
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        g_prof->clear();
    }
    *n_out = n;
    return DGPRF_OK;
}
extern "C" int dgprf_version(void) { return 100; }

// ---- workspace layout ------------------------------------------------------------------------
struct LayerWs {
    int CS;        // column splits of the backward (slabs of Dpart / Tpart / Rpart)
    int CSf;       // column splits of the forward (slabs of Fpart)
    int tc_cols;   // 0: SIMT forward; 32 | 64: tensor-core forward with that column-tile width
    int64_t n_phi, n_fpart, n_dpart, n_tpart, n_rpart;       // floats per chain
    size_t phi, fpart, dpart, tpart, rpart;                  // byte offsets of the [C][...] regions
    int phi_blocked;                                         // saved features in the tile-blocked layout (both sides pipelined)
    int tc2;                                                 // pipelined TC forward: prepped operand buffers below
    int64_t n_zt, n_wt, n_at, n_ot; size_t zt, wt, at, ot;
    int64_t n_wp; size_t wp; int bwd2;                       // pipelined TC backward: padded W rows
};
struct WsLayout {
    LayerWs L[DGPRF_MAX_LAYERS];
    int RS;        // row splits of the layered backward (gW slabs)
    int RSF;       // row groups of the row-fused step (0: not eligible) -- also gW slabs
    int K10;       // the row-fused step is the cluster-split tensor-pipe kernel (RSF = its row tiles), else K9
    int64_t w_len, h_len, n_dflast, n_gwpart, n_llpart;
    size_t dflast, gwpart, ghyp, llsum, likpart, llpart, gridbar, dfsum, hpart, fsum, sumctr, total;
    int64_t n_sumctr;
    int64_t n_dfsum, n_hpart, n_fsum;
};

static inline int layer_F(const dgprf_layer& l) { return l.kind == DGPRF_KIND_RBF ? 2 * l.M : l.M; }
static inline int layer_d(const dgprf_layer& l) { return l.d_prev + l.d_x; }

static int validate_model(const dgprf_model* m) {
    DGPRF_REQUIRE(m != nullptr, "model is NULL");
    DGPRF_REQUIRE(m->n_layers >= 1 && m->n_layers <= DGPRF_MAX_LAYERS, "n_layers=%d out of [1,%d]", m->n_layers, DGPRF_MAX_LAYERS);
    DGPRF_REQUIRE(m->n_chains >= 1, "n_chains=%d", m->n_chains);
    DGPRF_REQUIRE(m->likelihood == DGPRF_LIK_GAUSSIAN || m->likelihood == DGPRF_LIK_SOFTMAX, "unknown likelihood %d", m->likelihood);
    DGPRF_REQUIRE(m->precision == DGPRF_PREC_FP32 || m->precision == DGPRF_PREC_TF32, "unknown precision %d", m->precision);
    DGPRF_REQUIRE(m->w_base && m->h_base, "parameter buffers are NULL");
    DGPRF_REQUIRE(m->likelihood != DGPRF_LIK_GAUSSIAN || m->off_lik_log_var >= 0, "Gaussian likelihood needs lik_log_var");
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        DGPRF_REQUIRE(y.kind == DGPRF_KIND_RBF || y.kind == DGPRF_KIND_ARC, "layer %d: unknown kind %d", l, y.kind);
        DGPRF_REQUIRE(y.M >= 1 && y.g >= 1 && y.g <= 64, "layer %d: M=%d g=%d unsupported (need M>=1, 1<=g<=64)", l, y.M, y.g);
        DGPRF_REQUIRE(y.d_prev >= 0 && y.d_x >= 0 && layer_d(y) >= 1 && layer_d(y) <= 8192, "layer %d: input width %d unsupported", l, layer_d(y));
        DGPRF_REQUIRE(y.d_x <= m->d_in, "layer %d: d_x=%d > d_in=%d", l, y.d_x, m->d_in);
        DGPRF_REQUIRE(l == 0 ? y.d_prev == 0 : y.d_prev == m->layer[l - 1].g, "layer %d: d_prev=%d does not chain", l, y.d_prev);
        DGPRF_REQUIRE(y.z != nullptr, "layer %d: z is NULL", l);
        DGPRF_REQUIRE(y.off_W >= 0 && (y.off_W & 3) == 0, "layer %d: off_W must be a non-negative multiple of 4", l);
    }
    DGPRF_REQUIRE(m->layer[m->n_layers - 1].g == m->d_out, "last n_gp=%d != d_out=%d", m->layer[m->n_layers - 1].g, m->d_out);
    return DGPRF_OK;
}

static int make_layout(const dgprf_model* m, int B, int mode, WsLayout* w) {
    DGPRF_REQUIRE(B >= 1, "B=%d", B);
    DGPRF_REQUIRE(mode >= DGPRF_MODE_EVAL && mode <= DGPRF_MODE_HYPER, "unknown mode %d", mode);
    memset(w, 0, sizeof(*w));
    const int64_t C = m->n_chains;
    size_t off = 0;
    auto take = [&](int64_t floats_per_chain) {
        const size_t at = off;
        off += (size_t)round_up(floats_per_chain * C * (int64_t)sizeof(float), 256);
        return at;
    };
    w->RS = row_splits(B);
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        LayerWs& s = w->L[l];
        const int cs0 = col_splits(y.M);
        s.CS = cs0;
        s.CSf = cs0;
        s.tc_cols = 0;
        if (m->precision == DGPRF_PREC_TF32 && (y.M % 4) == 0 && y.g <= 64) {
            // the pipelined tensor-core forward when the layer has enough tiles for it, else the fp32 SIMT kernel
            const int c2 = dgprf_fwd_tc2_col_splits(dgprf_tc_tile_cols(B, y.M, m->n_chains), B, layer_d(y), y.M, y.g, m->n_chains);
            if (c2 > 0) { s.tc2 = 1; s.CSf = c2; s.tc_cols = 64; }
        }
        s.n_fpart = (int64_t)s.CSf * B * y.g;
        s.fpart = take(s.n_fpart);
        if (mode >= DGPRF_MODE_TRAIN && m->precision == DGPRF_PREC_TF32) {     // pipelined TC backward: its own column splits
            const int pc = dgprf_bwd_tc2_pick_cs(B, y.M, y.g, layer_d(y), y.d_prev, w->RS, m->n_chains, mode == DGPRF_MODE_HYPER);
            if (pc > 0) { s.bwd2 = 1; s.CS = pc; }
        }
        if (mode >= DGPRF_MODE_TRAIN) {
            s.n_phi = (int64_t)B * layer_F(y);
            // both sides of the saved-feature round trip are the pipelined kernels: keep it tile-blocked (kernels.cuh)
            if (s.tc2 && s.bwd2 && !getenv("DGPRF_PHI_ROWMAJOR") && !getenv("DGPRF_TC2_DIRECT_STORE")) {
                s.phi_blocked = 1;
                s.n_phi = dgprf_phi_blocked_floats(B, y.M, y.kind);
            }
            s.phi = take(s.n_phi);
            if (l > 0) { s.n_dpart = (int64_t)s.CS * B * y.d_prev; s.dpart = take(s.n_dpart); }
        }
        if (mode == DGPRF_MODE_HYPER) {
            s.n_tpart = (int64_t)s.CS * B * layer_d(y); s.tpart = take(s.n_tpart);
            s.n_rpart = (int64_t)s.CS * B;              s.rpart = take(s.n_rpart);
        }
        if (s.bwd2) {
            s.n_wp = dgprf_bwd_tc2_wp_floats(layer_F(y));
            s.wp = take(s.n_wp);
        }
        if (s.tc2) {
            if (layer_d(y) <= 128) {
                s.n_zt = dgprf_fwd_tc2_zt_floats(y.M);      // one copy per chain (only the first is used when z is shared)
                s.zt = take(s.n_zt);
            } else {                                        // WIDE variant: input and Omega^T, tf32 hi / lo
                s.n_at = dgprf_fwd_tc2_at_floats(B, layer_d(y));
                s.at = take(s.n_at);
                s.n_ot = dgprf_fwd_tc2_ot_floats(y.M, layer_d(y));
                s.ot = take(s.n_ot);
            }
            s.n_wt = dgprf_fwd_tc2_wt_floats(layer_F(y), y.g);
            s.wt = take(s.n_wt);
        }
        const int64_t wend = y.off_W + (int64_t)layer_F(y) * y.g;
        if (wend > w->w_len) w->w_len = wend;
        int64_t hend = y.off_log_amp + 1;
        if (y.off_log_inv_ls + layer_d(y) > hend) hend = y.off_log_inv_ls + layer_d(y);
        if (y.has_mean && y.off_mean + layer_d(y) > hend) hend = y.off_mean + layer_d(y);
        if (hend > w->h_len) w->h_len = hend;
    }
    if (m->off_lik_log_var + 1 > w->h_len) w->h_len = m->off_lik_log_var + 1;
    {   // when every layer takes the pipelined backward (128-row tiles), row splits beyond the last row tile would only be
        // zero slabs for the update to add: cap them (the per-layer column splits above already assumed min(RS, row tiles))
        bool all_bwd2 = mode >= DGPRF_MODE_TRAIN;
        for (int l = 0; l < m->n_layers; ++l) all_bwd2 = all_bwd2 && w->L[l].bwd2;
        const int n_rt128 = ceil_div(B, 128);
        if (all_bwd2 && w->RS > n_rt128) w->RS = n_rt128;
    }
    w->w_len = round_up(w->w_len, 4);
    w->h_len = round_up(w->h_len, 4);
    w->llsum = take(1);
    w->likpart = take(2 * 64);
    if (mode >= DGPRF_MODE_TRAIN) {
        w->n_dflast = (int64_t)B * m->d_out;
        w->dflast = take(w->n_dflast);
        const int k10_tiles = dgprf_step_cluster_tiles(m, B);
        w->K10 = k10_tiles > 0 ? 1 : 0;
        w->RSF = k10_tiles > 0 ? k10_tiles : (dgprf_step_rows_eligible(m, B) ? dgprf_step_rows_groups(B) : 0);
        w->n_gwpart = (int64_t)(w->RS > w->RSF ? w->RS : w->RSF) * w->w_len;
        w->gwpart = take(w->n_gwpart);
        w->n_llpart = w->RSF > 0 ? w->RSF : 1;
        w->llpart = take(w->n_llpart);
        w->gridbar = off;                               // two 32-bit words {count, generation}, zero-initialised
        off += 256;
    }
    if (mode == DGPRF_MODE_HYPER) {
        w->ghyp = take(w->h_len);
        int dmax = 1;
        for (int l = 0; l < m->n_layers; ++l) if (layer_d(m->layer[l]) > dmax) dmax = layer_d(m->layer[l]);
        w->n_hpart = (int64_t)dgprf_hyper_row_blocks(B) * (2 * dmax + 1);      // row-block partials of the hyper reduction
        w->hpart = take(w->n_hpart);
    }
    for (int l = 0; l + 1 < m->n_layers; ++l)          // pre-summed dF of a pipelined TC backward (one slab instead of CS)
        if (w->L[l].bwd2 && (int64_t)B * m->layer[l].g > w->n_dfsum) w->n_dfsum = (int64_t)B * m->layer[l].g;
    if (w->n_dfsum > 0) w->dfsum = take(2 * w->n_dfsum);      // two buffers: layer l reads one while its kernel's tail fills the other
    for (int l = 1; l < m->n_layers; ++l)              // pre-summed F_{l-1} for a pipelined TC forward whose input comes in several slabs
        if (w->L[l].tc2 && w->L[l - 1].CSf > 1 && (int64_t)B * m->layer[l - 1].g > w->n_fsum) w->n_fsum = (int64_t)B * m->layer[l - 1].g;
    if (w->n_fsum > 0) w->fsum = take(2 * w->n_fsum);
    // tickets of the fused slab sums (pipelined kernels: the last column split of a row block adds the slabs): [fwd | bwd] x
    // [chains][128-row blocks], zero between launches (the workspace is zero-initialised, a ticket resets itself)
    w->n_sumctr = ceil_div(B, 128);
    w->sumctr = take(2 * w->n_sumctr);
    w->total = off;
    return DGPRF_OK;
}

static inline float* wsf(void* ws, size_t off) { return reinterpret_cast<float*>(static_cast<char*>(ws) + off); }

static int check_ws(const dgprf_model* m, int B, int mode, void* ws, size_t ws_bytes, WsLayout* w) {
    // the layout is a pure function of (model description, B, mode, environment): keep the last one of this thread
    struct Cached { bool ok; dgprf_model m; int B, mode; uint64_t env; WsLayout w; };
    static thread_local Cached c = {};
    const uint64_t env = dgprf_env_signature();
    if (m && c.ok && c.B == B && c.mode == mode && c.env == env && memcmp(&c.m, m, sizeof(*m)) == 0) *w = c.w;
    else {
        c.ok = false;
        int rc = validate_model(m);
        if (rc) return rc;
        rc = make_layout(m, B, mode, w);
        if (rc) return rc;
        memcpy(&c.m, m, sizeof(*m)); c.B = B; c.mode = mode; c.env = env; c.w = *w; c.ok = true;
    }
    DGPRF_REQUIRE(ws != nullptr, "workspace is NULL");
    DGPRF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
    if (ws_bytes < w->total) {
        dgprf_set_error("workspace too small: %zu < %zu bytes", ws_bytes, w->total);
        return DGPRF_EWORKSPACE;
    }
    return DGPRF_OK;
}

extern "C" int dgprf_workspace_bytes(const dgprf_model* m, int B, int mode, size_t* bytes) {
    DGPRF_REQUIRE(bytes != nullptr, "bytes is NULL");
    int rc = validate_model(m);
    if (rc) return rc;
    WsLayout w;
    rc = make_layout(m, B, mode, &w);
    if (rc) return rc;
    *bytes = w.total;
    return DGPRF_OK;
}

static SlabMat fpart_of(const dgprf_model* m, const WsLayout& w, void* ws, int l, int B) {
    SlabMat s;
    s.ptr = wsf(ws, w.L[l].fpart);
    s.cs = w.L[l].n_fpart;
    s.ss = (int64_t)B * m->layer[l].g;
    s.ld = m->layer[l].g;
    s.n_slabs = w.L[l].CSf;
    return s;
}

// ---- forward -----------------------------------------------------------------------------------
static int forward_impl(const dgprf_model* m, const WsLayout& w, const float* X, int64_t x_cs, int B, int mode,
                        void* ws, float* F_out, cudaStream_t st) {
    // ---- one launch prepares the TMA operands of every pipelined tensor-core layer (they depend on parameters only) ----
    bool prepped = false;
    {
        PrepArgs p;
        memset(&p, 0, sizeof(p));
        p.n_layers = m->n_layers;
        for (int l = 0; l < m->n_layers; ++l) {
            const dgprf_layer& y = m->layer[l];
            const LayerWs& s = w.L[l];
            PrepLayer& q = p.L[l];
            q.z = y.z; q.z_cs = y.z_cs;
            q.log_inv_ls = m->h_base + y.off_log_inv_ls; q.mean = y.has_mean ? m->h_base + y.off_mean : nullptr; q.h_cs = m->h_cs;
            q.W = m->w_base + y.off_W; q.w_cs = m->w_cs;
            q.d = layer_d(y); q.M = y.M; q.F = layer_F(y); q.g = y.g; q.has_mean = y.has_mean;
            q.NG = y.g <= 16 ? 16 : (y.g <= 32 ? 32 : 64);
            q.Kp = (q.d + 31) & ~31;
            if (s.tc2) {
                q.wt = wsf(ws, s.wt); q.n_wt = ceil_div(q.F, 32) * ceil_div(q.NG, 32);
                if (s.n_zt > 0) { q.zt = wsf(ws, s.zt); q.n_zt = ceil_div(q.M, 32) * 4; }
                if (s.n_ot > 0) { q.ot = wsf(ws, s.ot); q.n_ot = ceil_div(q.M, 32) * (q.Kp / 32); }
                prepped = true;
            }
            if (mode >= DGPRF_MODE_TRAIN && s.bwd2) {
                q.wp = wsf(ws, s.wp); q.n_wp = ceil_div(q.F * 8, 256);
                prepped = true;
            }
        }
        if (prepped) {
            const int rc = dgprf_launch_prep_layers(p, m->n_chains, st);
            if (rc) return rc;
        }
    }
    // opt-in (DGPRF_FUSED_SLAB_SUMS=1), measured slower than the launches it removes: the ticket + sum are three to five DEPENDENT
    // round trips at a CTA's tail (~1.8 us each under the other CTAs' streams), a PDL-overlapped k_sum_slabs launch is 7-17 us
    // off the critical path of nothing (profiles/r02_fused_slab_sums.txt: configs[4] layer backward 523 -> 744 us, forward +8 us)
    const bool fuse_sums = getenv("DGPRF_FUSED_SLAB_SUMS") != nullptr;
    bool fsum_ready = false;                 // the forward kernel of layer l - 1 already summed its slabs into fsum[(l - 1) & 1]
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        FwdArgs a;
        memset(&a, 0, sizeof(a));
        a.kind = y.kind; a.B = B; a.d_prev = y.d_prev; a.d_x = y.d_x; a.d = layer_d(y);
        a.M = y.M; a.g = y.g; a.F = layer_F(y); a.CS = w.L[l].CSf; a.ldx = m->d_in; a.do_gemm2 = 1;
        a.has_mean = y.has_mean; a.tile_cols = w.L[l].tc_cols;
        if (l > 0) a.Fprev = fpart_of(m, w, ws, l - 1, B);
        a.X = X; a.x_cs = x_cs;
        a.z = y.z; a.z_cs = y.z_cs;
        a.log_inv_ls = m->h_base + y.off_log_inv_ls;
        a.log_amp = m->h_base + y.off_log_amp;
        a.mean = y.has_mean ? m->h_base + y.off_mean : nullptr;
        a.h_cs = m->h_cs;
        a.W = m->w_base + y.off_W; a.w_cs = m->w_cs;
        a.Phi = mode >= DGPRF_MODE_TRAIN ? wsf(ws, w.L[l].phi) : nullptr;
        a.phi_cs = w.L[l].n_phi;
        a.phi_blocked = (mode >= DGPRF_MODE_TRAIN) ? w.L[l].phi_blocked : 0;
        a.Fpart = wsf(ws, w.L[l].fpart); a.fpart_cs = w.L[l].n_fpart;
        if (w.L[l].tc2) {
            a.prepped = prepped ? 1 : 0;
            a.wt = wsf(ws, w.L[l].wt);
            if (w.L[l].n_zt > 0) a.zt = wsf(ws, w.L[l].zt);
            if (w.L[l].n_at > 0) { a.at = wsf(ws, w.L[l].at); a.ot = wsf(ws, w.L[l].ot); }
        }
        int rc;
        if (w.L[l].tc_cols != 0 && dgprf_fwd_tc2_supported(a)) {                                                // pipelined
            if (l > 0 && a.Fprev.n_slabs > 1 && w.n_fsum > 0 && a.d <= 128) {      // (the WIDE variant's input split sums the slabs itself)
                // every CTA of a row block reads the whole input tile: the partial slabs of F_{l-1} are summed once (even two
                // slabs: adding them in the CTA's own prologue costs a second dependent load round per CTA -- measured +50 us
                // per layer at 65 536 rows) -- by the tail of the previous layer's kernel when that was the pipelined one, by
                // a launch otherwise
                float* fs = wsf(ws, w.fsum) + (int64_t)((l - 1) & 1) * m->n_chains * w.n_fsum;
                if (!fsum_ready) {
                    rc = dgprf_launch_sum_slabs(a.Fprev, B, y.d_prev, fs, w.n_fsum, m->n_chains, st);
                    if (rc) return rc;
                }
                a.Fprev.ptr = fs; a.Fprev.cs = w.n_fsum; a.Fprev.ss = 0; a.Fprev.n_slabs = 1;
            }
            fsum_ready = false;
            if (fuse_sums && a.CS > 1 && l + 1 < m->n_layers && w.L[l + 1].tc2 && w.n_fsum > 0 && layer_d(m->layer[l + 1]) <= 128 &&
                (int64_t)B * y.g <= w.n_fsum) {
                a.Fsum = wsf(ws, w.fsum) + (int64_t)(l & 1) * m->n_chains * w.n_fsum; a.fsum_cs = w.n_fsum;
                a.sum_ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + w.sumctr);
                fsum_ready = true;
            }
            rc = dgprf_launch_fwd_tc2(a, m->n_chains, st);
        }
        else {
            DGPRF_REQUIRE(!a.phi_blocked, "layer %d: blocked saved features without the pipelined forward", l);
            fsum_ready = false;
            rc = dgprf_launch_fwd_simt(a, m->n_chains, st);
        }
        if (rc) return rc;
    }
    if (F_out) {
        const int L = m->n_layers - 1;
        return dgprf_launch_sum_slabs(fpart_of(m, w, ws, L, B), B, m->d_out, F_out, (int64_t)B * m->d_out, m->n_chains, st);
    }
    return DGPRF_OK;
}

extern "C" int dgprf_forward(const dgprf_model* m, const float* X, int64_t x_cs, int B, int mode,
                             void* ws, size_t ws_bytes, float* F_out, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(X != nullptr, "X is NULL");
    return forward_impl(m, w, X, x_cs, B, mode, ws, F_out, (cudaStream_t)stream);
}

// ---- likelihood ----------------------------------------------------------------------------------
static int loglik_impl(const dgprf_model* m, const WsLayout& w, const float* Y, int64_t y_cs, int B, int mode,
                       void* ws, float* ll_rows, float* aux_rows, float* ll_sum, float inv_B, cudaStream_t st) {
    const int L = m->n_layers - 1;
    LikArgs a;
    memset(&a, 0, sizeof(a));
    a.likelihood = m->likelihood; a.B = B; a.D = m->d_out;
    a.F = fpart_of(m, w, ws, L, B);
    a.Y = Y; a.y_cs = y_cs;
    a.lik_log_var = m->likelihood == DGPRF_LIK_GAUSSIAN ? m->h_base + m->off_lik_log_var : nullptr;
    a.h_cs = m->h_cs;
    a.ll_rows = ll_rows; a.aux_rows = aux_rows;
    a.ll_sum = ll_sum ? ll_sum : wsf(ws, w.llsum);
    a.part = wsf(ws, w.likpart);
    a.inv_B = inv_B;
    if (inv_B > 0.f) {
        DGPRF_REQUIRE(mode >= DGPRF_MODE_TRAIN, "dU/dF needs a TRAIN/HYPER workspace");
        a.dF = wsf(ws, w.dflast); a.df_cs = w.n_dflast;
        if (mode == DGPRF_MODE_HYPER) {
            DGPRF_CHECK_CUDA(cudaMemsetAsync(wsf(ws, w.ghyp), 0, sizeof(float) * w.h_len * m->n_chains, st));
            if (m->likelihood == DGPRF_LIK_GAUSSIAN) {
                a.g_lik_log_var = wsf(ws, w.ghyp) + m->off_lik_log_var;
                a.g_cs = w.h_len;
            }
        }
    }
    return dgprf_launch_loglik(a, m->n_chains, st);
}

extern "C" int dgprf_loglik(const dgprf_model* m, const float* Y, int64_t y_cs, int B, int mode,
                            void* ws, size_t ws_bytes, float* ll_rows, float* aux_rows, float* ll_sum,
                            float inv_B, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(Y != nullptr, "Y is NULL");
    return loglik_impl(m, w, Y, y_cs, B, mode, ws, ll_rows, aux_rows, ll_sum, inv_B, (cudaStream_t)stream);
}

// ---- backward -------------------------------------------------------------------------------------
// Per-layer hook of the layered reverse pass (thread-local): called on the host right after the kernels of layer l have been
// enqueued, top layer first.  The data-parallel step uses it to start the all-reduce of a layer's gradient slice on a side
// stream while the layers below are still running (dgprf/dist.py: data_parallel_step).
static thread_local dgprf_layer_hook g_bwd_hook = nullptr;
static thread_local void* g_bwd_hook_user = nullptr;

extern "C" int dgprf_set_backward_hook(dgprf_layer_hook hook, void* user) {
    g_bwd_hook = hook;
    g_bwd_hook_user = hook ? user : nullptr;
    return DGPRF_OK;
}

static int backward_impl(const dgprf_model* m, const WsLayout& w, const float* X, int64_t x_cs, int B, int mode,
                         void* ws, cudaStream_t st) {
    const int hyper = mode == DGPRF_MODE_HYPER;
    // opt-in (DGPRF_FUSED_SLAB_SUMS=1), measured slower than the launches it removes: the ticket + sum are three to five DEPENDENT
    // round trips at a CTA's tail (~1.8 us each under the other CTAs' streams), a PDL-overlapped k_sum_slabs launch is 7-17 us
    // off the critical path of nothing (profiles/r02_fused_slab_sums.txt: configs[4] layer backward 523 -> 744 us, forward +8 us)
    const bool fuse_sums = getenv("DGPRF_FUSED_SLAB_SUMS") != nullptr;
    bool dsum_ready = false;                 // the backward kernel of layer l + 1 already summed its dF slabs into dfsum[(l + 1) & 1]
    for (int l = m->n_layers - 1; l >= 0; --l) {
        const dgprf_layer& y = m->layer[l];
        BwdArgs a;
        memset(&a, 0, sizeof(a));
        a.kind = y.kind; a.B = B; a.d_prev = y.d_prev; a.d_x = y.d_x; a.d = layer_d(y);
        a.M = y.M; a.g = y.g; a.F = layer_F(y); a.CS = w.L[l].CS; a.RS = w.RS; a.ldx = m->d_in;
        a.has_mean = y.has_mean; a.hyper = hyper;
        if (l == m->n_layers - 1) {
            a.dF.ptr = wsf(ws, w.dflast); a.dF.cs = w.n_dflast; a.dF.ss = 0; a.dF.ld = y.g; a.dF.n_slabs = 1;
        } else {
            a.dF.ptr = wsf(ws, w.L[l + 1].dpart); a.dF.cs = w.L[l + 1].n_dpart;
            a.dF.ss = (int64_t)B * y.g; a.dF.ld = y.g; a.dF.n_slabs = w.L[l + 1].CS;
        }
        a.Phi = wsf(ws, w.L[l].phi); a.phi_cs = w.L[l].n_phi; a.phi_blocked = w.L[l].phi_blocked;
        a.z = y.z; a.z_cs = y.z_cs;
        a.log_inv_ls = m->h_base + y.off_log_inv_ls;
        a.log_amp = m->h_base + y.off_log_amp;
        a.mean = y.has_mean ? m->h_base + y.off_mean : nullptr;
        a.h_cs = m->h_cs;
        a.W = m->w_base + y.off_W; a.w_cs = m->w_cs;
        a.gWpart = wsf(ws, w.gwpart) + y.off_W; a.gw_cs = w.n_gwpart; a.gw_ss = w.w_len;
        a.Dpart = l > 0 ? wsf(ws, w.L[l].dpart) : nullptr; a.d_cs = w.L[l].n_dpart;
        a.Tpart = hyper ? wsf(ws, w.L[l].tpart) : nullptr; a.t_cs = w.L[l].n_tpart;
        a.Rpart = hyper ? wsf(ws, w.L[l].rpart) : nullptr; a.r_cs = w.L[l].n_rpart;
        a.wp = w.L[l].bwd2 ? wsf(ws, w.L[l].wp) : nullptr;
        a.prepped = w.L[l].bwd2 ? 1 : 0;          // forward_impl of this step wrote wp (k_prep_layers)
        int rc;
        if (m->precision == DGPRF_PREC_TF32 && dgprf_bwd_tc2_supported(a)) {
            if (a.dF.n_slabs > 1) {
                // every column-split CTA reads the whole dF tile: the partial slabs are summed once instead of CS times -- by
                // the tail of the kernel of the layer above when that was the pipelined one, by a launch otherwise
                float* ds = wsf(ws, w.dfsum) + (int64_t)((l + 1) & 1) * m->n_chains * w.n_dfsum;
                if (!dsum_ready) {
                    rc = dgprf_launch_sum_slabs(a.dF, B, y.g, ds, w.n_dfsum, m->n_chains, st);
                    if (rc) return rc;
                }
                a.dF.ptr = ds; a.dF.cs = w.n_dfsum; a.dF.ss = 0; a.dF.n_slabs = 1;
            }
            dsum_ready = false;
            if (fuse_sums && l > 0 && a.CS > 1 && w.L[l - 1].bwd2 && w.n_dfsum >= (int64_t)B * y.d_prev) {
                a.Dsum = wsf(ws, w.dfsum) + (int64_t)(l & 1) * m->n_chains * w.n_dfsum; a.dsum_cs = w.n_dfsum;
                a.sum_ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + w.sumctr) + (int64_t)m->n_chains * w.n_sumctr;
                dsum_ready = true;
            }
            rc = dgprf_launch_bwd_tc2(a, m->n_chains, st);
        }
        else {
            DGPRF_REQUIRE(!a.phi_blocked, "layer %d: blocked saved features without the pipelined backward", l);
            dsum_ready = false;
            rc = dgprf_launch_bwd_simt(a, m->n_chains, st);
        }
        if (rc) return rc;
        if (hyper) {
            HypArgs h;
            memset(&h, 0, sizeof(h));
            h.B = B; h.d = a.d; h.d_prev = y.d_prev; h.d_x = y.d_x; h.g = y.g; h.ldx = m->d_in; h.has_mean = y.has_mean;
            if (l > 0) h.Fprev = fpart_of(m, w, ws, l - 1, B);
            h.X = X; h.x_cs = x_cs;
            h.T.ptr = a.Tpart; h.T.cs = a.t_cs; h.T.ss = (int64_t)B * a.d; h.T.ld = a.d; h.T.n_slabs = a.CS;
            h.R.ptr = a.Rpart; h.R.cs = a.r_cs; h.R.ss = B; h.R.ld = 1; h.R.n_slabs = a.CS;
            h.dF = a.dF;
            h.Fcur = fpart_of(m, w, ws, l, B);
            h.log_inv_ls = a.log_inv_ls; h.h_cs = m->h_cs;
            h.gH = wsf(ws, w.ghyp); h.gh_cs = w.h_len;
            h.part = wsf(ws, w.hpart); h.part_cs = w.n_hpart;
            h.off_log_amp = y.off_log_amp; h.off_log_inv_ls = y.off_log_inv_ls; h.off_mean = y.off_mean;
            rc = dgprf_launch_hyper_reduce(h, m->n_chains, st);
            if (rc) return rc;
        }
        if (g_bwd_hook) g_bwd_hook(l, g_bwd_hook_user);
    }
    return DGPRF_OK;
}

extern "C" int dgprf_backward(const dgprf_model* m, const float* X, int64_t x_cs, int B, int mode,
                              void* ws, size_t ws_bytes, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(mode >= DGPRF_MODE_TRAIN, "backward needs a TRAIN/HYPER workspace");
    DGPRF_REQUIRE(X != nullptr, "X is NULL");
    return backward_impl(m, w, X, x_cs, B, mode, ws, (cudaStream_t)stream);
}

extern "C" int dgprf_grad_finalize(const dgprf_model* m, int B, int mode, void* ws, size_t ws_bytes,
                                   float* gW, int64_t gw_cs, float* gH, int64_t gh_cs,
                                   float prior_inv_N, int prior_hyper, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(mode >= DGPRF_MODE_TRAIN, "grad_finalize needs a TRAIN/HYPER workspace");
    cudaStream_t st = (cudaStream_t)stream;
    if (gW) {
        DGPRF_REQUIRE(gw_cs >= w.w_len, "gW chain stride %lld < %lld", (long long)gw_cs, (long long)w.w_len);
        rc = dgprf_launch_grad_finalize(wsf(ws, w.gwpart), w.n_gwpart, w.w_len, w.RS,
                                        m->w_base, m->w_cs, prior_inv_N, gW, gw_cs, w.w_len, m->n_chains, st);
        if (rc) return rc;
    }
    if (gH) {
        DGPRF_REQUIRE(mode == DGPRF_MODE_HYPER, "hyper gradients need a HYPER workspace");
        DGPRF_REQUIRE(gh_cs >= w.h_len, "gH chain stride %lld < %lld", (long long)gh_cs, (long long)w.h_len);
        rc = dgprf_launch_grad_finalize(wsf(ws, w.ghyp), w.h_len, 0, 1, m->h_base, m->h_cs,
                                        prior_hyper ? prior_inv_N : 0.f, gH, gh_cs, w.h_len, m->n_chains, st);
        if (rc) return rc;
    }
    return DGPRF_OK;
}

extern "C" int dgprf_grad_finalize_layer(const dgprf_model* m, int layer, int B, int mode, void* ws, size_t ws_bytes,
                                         float* gW, int64_t gw_cs, float prior_inv_N, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(mode >= DGPRF_MODE_TRAIN, "grad_finalize_layer needs a TRAIN/HYPER workspace");
    DGPRF_REQUIRE(layer >= 0 && layer < m->n_layers, "layer %d out of range [0, %d)", layer, m->n_layers);
    DGPRF_REQUIRE(gW != nullptr, "gW is NULL");
    DGPRF_REQUIRE(gw_cs >= w.w_len, "gW chain stride %lld < %lld", (long long)gw_cs, (long long)w.w_len);
    const dgprf_layer& y = m->layer[layer];
    // the slice runs to the next tensor's offset (tensors are padded to 128-bit lanes; the whole-buffer finalize covers the
    // padding too, and the update kernel walks whole lanes)
    int64_t n = w.w_len - y.off_W;
    for (int l = 0; l < m->n_layers; ++l)
        if (m->layer[l].off_W > y.off_W && m->layer[l].off_W - y.off_W < n) n = m->layer[l].off_W - y.off_W;
    DGPRF_REQUIRE(n >= (int64_t)layer_F(y) * y.g, "layer %d: W slice overlaps the next tensor", layer);
    return dgprf_launch_grad_finalize(wsf(ws, w.gwpart) + y.off_W, w.n_gwpart, w.w_len, w.RS, m->w_base + y.off_W, m->w_cs,
                                      prior_inv_N, gW + y.off_W, gw_cs, n, m->n_chains, (cudaStream_t)stream);
}

extern "C" int dgprf_gradients(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y, int64_t y_cs, int B, int mode,
                               void* ws, size_t ws_bytes, float* gW, int64_t gw_cs, float* gH, int64_t gh_cs,
                               float prior_inv_N, int prior_hyper, float* ll_sum, float inv_B, int allow_fused, void* stream) {
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(mode >= DGPRF_MODE_TRAIN, "gradients need a TRAIN/HYPER workspace");
    DGPRF_REQUIRE(X && Y, "X/Y is NULL");
    DGPRF_REQUIRE(gH == nullptr || mode == DGPRF_MODE_HYPER, "hyper gradients need a HYPER workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const float invB_default = 1.f / (float)B;
    if (inv_B <= 0.f) inv_B = invB_default;
    int n_part = w.RS;
    if (allow_fused && mode == DGPRF_MODE_TRAIN && w.RSF > 0 && inv_B == invB_default) {
        bool fused = false;                              // (no update requested: the kernel stops after the gradient slabs)
        rc = (w.K10 ? dgprf_launch_step_cluster : dgprf_launch_step_rows)(
            m, X, x_cs, Y, y_cs, B, wsf(ws, w.gwpart), w.n_gwpart, w.w_len, wsf(ws, w.llpart), w.n_llpart, nullptr, nullptr, 0,
            reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + w.gridbar), nullptr, &fused, st);
        if (rc) return rc;
        if (ll_sum) {
            rc = dgprf_launch_sum_rows(wsf(ws, w.llpart), w.n_llpart, w.RSF, ll_sum, m->n_chains, st);
            if (rc) return rc;
        }
        n_part = w.RSF;
    } else {
        rc = forward_impl(m, w, X, x_cs, B, mode, ws, nullptr, st);
        if (rc) return rc;
        rc = loglik_impl(m, w, Y, y_cs, B, mode, ws, nullptr, nullptr, ll_sum, inv_B, st);
        if (rc) return rc;
        rc = backward_impl(m, w, X, x_cs, B, mode, ws, st);
        if (rc) return rc;
    }
    if (gW) {
        DGPRF_REQUIRE(gw_cs >= w.w_len, "gW chain stride %lld < %lld", (long long)gw_cs, (long long)w.w_len);
        if (prior_inv_N == 0.f) {
            // no prior term (the data-parallel step adds theta / N once, in the update): the plain slab sum, 128-bit lanes with
            // four slabs in flight per thread (same slab order as k_grad_finalize; 19 -> ~11 us for 18 slabs of 4 MB)
            SlabMat s;
            s.ptr = wsf(ws, w.gwpart); s.cs = w.n_gwpart; s.ss = w.w_len; s.ld = 1; s.n_slabs = n_part;
            rc = dgprf_launch_sum_slabs(s, (int)w.w_len, 1, gW, gw_cs, m->n_chains, st);
        } else
            rc = dgprf_launch_grad_finalize(wsf(ws, w.gwpart), w.n_gwpart, w.w_len, n_part, m->w_base, m->w_cs, prior_inv_N, gW, gw_cs,
                                            w.w_len, m->n_chains, st);
        if (rc) return rc;
    }
    if (gH) {
        DGPRF_REQUIRE(gh_cs >= w.h_len, "gH chain stride %lld < %lld", (long long)gh_cs, (long long)w.h_len);
        rc = dgprf_launch_grad_finalize(wsf(ws, w.ghyp), w.h_len, 0, 1, m->h_base, m->h_cs, prior_hyper ? prior_inv_N : 0.f, gH, gh_cs,
                                        w.h_len, m->n_chains, st);
        if (rc) return rc;
    }
    return DGPRF_OK;
}

// ---- update -----------------------------------------------------------------------------------------
// the device-resident step base of dgprf_sgmcmc_step_graph, attached to every UpdArgs built during that call
static thread_local const unsigned long long* g_step_dev = nullptr;

static int fill_upd_args(UpdArgs& a, float* theta, float* mom, int64_t cs, int64_t n,
                         const float* grad, int64_t grad_cs, int n_part, int64_t part_stride,
                         const dgprf_segment* segs, int n_seg, float lr, float data_size, float beta,
                         float temperature, int resample, uint64_t seed, uint64_t step, uint32_t stream_base,
                         const float* eps_inject, const float* mom_inject) {
    DGPRF_REQUIRE(theta && mom && grad && segs, "update: NULL buffer");
    DGPRF_REQUIRE(lr > 0.f && data_size > 0.f && beta >= 0.f && beta < 1.f && temperature >= 0.f,
                  "update: need lr>0, N>0, 0<=beta<1, T>=0");
    memset(&a, 0, sizeof(a));
    a.theta = theta; a.mom = mom; a.cs = cs; a.n = n;
    a.grad = grad; a.grad_cs = grad_cs; a.n_part = n_part; a.part_stride = part_stride; a.n_seg = n_seg;
    a.h = sqrtf(lr / data_size); a.hN = a.h * data_size; a.beta = beta;
    a.noise_scale = sqrtf(2.f * (1.f - beta) * temperature);
    a.inv_N = 1.f / data_size;
    a.resample = resample; a.seed = seed; a.step = step; a.stream_base = stream_base;
    a.eps_inject = eps_inject; a.mom_inject = mom_inject;
    a.step_dev = g_step_dev;
    return DGPRF_OK;
}

static int update_impl(float* theta, float* mom, int64_t cs, int64_t n, int n_chains,
                       const float* grad, int64_t grad_cs, int n_part, int64_t part_stride,
                       const dgprf_segment* segs, int n_seg, float lr, float data_size, float beta,
                       float temperature, int resample, uint64_t seed, uint64_t step, uint32_t stream_base,
                       const float* eps_inject, const float* mom_inject, cudaStream_t st) {
    UpdArgs a;
    const int rc = fill_upd_args(a, theta, mom, cs, n, grad, grad_cs, n_part, part_stride, segs, n_seg, lr, data_size,
                                 beta, temperature, resample, seed, step, stream_base, eps_inject, mom_inject);
    if (rc) return rc;
    return dgprf_launch_update(a, segs, n_seg, n_chains, st);
}

extern "C" int dgprf_sgmcmc_update(float* theta, float* mom, int64_t cs, int64_t n, int n_chains,
                                   const float* grad, int64_t grad_cs, int n_part, int64_t part_stride,
                                   const dgprf_segment* segs, int n_seg,
                                   float lr, float data_size, float momentum_decay, float temperature,
                                   int resample_moments, uint64_t seed, uint64_t step,
                                   const float* eps_inject, const float* mom_inject, void* stream) {
    DGPRF_REQUIRE(n_part >= 1 && n_chains >= 1, "update: n_part/n_chains must be >= 1");
    return update_impl(theta, mom, cs, n, n_chains, grad, grad_cs, n_part, part_stride, segs, n_seg, lr, data_size,
                       momentum_decay, temperature, resample_moments, seed, step, 0u, eps_inject, mom_inject,
                       (cudaStream_t)stream);
}

extern "C" int dgprf_sgmcmc_step(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y,
                                 int64_t y_cs, int B, int full_bayesian,
                                 float* theta_w, float* mom_w, int64_t w_len,
                                 const dgprf_segment* segs_w, int n_seg_w,
                                 float* theta_h, float* mom_h, int64_t h_len,
                                 const dgprf_segment* segs_h, int n_seg_h,
                                 float lr, float data_size, float momentum_decay, float temperature,
                                 int resample_moments, uint64_t seed, uint64_t step,
                                 const float* eps_w, const float* res_w, const float* eps_h, const float* res_h,
                                 void* ws, size_t ws_bytes, float* u_out, void* stream);

extern "C" int dgprf_sgmcmc_step_graph(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y,
                                       int64_t y_cs, int B, int full_bayesian,
                                       float* theta_w, float* mom_w, int64_t w_len,
                                       const dgprf_segment* segs_w, int n_seg_w,
                                       float* theta_h, float* mom_h, int64_t h_len,
                                       const dgprf_segment* segs_h, int n_seg_h,
                                       float lr, float data_size, float momentum_decay, float temperature,
                                       int resample_moments, uint64_t seed, uint64_t step, const uint64_t* step_base_dev,
                                       void* ws, size_t ws_bytes, float* u_out, void* stream) {
    DGPRF_REQUIRE(step_base_dev != nullptr, "step_graph: step_base_dev is NULL");
    g_step_dev = reinterpret_cast<const unsigned long long*>(step_base_dev);
    const int rc = dgprf_sgmcmc_step(m, X, x_cs, Y, y_cs, B, full_bayesian, theta_w, mom_w, w_len, segs_w, n_seg_w, theta_h, mom_h,
                                     h_len, segs_h, n_seg_h, lr, data_size, momentum_decay, temperature, resample_moments, seed,
                                     step, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes, u_out, stream);
    g_step_dev = nullptr;
    return rc;
}

extern "C" int dgprf_sgmcmc_step(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y,
                                 int64_t y_cs, int B, int full_bayesian,
                                 float* theta_w, float* mom_w, int64_t w_len,
                                 const dgprf_segment* segs_w, int n_seg_w,
                                 float* theta_h, float* mom_h, int64_t h_len,
                                 const dgprf_segment* segs_h, int n_seg_h,
                                 float lr, float data_size, float momentum_decay, float temperature,
                                 int resample_moments, uint64_t seed, uint64_t step,
                                 const float* eps_w, const float* res_w, const float* eps_h, const float* res_h,
                                 void* ws, size_t ws_bytes, float* u_out, void* stream) {
    const int mode = full_bayesian ? DGPRF_MODE_HYPER : DGPRF_MODE_TRAIN;
    WsLayout w;
    int rc = check_ws(m, B, mode, ws, ws_bytes, &w);
    if (rc) return rc;
    DGPRF_REQUIRE(X && Y, "X/Y is NULL");
    DGPRF_REQUIRE(theta_w == m->w_base, "theta_w must alias model.w_base");
    DGPRF_REQUIRE(w_len == w.w_len, "w_len=%lld, layout expects %lld", (long long)w_len, (long long)w.w_len);
    DGPRF_REQUIRE(m->w_cs >= w_len || m->n_chains == 1, "w chain stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    int n_part = w.RS;
    if (!full_bayesian && w.RSF > 0) {
        // small minibatch: whole forward + likelihood seed + backward in one row-fused kernel, and (when
        // every CTA is co-resident) the update behind a grid barrier in the same launch
        UpdArgs ua;
        rc = fill_upd_args(ua, theta_w, mom_w, m->w_cs, w_len, wsf(ws, w.gwpart), w.n_gwpart, w.RSF, w.w_len, segs_w,
                           n_seg_w, lr, data_size, momentum_decay, temperature, resample_moments, seed, step, 0u,
                           eps_w, res_w);
        if (rc) return rc;
        ua.n_seg = n_seg_w;
        bool fused = false;
        rc = (w.K10 ? dgprf_launch_step_cluster : dgprf_launch_step_rows)(
            m, X, x_cs, Y, y_cs, B, wsf(ws, w.gwpart), w.n_gwpart, w.w_len, wsf(ws, w.llpart), w.n_llpart, &ua, segs_w,
            n_seg_w, reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + w.gridbar), u_out, &fused, st);
        if (rc) return rc;
        if (fused) return DGPRF_OK;
        if (u_out) {
            rc = dgprf_launch_sum_rows(wsf(ws, w.llpart), w.n_llpart, w.RSF, u_out, m->n_chains, st);
            if (rc) return rc;
        }
        n_part = w.RSF;
    } else {
        rc = forward_impl(m, w, X, x_cs, B, mode, ws, nullptr, st);
        if (rc) return rc;
        rc = loglik_impl(m, w, Y, y_cs, B, mode, ws, nullptr, nullptr, u_out, 1.f / (float)B, st);
        if (rc) return rc;
        rc = backward_impl(m, w, X, x_cs, B, mode, ws, st);
        if (rc) return rc;
    }
    rc = update_impl(theta_w, mom_w, m->w_cs, w_len, m->n_chains, wsf(ws, w.gwpart), w.n_gwpart, n_part,
                     w.w_len, segs_w, n_seg_w, lr, data_size, momentum_decay, temperature, resample_moments, seed,
                     step, 0u, eps_w, res_w, st);
    if (rc) return rc;
    if (full_bayesian && n_seg_h > 0) {
        DGPRF_REQUIRE(theta_h == m->h_base, "theta_h must alias model.h_base");
        DGPRF_REQUIRE(h_len == w.h_len, "h_len=%lld, layout expects %lld", (long long)h_len, (long long)w.h_len);
        rc = update_impl(theta_h, mom_h, m->h_cs, h_len, m->n_chains, wsf(ws, w.ghyp), w.h_len, 1, 0, segs_h,
                         n_seg_h, lr, data_size, momentum_decay, temperature, resample_moments, seed, step, 2u,
                         eps_h, res_h, st);
        if (rc) return rc;
    }
    return DGPRF_OK;
}

extern "C" int dgprf_sgmcmc_step_host(const dgprf_model* m, const float* X_host, const float* Y_host, int y_cols, int B,
                                      float* X_dev, float* Y_dev, int zero_copy, int full_bayesian,
                                      float* theta_w, float* mom_w, int64_t w_len,
                                      const dgprf_segment* segs_w, int n_seg_w,
                                      float* theta_h, float* mom_h, int64_t h_len,
                                      const dgprf_segment* segs_h, int n_seg_h,
                                      float lr, float data_size, float momentum_decay, float temperature,
                                      int resample_moments, uint64_t seed, uint64_t step,
                                      void* ws, size_t ws_bytes, float* u_dev, float* u_host, void* stream) {
    DGPRF_REQUIRE(m && X_host && Y_host && X_dev && Y_dev && B >= 1 && y_cols >= 1, "step_host: bad arguments");
    DGPRF_REQUIRE(u_host == nullptr || u_dev != nullptr, "step_host: u_host needs the u_dev staging word");
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_copy == 1)  // pinned host buffers are device-visible: no staging copies, the kernels read them over the bus
        return dgprf_sgmcmc_step(m, X_host, 0, Y_host, 0, B, full_bayesian, theta_w, mom_w, w_len, segs_w, n_seg_w,
                                 theta_h, mom_h, h_len, segs_h, n_seg_h, lr, data_size, momentum_decay, temperature,
                                 resample_moments, seed, step, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes,
                                 u_host, stream);
    const size_t xb = sizeof(float) * (size_t)B * m->d_in, yb = sizeof(float) * (size_t)B * y_cols;
    if (zero_copy == 2) {
        // Pipelined staging: the minibatch of step n+1 is copied on a side stream into the other half of the (double-sized)
        // staging buffers while the kernels of step n run; the step's kernels wait on the copy's event, and the copy that
        // re-uses a half waits on the event recorded after the step that read it.  The side stream and the four events are
        // the only resources this library ever creates (once per thread and device, never freed).
        struct Pipe { cudaStream_t cs; cudaEvent_t copied[2], freed[2]; int dev; unsigned n; bool ok; };
        static thread_local Pipe pl = {nullptr, {nullptr, nullptr}, {nullptr, nullptr}, -1, 0u, false};
        int dev = 0;
        DGPRF_CHECK_CUDA(cudaGetDevice(&dev));
        if (!pl.ok || pl.dev != dev) {
            DGPRF_CHECK_CUDA(cudaStreamCreateWithFlags(&pl.cs, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                DGPRF_CHECK_CUDA(cudaEventCreateWithFlags(&pl.copied[i], cudaEventDisableTiming));
                DGPRF_CHECK_CUDA(cudaEventCreateWithFlags(&pl.freed[i], cudaEventDisableTiming));
            }
            pl.dev = dev; pl.n = 0; pl.ok = true;
        }
        const int slot = (int)(pl.n++ & 1u);
        float* Xs = X_dev + (size_t)slot * B * m->d_in;
        float* Ys = Y_dev + (size_t)slot * B * y_cols;
        DGPRF_CHECK_CUDA(cudaStreamWaitEvent(pl.cs, pl.freed[slot], 0));        // a never-recorded event counts as complete
        DGPRF_CHECK_CUDA(cudaMemcpyAsync(Xs, X_host, xb, cudaMemcpyHostToDevice, pl.cs));
        DGPRF_CHECK_CUDA(cudaMemcpyAsync(Ys, Y_host, yb, cudaMemcpyHostToDevice, pl.cs));
        DGPRF_CHECK_CUDA(cudaEventRecord(pl.copied[slot], pl.cs));
        DGPRF_CHECK_CUDA(cudaStreamWaitEvent(st, pl.copied[slot], 0));
        const int rc = dgprf_sgmcmc_step(m, Xs, 0, Ys, 0, B, full_bayesian, theta_w, mom_w, w_len, segs_w, n_seg_w,
                                         theta_h, mom_h, h_len, segs_h, n_seg_h, lr, data_size, momentum_decay, temperature,
                                         resample_moments, seed, step, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes,
                                         u_host /* pinned: written in place by the kernel */, stream);
        if (rc) return rc;
        DGPRF_CHECK_CUDA(cudaEventRecord(pl.freed[slot], st));
        return DGPRF_OK;
    }
    DGPRF_CHECK_CUDA(cudaMemcpyAsync(X_dev, X_host, xb, cudaMemcpyHostToDevice, st));
    DGPRF_CHECK_CUDA(cudaMemcpyAsync(Y_dev, Y_host, yb, cudaMemcpyHostToDevice, st));
    const int rc = dgprf_sgmcmc_step(m, X_dev, 0, Y_dev, 0, B, full_bayesian, theta_w, mom_w, w_len, segs_w, n_seg_w,
                                     theta_h, mom_h, h_len, segs_h, n_seg_h, lr, data_size, momentum_decay, temperature,
                                     resample_moments, seed, step, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes,
                                     u_host ? u_dev : nullptr, stream);
    if (rc) return rc;
    if (u_host) DGPRF_CHECK_CUDA(cudaMemcpyAsync(u_host, u_dev, sizeof(float) * m->n_chains, cudaMemcpyDeviceToHost, st));
    return DGPRF_OK;
}

// ---- stand-alone layer ops -----------------------------------------------------------------------------
extern "C" int dgprf_rf_features(int kind, const float* X, int B, int d, const float* z, const float* log_inv_ls,
                                 const float* log_amp, const float* mean, int M, float* Phi, void* stream) {
    DGPRF_REQUIRE(kind == DGPRF_KIND_RBF || kind == DGPRF_KIND_ARC, "unknown kind %d", kind);
    DGPRF_REQUIRE(X && z && log_inv_ls && log_amp && Phi && B >= 0 && d >= 1 && M >= 1, "rf_features: bad arguments");
    if (B == 0) return DGPRF_OK;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.kind = kind; a.B = B; a.d_prev = 0; a.d_x = d; a.d = d; a.M = M; a.g = 1;
    a.F = kind == DGPRF_KIND_RBF ? 2 * M : M; a.CS = col_splits(M); a.ldx = d; a.do_gemm2 = 0;
    a.has_mean = mean != nullptr;
    a.X = X; a.z = z; a.log_inv_ls = log_inv_ls; a.log_amp = log_amp; a.mean = mean;
    a.W = nullptr; a.Phi = Phi;
    return dgprf_launch_fwd_simt(a, 1, (cudaStream_t)stream);
}

extern "C" int dgprf_gaussian_log_prob(const float* F, const float* Y, const float* lik_log_var, int B, int D,
                                       float* out_rows, void* stream) {
    DGPRF_REQUIRE(F && Y && lik_log_var && out_rows && B >= 0 && D >= 1, "gaussian_log_prob: bad arguments");
    if (B == 0) return DGPRF_OK;
    LikArgs a;
    memset(&a, 0, sizeof(a));
    a.likelihood = DGPRF_LIK_GAUSSIAN; a.B = B; a.D = D;
    a.F.ptr = F; a.F.ld = D; a.F.n_slabs = 1;
    a.Y = Y; a.lik_log_var = lik_log_var; a.ll_rows = out_rows;
    return dgprf_launch_loglik(a, 1, (cudaStream_t)stream);
}

extern "C" int dgprf_softmax_log_prob(const float* F, const float* Y, int B, int C, float* out_rows,
                                      float* probs, void* stream) {
    DGPRF_REQUIRE(F && (Y || !out_rows) && B >= 0 && C >= 1, "softmax_log_prob: bad arguments");
    if (B == 0) return DGPRF_OK;
    LikArgs a;
    memset(&a, 0, sizeof(a));
    a.likelihood = DGPRF_LIK_SOFTMAX; a.B = B; a.D = C;
    a.F.ptr = F; a.F.ld = C; a.F.n_slabs = 1;
    a.Y = Y; a.ll_rows = out_rows; a.probs = probs;
    return dgprf_launch_loglik(a, 1, (cudaStream_t)stream);
}
