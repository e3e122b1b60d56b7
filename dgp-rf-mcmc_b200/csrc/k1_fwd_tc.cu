// K1 (tensor-core variant, DGPRF_PREC_TF32): fused [RF layer -> GP layer] forward on tcgen05.
//
//   GEMM #1   P[128 x 64]   = (in * exp(log_inv_ls)) @ z-tile     tcgen05.mma kind::tf32, D in TMEM,
//             3xTF32 split (a = a_hi + a_lo: hi*hi + lo*hi + hi*lo) so P is fp32-accurate: the K of this
//             GEMM is tiny, and tf32-rounded phases would flip ReLU gates / shift cos,sin by ~1e-3
//   epilogue  Phi = amp/sqrt(M) [cos P, sin P] | sqrt(2) amp/sqrt(M) relu(P)   (tcgen05.ld -> registers;
//             Cody-Waite + MUFU sincos) written to shared memory in the K-major SWIZZLE_128B layout
//   TMA       the same shared tile is stored to the saved-feature matrix with cp.async.bulk.tensor
//   GEMM #2   F[128 x g]   += Phi-tile @ W-tile                   tcgen05.mma reading Phi from smem,
//             accumulated in TMEM over all column tiles of the CTA
//
// P never leaves the SM.  exp(log_inv_ls) is folded into the A operand (P = (in*s) z + in.mean), so the
// B operand of GEMM #1 is the raw fixed draw z.  One CTA = 128 batch rows x every CS-th 64-column tile.
// References: layers/rf_layers.py:36-45, 82-91; layers/GP_weight_layers.py:13; utils.py:42.
#include "kernels.cuh"
#include "tc_common.cuh"

constexpr int TC_BM = 128;       // UMMA M
constexpr int TC_KG = 2;         // 32-wide K blocks of GEMM #1 staged per group (hi and lo copies of each)
constexpr int TC_THREADS = 256;
constexpr int TC_A_BLK = TC_BM * 128;   // bytes of one [128 x 32 tf32] block
constexpr uint32_t TC_TMEM_COLS = 128;  // D1: BN (<= 64) columns, D2: up to 64 columns

// NG: padded n_gp (UMMA N of GEMM #2).  BN: random-feature columns per tile (UMMA N of GEMM #1):
// 64 for large problems, 32 when the grid would otherwise leave most SMs idle.
template <int NG, int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k1_fwd_tc(const FwdArgs a, const __grid_constant__ CUtensorMap map_cos, const __grid_constant__ CUtensorMap map_sin) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by offsetting the shared array itself: the pointer keeps its shared-memory address
    // space (a round trip through uintptr_t makes every access a generic LD/ST)
    uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA1 = sm;                                 // [hi blocks 0..KG) | lo blocks 0..KG)]
    constexpr int TC_BN = BN;
    constexpr int TC_B_BLK = BN * 128;
    constexpr int NPB = BN / 32;                       // 32-wide Phi blocks per half (cos | sin)
    uint8_t* sB1 = sA1 + 2 * TC_KG * TC_A_BLK;         // same split
    uint8_t* sPhi = sB1 + 2 * TC_KG * TC_B_BLK;        // 2*NPB blocks: cos 0..NPB) | sin NPB..2NPB)  (ARC: cos only)
    uint8_t* sW = sPhi + 2 * NPB * TC_A_BLK;           // 2*NPB blocks of [NG x 128 B]
    float* bias_s = reinterpret_cast<float*>(sW + 2 * NPB * NG * 128);
    uint64_t* bar = reinterpret_cast<uint64_t*>(bias_s + TC_BM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.z, cs = blockIdx.y, row0 = blockIdx.x * TC_BM;

    const float* z = a.z + chain * a.z_cs;
    const float* X = a.X + chain * a.x_cs;
    const float* W = a.W + chain * a.w_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float amp = __expf(__ldg(a.log_amp + chain * a.h_cs));
    const float scale = (rbf ? 1.f : 1.41421356237f) * amp * rsqrtf((float)a.M);

    if (warp == 0) tc::tmem_alloc(tmem_slot, TC_TMEM_COLS);
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + TC_BN;
    uint32_t phase = 0;

    constexpr uint32_t IDESC1 = tc::make_idesc_tf32(TC_BM, TC_BN);
    constexpr uint32_t IDESC2 = tc::make_idesc_tf32(TC_BM, NG);
    const int n_kb = (a.d + 31) / 32;                    // 32-wide K blocks of GEMM #1
    const int n_kg = (n_kb + TC_KG - 1) / TC_KG;
    const bool a_resident = n_kg == 1;                   // the whole A operand stays in shared memory
    const int nb2 = rbf ? 2 * NPB : NPB;                 // 32-wide K blocks of GEMM #2 per column tile
    bool d2_started = false;

    const int n_ct = (a.M + TC_BN - 1) / TC_BN;
    for (int ct = cs; ct < n_ct; ct += a.CS) {
        const int c0 = ct * TC_BN;
        if (tid == 0) tc::tma_wait_read0();              // previous tile's Phi store has drained sPhi
        float bsum[TC_BM / 8];
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) bsum[i] = 0.f;

        for (int kg = 0; kg < n_kg; ++kg) {
            const int nblk = min(TC_KG, n_kb - kg * TC_KG);
            // ---- stage A = tf32(in * s): lanes along K (coalesced rows of X / F_prev) ----
            if (!a_resident || ct == cs) {
                float v[TC_KG][TC_BM / 8], sq[TC_KG], mq[TC_KG];
#pragma unroll
                for (int kbl = 0; kbl < TC_KG; ++kbl) {          // every load of the group is issued first
                    const int q = (kg * TC_KG + kbl) * 32 + lane;
                    const bool qok = kbl < nblk && q < a.d;
                    sq[kbl] = qok ? expf(__ldg(ls + q)) : 0.f;
                    mq[kbl] = (qok && mean) ? __ldg(mean + q) : 0.f;
#pragma unroll
                    for (int rr = 0; rr < TC_BM / 8; ++rr) {
                        const int64_t row = row0 + warp + 8 * rr;
                        float t = 0.f;
                        if (row < a.B && qok)
                            t = q < a.d_prev ? slab_load(a.Fprev, chain, row, q) : __ldg(X + row * a.ldx + (q - a.d_prev));
                        v[kbl][rr] = t;
                    }
                }
#pragma unroll
                for (int kbl = 0; kbl < TC_KG; ++kbl) {
                    if (kbl < nblk) {
#pragma unroll
                        for (int rr = 0; rr < TC_BM / 8; ++rr) {
                            const int r = warp + 8 * rr;
                            const float x = v[kbl][rr] * sq[kbl], hi = tc::to_tf32(x);
                            *reinterpret_cast<float*>(sA1 + kbl * TC_A_BLK + tc::sw128_off(r, lane)) = hi;
                            *reinterpret_cast<float*>(sA1 + (TC_KG + kbl) * TC_A_BLK + tc::sw128_off(r, lane)) = tc::to_tf32(x - hi);
                            bsum[rr] = fmaf(v[kbl][rr], mq[kbl], bsum[rr]);
                        }
                    }
                }
            }
            // ---- stage B = tf32(z tile), rows = feature columns, 16-byte chunks along K ----
            {
                constexpr int NU = TC_KG * TC_BN * 8 / TC_THREADS;        // (block, row, chunk) items per thread
                float zx[NU][4];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int e = tid + u * TC_THREADS;
                    const int n = e % TC_BN, ch = (e / TC_BN) & 7, kbl = e / (TC_BN * 8);
                    const int col = c0 + n;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int q = (kg * TC_KG + kbl) * 32 + ch * 4 + i;
                        zx[u][i] = (kbl < nblk && q < a.d && col < a.M) ? __ldg(z + (int64_t)q * a.M + col) : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int e = tid + u * TC_THREADS;
                    const int n = e % TC_BN, ch = (e / TC_BN) & 7, kbl = e / (TC_BN * 8);
                    if (kbl < nblk) {
                        float4 o, ol;
                        o.x = tc::to_tf32(zx[u][0]); o.y = tc::to_tf32(zx[u][1]); o.z = tc::to_tf32(zx[u][2]); o.w = tc::to_tf32(zx[u][3]);
                        ol.x = tc::to_tf32(zx[u][0] - o.x); ol.y = tc::to_tf32(zx[u][1] - o.y);
                        ol.z = tc::to_tf32(zx[u][2] - o.z); ol.w = tc::to_tf32(zx[u][3] - o.w);
                        *reinterpret_cast<float4*>(sB1 + kbl * TC_B_BLK + tc::sw128_chunk(n, ch)) = o;
                        *reinterpret_cast<float4*>(sB1 + (TC_KG + kbl) * TC_B_BLK + tc::sw128_chunk(n, ch)) = ol;
                    }
                }
            }
            // ---- stage the W tile of GEMM #2 once per column tile ----
            if (kg == 0 && a.do_gemm2) {
                constexpr int NW = 2 * BN * NG / TC_THREADS;             // W elements per thread (both halves)
                float wv[NW];
#pragma unroll
                for (int u = 0; u < NW; ++u) {
                    const int e = tid + u * TC_THREADS;
                    const int j = e % NG, kf = e / NG;                  // kf in [0, 2*BN): cos half | sin half
                    const int col = c0 + (kf % BN);
                    const int64_t frow = (kf >= BN ? a.M : 0) + col;
                    wv[u] = (j < a.g && col < a.M && (rbf || kf < BN)) ? __ldg(W + frow * a.g + j) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < NW; ++u) {
                    const int e = tid + u * TC_THREADS;
                    const int j = e % NG, kf = e / NG;
                    if (rbf || kf < BN)
                        *reinterpret_cast<float*>(sW + (kf >> 5) * (NG * 128) + tc::sw128_off(j, kf & 31)) = tc::to_tf32(wv[u]);
                }
            }
            if (a.has_mean && kg == n_kg - 1 && (!a_resident || ct == cs)) {
#pragma unroll
                for (int rr = 0; rr < TC_BM / 8; ++rr) {
                    const float b = warp_sum(bsum[rr]);
                    if (lane == 0) bias_s[warp + 8 * rr] = b;
                }
            }
            tc::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc::tc_fence_after();
                for (int kbl = 0; kbl < nblk; ++kbl) {
                    const int kleft = a.d - (kg * TC_KG + kbl) * 32;
                    const int ksteps = kleft >= 32 ? 4 : (kleft + 7) / 8;
                    for (int k4 = 0; k4 < ksteps; ++k4) {
                        const uint64_t da = tc::make_desc_sw128(tc::smem_u32(sA1 + kbl * TC_A_BLK) + k4 * 32);
                        const uint64_t db = tc::make_desc_sw128(tc::smem_u32(sB1 + kbl * TC_B_BLK) + k4 * 32);
                        const uint64_t dal = tc::make_desc_sw128(tc::smem_u32(sA1 + (TC_KG + kbl) * TC_A_BLK) + k4 * 32);
                        const uint64_t dbl = tc::make_desc_sw128(tc::smem_u32(sB1 + (TC_KG + kbl) * TC_B_BLK) + k4 * 32);
                        tc::umma_tf32(tmem_d1, dal, db, IDESC1, (kg | kbl | k4) != 0);     // small terms first
                        tc::umma_tf32(tmem_d1, da, dbl, IDESC1, 1u);
                        tc::umma_tf32(tmem_d1, da, db, IDESC1, 1u);
                    }
                }
                tc::umma_commit(bar);
            }
            tc::mbar_wait(bar, phase);
            phase ^= 1;
            tc::tc_fence_after();
        }

        // ---- epilogue: TMEM -> registers -> activation -> swizzled shared tile ----
        {
            constexpr int CW = BN / 2;                           // columns per warp: 32 (BN=64) | 16 (BN=32)
            const int lq = warp & 3, chh = warp >> 2;            // TMEM lane quarter, column half
            const int r = 32 * lq + lane;
            float p[CW];
            if constexpr (CW == 32) tc::tmem_ld32(tmem_d1 + ((uint32_t)(32 * lq) << 16) + CW * chh, p);
            else tc::tmem_ld16(tmem_d1 + ((uint32_t)(32 * lq) << 16) + CW * chh, p);
            tc::tmem_ld_wait();
            const float bias = a.has_mean ? bias_s[r] : 0.f;
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                float4 f0, f1;
                float* f0p = reinterpret_cast<float*>(&f0);
                float* f1p = reinterpret_cast<float*>(&f1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool live = (c0 + CW * chh + 4 * c4 + i) < a.M;
                    const float x = p[4 * c4 + i] + bias;
                    if (rbf) {
                        float s, c;
                        sincos_cw(x, &s, &c);
                        f0p[i] = live ? scale * c : 0.f;
                        f1p[i] = live ? scale * s : 0.f;
                    } else {
                        f0p[i] = live ? scale * fmaxf(x, 0.f) : 0.f;
                        f1p[i] = 0.f;
                    }
                }
                // column CW*chh + 4*c4 of the tile -> 32-wide block and 16-byte chunk inside it
                const int cb = (CW * chh + 4 * c4) >> 5, cc = ((CW * chh + 4 * c4) & 31) >> 2;
                *reinterpret_cast<float4*>(sPhi + cb * TC_A_BLK + tc::sw128_chunk(r, cc)) = f0;
                if (rbf) *reinterpret_cast<float4*>(sPhi + (NPB + cb) * TC_A_BLK + tc::sw128_chunk(r, cc)) = f1;
            }
        }
        tc::tc_fence_before();
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            if (a.do_gemm2) {
                for (int b = 0; b < nb2; ++b)
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t da = tc::make_desc_sw128(tc::smem_u32(sPhi + b * TC_A_BLK) + k4 * 32);
                        const uint64_t db = tc::make_desc_sw128(tc::smem_u32(sW + b * (NG * 128)) + k4 * 32);
                        tc::umma_tf32(tmem_d2, da, db, IDESC2, (d2_started || (b | k4) != 0) ? 1u : 0u);
                    }
                tc::umma_commit(bar);
            }
            if (a.Phi != nullptr) {                              // saved features for the backward
                for (int b = 0; b < NPB; ++b) {
                    if (c0 + 32 * b >= a.M) break;
                    tc::tma_store_3d(&map_cos, tc::smem_u32(sPhi + b * TC_A_BLK), c0 + 32 * b, row0, chain);
                    if (rbf) tc::tma_store_3d(&map_sin, tc::smem_u32(sPhi + (NPB + b) * TC_A_BLK), c0 + 32 * b, row0, chain);
                }
                tc::tma_commit();
            }
        }
        if (a.do_gemm2) {
            tc::mbar_wait(bar, phase);
            phase ^= 1;
            tc::tc_fence_after();
            d2_started = true;
        }
    }

    // ---- final epilogue: the F partial slab of this column split ----
    if (a.do_gemm2 && warp < 4) {
        const int r = 32 * warp + lane;
        const int64_t row = row0 + r;
#pragma unroll
        for (int c16 = 0; c16 < NG / 16; ++c16) {
            float v[16];
            tc::tmem_ld16(tmem_d2 + ((uint32_t)(32 * warp) << 16) + 16 * c16, v);
            tc::tmem_ld_wait();
            if (row < a.B) {
                float* dst = a.Fpart + chain * a.fpart_cs + ((int64_t)cs * a.B + row) * a.g;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (16 * c16 + j < a.g) dst[16 * c16 + j] = d2_started ? v[j] : 0.f;
            }
        }
    }
    if (tid == 0) tc::tma_wait0();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

static size_t tc_fwd_smem_bytes(int NG, int BN) {
    return 1024 + 2 * (size_t)TC_KG * TC_A_BLK + 2 * (size_t)TC_KG * BN * 128 + 2 * (size_t)(BN / 32) * TC_A_BLK +
           2 * (size_t)(BN / 32) * NG * 128 + TC_BM * sizeof(float) + 16;
}

template <int NG, int BN>
static int launch_fwd_tc(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const size_t smem = tc_fwd_smem_bytes(NG, BN);
    { const int rc_s = dgprf_ensure_smem((const void*)k1_fwd_tc<NG, BN>, (size_t)smem); if (rc_s) return rc_s; }
    CUtensorMap mc, ms;
    memset(&mc, 0, sizeof(mc));
    memset(&ms, 0, sizeof(ms));
    if (a.Phi != nullptr) {
        int rc = dgprf_make_tmap_3d(&mc, a.Phi, a.M, a.B, n_chains, a.F, a.phi_cs, TC_BM);
        if (rc) return rc;
        if (a.kind == DGPRF_KIND_RBF) {
            rc = dgprf_make_tmap_3d(&ms, a.Phi + a.M, a.M, a.B, n_chains, a.F, a.phi_cs, TC_BM);
            if (rc) return rc;
        }
    }
    dim3 grid(ceil_div(a.B, TC_BM), a.CS, n_chains);
    { ProfScope _ps("k1_fwd_tc", st); k1_fwd_tc<NG, BN><<<grid, TC_THREADS, smem, st>>>(a, mc, ms); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// true when the tensor-core kernel supports this layer shape (else the caller uses the SIMT kernel)
bool dgprf_fwd_tc_supported(const FwdArgs& a) {
    return (a.M % 4 == 0) && a.g <= 64 && (a.Phi == nullptr || (a.phi_cs % 4) == 0);
}

// Column-tile width of the tensor-core kernels for a layer: 32 when 64-wide tiles would leave most of the
// 148 SMs without a CTA (small minibatches), else 64.  The workspace layout (api.cu) uses the same rule.
int dgprf_tc_tile_cols(int B, int M, int n_chains) {
    const int64_t ctas64 = (int64_t)ceil_div(B, TC_BM) * (ceil_div(M, 64) < kMaxCS ? ceil_div(M, 64) : kMaxCS) * n_chains;
    return ctas64 < 120 ? 32 : 64;
}

int dgprf_launch_fwd_tc(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const int g = a.do_gemm2 ? a.g : 1;
    if (a.tile_cols == 32) {
        if (g <= 16) return launch_fwd_tc<16, 32>(a, n_chains, st);
        if (g <= 32) return launch_fwd_tc<32, 32>(a, n_chains, st);
        return launch_fwd_tc<64, 32>(a, n_chains, st);
    }
    if (g <= 16) return launch_fwd_tc<16, 64>(a, n_chains, st);
    if (g <= 32) return launch_fwd_tc<32, 64>(a, n_chains, st);
    return launch_fwd_tc<64, 64>(a, n_chains, st);
}
