// K1 v2 (tensor-core, DGPRF_PREC_TF32): warp-specialised, pipelined fused [RF layer -> GP layer] forward for
// large problems (input width <= 128).  Same arithmetic as k1_fwd_tc.cu; what changes is the execution model:
//
//   * the A operand of GEMM #1, (in * exp(log_inv_ls)) split into tf32 hi + lo, is written ONCE per CTA into
//     TENSOR MEMORY (tcgen05.st) and consumed from there by every column tile (tcgen05.mma with A in TMEM),
//     so shared memory only holds the streamed operands;
//   * roles:  warps 0-7  epilogue   tcgen05.ld P -> sincos/relu -> Phi tile (smem, UMMA/TMA swizzle)
//             warp  8    MMA        one thread issues GEMM #1 (3xTF32) of tile t+1 and GEMM #2 of tile t
//             warps 9-11 stagers    z tile (hi/lo) and W tile of the next column tile -> smem
//             warp  12   store      TMA store of the Phi tile (saved features)
//     connected by mbarriers; P (D1) is double-buffered in TMEM so GEMM #1 of tile t+1 and the staging of its
//     operands run under the epilogue of tile t (the MUFU/issue-bound part).
#include <stdlib.h>
#include "kernels.cuh"
#include "tc_common.cuh"

constexpr int V2_BM = 128, V2_BN = 64;
constexpr int V2_EPI_WARPS = 8, V2_STAGE_WARPS = 3;
constexpr int V2_THREADS = (V2_EPI_WARPS + 1 + V2_STAGE_WARPS + 1) * 32;      // 416
constexpr int V2_BLK = V2_BM * 128;               // [128 x 32 tf32] block
constexpr int V2_BBLK = V2_BN * 128;              // [64 x 32 tf32] block
constexpr int V2_KB = 4;                          // K blocks of GEMM #1 (input width <= 128)
constexpr uint32_t V2_TMEM_COLS = 512;            // A_hi 128 | A_lo 128 | D1 2x64 | D2 64

namespace tc {
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tc

template <int NG>
__global__ void __launch_bounds__(V2_THREADS, 1)
k1_fwd_tc2(const FwdArgs a, const __grid_constant__ CUtensorMap map_cos, const __grid_constant__ CUtensorMap map_sin) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sB1 = sm;                                   // [hi blocks 0..KB) | lo blocks 0..KB)] of [64 x 32]
    uint8_t* sPhi = sB1 + 2 * V2_KB * V2_BBLK;           // 4 blocks [128 x 32]: cos 0,1 | sin 2,3
    uint8_t* sW = sPhi + 4 * V2_BLK;                     // 2 stages x 4 blocks [NG x 32]
    float* bias_s = reinterpret_cast<float*>(sW + 2 * 4 * NG * 128);
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + V2_BM);
    uint64_t* b1_full = bars + 0;     // stagers (3 warps)        -> MMA
    uint64_t* b1_empty = bars + 1;    // MMA commit               -> stagers
    uint64_t* d1_full = bars + 2;     // [2] MMA commit           -> epilogue
    uint64_t* d1_empty = bars + 4;    // [2] epilogue (8 warps)   -> MMA
    uint64_t* phi_full = bars + 6;    // epilogue (8 warps)       -> MMA, store
    uint64_t* phi_empty = bars + 7;   // MMA commit + store warp  -> epilogue      (count 2)
    uint64_t* w_full = bars + 8;      // [2] stagers              -> MMA
    uint64_t* w_empty = bars + 10;    // [2] MMA commit           -> stagers
    uint64_t* d2_full = bars + 12;    // MMA commit               -> final epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.z, cs = blockIdx.y, row0 = blockIdx.x * V2_BM;
    const float* z = a.z + chain * a.z_cs;
    const float* X = a.X + chain * a.x_cs;
    const float* W = a.W + chain * a.w_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float scale = (rbf ? 1.f : 1.41421356237f) * __expf(__ldg(a.log_amp + chain * a.h_cs)) * rsqrtf((float)a.M);
    const int n_ct = (a.M + V2_BN - 1) / V2_BN;
    const int n_my = cs < n_ct ? (n_ct - cs + a.CS - 1) / a.CS : 0;        // column tiles of this CTA
    const int n_kb = (a.d + 31) / 32;
    const int nb2 = rbf ? 4 : 2;

    if (warp == V2_EPI_WARPS) tc::tmem_alloc(tmem_slot, V2_TMEM_COLS);
    if (tid == 0) {
        tc::mbar_init(b1_full, V2_STAGE_WARPS);
        tc::mbar_init(b1_empty, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(d1_full + i, 1);
            tc::mbar_init(d1_empty + i, V2_EPI_WARPS);
            tc::mbar_init(w_full + i, V2_STAGE_WARPS);
            tc::mbar_init(w_empty + i, 1);
        }
        tc::mbar_init(phi_full, V2_EPI_WARPS);
        tc::mbar_init(phi_empty, 2);
        tc::mbar_init(d2_full, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_ahi = tmem_base, tm_alo = tmem_base + 128, tm_d1 = tmem_base + 256, tm_d2 = tmem_base + 384;

    // ---- A operand -> TMEM (epilogue warps: thread = row, warp half = 64 K columns) ----
    if (warp < V2_EPI_WARPS) {
        const int lq = warp & 3, kh = warp >> 2;
        const int r = 32 * lq + lane;
        const int64_t row = row0 + r;
        float bsum = 0.f;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int q = 64 * kh + 16 * c16 + i;
                float v = 0.f;
                if (row < a.B && q < a.d)
                    v = q < a.d_prev ? slab_load(a.Fprev, chain, row, q) : __ldg(X + row * a.ldx + (q - a.d_prev));
                const float sq = q < a.d ? expf(__ldg(ls + q)) : 0.f;
                if (mean && q < a.d) bsum = fmaf(v, __ldg(mean + q), bsum);
                const float x = v * sq;
                hi[i] = tc::to_tf32(x);
                lo[i] = tc::to_tf32(x - hi[i]);
            }
            const uint32_t col = 64 * kh + 16 * c16;
            tc::tmem_st16(tm_ahi + ((uint32_t)(32 * lq) << 16) + col, hi);
            tc::tmem_st16(tm_alo + ((uint32_t)(32 * lq) << 16) + col, lo);
        }
        tc::tmem_st_wait();
        if (a.has_mean) {                         // bias_r = sum_q in[r][q] mean[q]: the two K halves of a row
            if (kh == 0) bias_s[r] = bsum;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (kh == 1) bias_s[r] += bsum;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    constexpr uint32_t IDESC1 = tc::make_idesc_tf32(V2_BM, V2_BN);
    constexpr uint32_t IDESC2 = tc::make_idesc_tf32(V2_BM, NG);

    if (warp < V2_EPI_WARPS) {
        // ===================================== EPILOGUE =====================================
        const int lq = warp & 3, chh = warp >> 2;
        const int r = 32 * lq + lane;
        const float bias = a.has_mean ? bias_s[r] : 0.f;
        for (int t = 0; t < n_my; ++t) {
            const int c0 = (cs + t * a.CS) * V2_BN;
            const int buf = t & 1;
            tc::mbar_wait(d1_full + buf, (t >> 1) & 1);
            tc::tc_fence_after();
            float p[32];
            tc::tmem_ld32(tm_d1 + 64 * buf + ((uint32_t)(32 * lq) << 16) + 32 * chh, p);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(d1_empty + buf);          // P is in registers: the buffer is free
            float f1[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const bool live = (c0 + 32 * chh + i) < a.M;
                const float x = p[i] + bias;
                if (rbf) {
                    float s, c;
                    sincos_cw(x, &s, &c);
                    p[i] = live ? scale * c : 0.f;
                    f1[i] = live ? scale * s : 0.f;
                } else {
                    p[i] = live ? scale * fmaxf(x, 0.f) : 0.f;
                }
            }
            tc::mbar_wait(phi_empty, (t & 1) ^ 1);                   // GEMM #2 and the store of tile t-1 are done
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                *reinterpret_cast<float4*>(sPhi + chh * V2_BLK + tc::sw128_chunk(r, c4)) =
                    make_float4(p[4 * c4], p[4 * c4 + 1], p[4 * c4 + 2], p[4 * c4 + 3]);
                if (rbf)
                    *reinterpret_cast<float4*>(sPhi + (2 + chh) * V2_BLK + tc::sw128_chunk(r, c4)) =
                        make_float4(f1[4 * c4], f1[4 * c4 + 1], f1[4 * c4 + 2], f1[4 * c4 + 3]);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(phi_full);
        }
        // ---- final: F partial slab of this column split ----
        if (a.do_gemm2 && n_my > 0 && warp < 4) {
            tc::mbar_wait(d2_full, 0);
            tc::tc_fence_after();
            const int64_t row = row0 + r;
#pragma unroll
            for (int c16 = 0; c16 < NG / 16; ++c16) {
                float v[16];
                tc::tmem_ld16(tm_d2 + ((uint32_t)(32 * warp) << 16) + 16 * c16, v);
                tc::tmem_ld_wait();
                if (row < a.B) {
                    float* dst = a.Fpart + chain * a.fpart_cs + ((int64_t)cs * a.B + row) * a.g;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (16 * c16 + j < a.g) dst[16 * c16 + j] = v[j];
                }
            }
        }
    } else if (warp == V2_EPI_WARPS) {
        // ===================================== MMA ISSUER =====================================
        if (lane == 0) {
            // software pipeline: GEMM #1 of tile t is issued before GEMM #2 of tile t-1
            for (int t = 0; t <= n_my; ++t) {
                if (t < n_my) {
                    const int buf = t & 1;
                    tc::mbar_wait(d1_empty + buf, ((t >> 1) & 1) ^ 1);
                    tc::mbar_wait(b1_full, t & 1);
                    tc::tc_fence_after();
                    for (int kb = 0; kb < n_kb; ++kb) {
                        const int kleft = a.d - 32 * kb;
                        const int ksteps = kleft >= 32 ? 4 : (kleft + 7) / 8;
                        for (int k4 = 0; k4 < ksteps; ++k4) {
                            const uint32_t acol = 32 * kb + 8 * k4;
                            const uint64_t dbh = tc::make_desc_sw128(tc::smem_u32(sB1 + kb * V2_BBLK) + k4 * 32);
                            const uint64_t dbl = tc::make_desc_sw128(tc::smem_u32(sB1 + (V2_KB + kb) * V2_BBLK) + k4 * 32);
                            tc::umma_tf32_ts(tm_d1 + 64 * buf, tm_alo + acol, dbh, IDESC1, (kb | k4) != 0);
                            tc::umma_tf32_ts(tm_d1 + 64 * buf, tm_ahi + acol, dbl, IDESC1, 1u);
                            tc::umma_tf32_ts(tm_d1 + 64 * buf, tm_ahi + acol, dbh, IDESC1, 1u);
                        }
                    }
                    tc::umma_commit(b1_empty);             // z tile consumed
                    tc::umma_commit(d1_full + buf);        // P ready
                }
                if (t > 0) {
                    const int u = t - 1, ws = u & 1;
                    tc::mbar_wait(phi_full, u & 1);
                    if (a.do_gemm2) {
                        tc::mbar_wait(w_full + ws, (u >> 1) & 1);
                        tc::tc_fence_after();
                        for (int b = 0; b < nb2; ++b)
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::umma_tf32(tm_d2, tc::make_desc_sw128(tc::smem_u32(sPhi + b * V2_BLK) + k4 * 32),
                                              tc::make_desc_sw128(tc::smem_u32(sW + (ws * 4 + b) * (NG * 128)) + k4 * 32),
                                              IDESC2, (u | b | k4) != 0);
                        tc::umma_commit(w_empty + ws);
                    }
                    tc::umma_commit(phi_empty);            // Phi tile consumed by the tensor core (1 of 2 arrivals)
                    if (u == n_my - 1) tc::umma_commit(d2_full);
                }
            }
        }
    } else if (warp < V2_EPI_WARPS + 1 + V2_STAGE_WARPS) {
        // ===================================== STAGERS =====================================
        const int st = tid - (V2_EPI_WARPS + 1) * 32;                // 0..95
        constexpr int NST = V2_STAGE_WARPS * 32;
        for (int t = 0; t < n_my; ++t) {
            const int c0 = (cs + t * a.CS) * V2_BN;
            // ---- z tile (B of GEMM #1): rows = feature columns, 16-byte chunks along K, tf32 hi / lo ----
            tc::mbar_wait(b1_empty, (t & 1) ^ 1);
            {
                constexpr int ZB = 6;                                  // items (4 loads each) in flight per thread
                const int n_items = n_kb * V2_BN * 8;
                for (int e0 = st; e0 < n_items; e0 += NST * ZB) {
                    float zx[ZB][4];
#pragma unroll
                    for (int u = 0; u < ZB; ++u) {
                        const int e = e0 + u * NST;
                        const int n = e % V2_BN, ch = (e / V2_BN) & 7, kb = e / (V2_BN * 8);
                        const int col = c0 + n;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int q = kb * 32 + ch * 4 + i;
                            zx[u][i] = (e < n_items && q < a.d && col < a.M) ? __ldg(z + (int64_t)q * a.M + col) : 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < ZB; ++u) {
                        const int e = e0 + u * NST;
                        if (e < n_items) {
                            const int n = e % V2_BN, ch = (e / V2_BN) & 7, kb = e / (V2_BN * 8);
                            float4 o, ol;
                            o.x = tc::to_tf32(zx[u][0]); o.y = tc::to_tf32(zx[u][1]); o.z = tc::to_tf32(zx[u][2]); o.w = tc::to_tf32(zx[u][3]);
                            ol.x = tc::to_tf32(zx[u][0] - o.x); ol.y = tc::to_tf32(zx[u][1] - o.y);
                            ol.z = tc::to_tf32(zx[u][2] - o.z); ol.w = tc::to_tf32(zx[u][3] - o.w);
                            *reinterpret_cast<float4*>(sB1 + kb * V2_BBLK + tc::sw128_chunk(n, ch)) = o;
                            *reinterpret_cast<float4*>(sB1 + (V2_KB + kb) * V2_BBLK + tc::sw128_chunk(n, ch)) = ol;
                        }
                    }
                }
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(b1_full);
            // ---- W tile (B of GEMM #2) into ring stage t & 1 ----
            if (a.do_gemm2) {
                const int ws = t & 1;
                tc::mbar_wait(w_empty + ws, ((t >> 1) & 1) ^ 1);
                uint8_t* dstW = sW + ws * 4 * (NG * 128);
                {
                    constexpr int WB = 16;                             // loads in flight per thread
                    const int n_el = nb2 * 32 * NG;
                    for (int e0 = st; e0 < n_el; e0 += NST * WB) {
                        float wv[WB];
#pragma unroll
                        for (int u = 0; u < WB; ++u) {
                            const int e = e0 + u * NST;
                            const int j = e % NG, kf = e / NG;
                            const int col = c0 + (kf & 63);
                            const int64_t frow = (kf >= 64 ? a.M : 0) + col;
                            wv[u] = (e < n_el && j < a.g && col < a.M) ? __ldg(W + frow * a.g + j) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < WB; ++u) {
                            const int e = e0 + u * NST;
                            if (e < n_el) {
                                const int j = e % NG, kf = e / NG;
                                *reinterpret_cast<float*>(dstW + (kf >> 5) * (NG * 128) + tc::sw128_off(j, kf & 31)) = tc::to_tf32(wv[u]);
                            }
                        }
                    }
                }
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(w_full + ws);
            }
        }
    } else {
        // ===================================== STORE WARP =====================================
        if (lane == 0) {
            for (int t = 0; t < n_my; ++t) {
                const int c0 = (cs + t * a.CS) * V2_BN;
                tc::mbar_wait(phi_full, t & 1);
                if (a.Phi != nullptr) {
                    for (int b = 0; b < 2; ++b) {
                        if (c0 + 32 * b >= a.M) break;
                        tc::tma_store_3d(&map_cos, tc::smem_u32(sPhi + b * V2_BLK), c0 + 32 * b, row0, chain);
                        if (rbf) tc::tma_store_3d(&map_sin, tc::smem_u32(sPhi + (2 + b) * V2_BLK), c0 + 32 * b, row0, chain);
                    }
                    tc::tma_commit();
                    tc::tma_wait_read0();
                }
                tc::mbar_arrive(phi_empty);                // 2 of 2 arrivals: the Phi tile may be overwritten
            }
            tc::tma_wait0();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == V2_EPI_WARPS) tc::tmem_dealloc(tmem_base, V2_TMEM_COLS);
}

static size_t tc2_smem_bytes(int NG) {
    return 1024 + 2 * (size_t)V2_KB * V2_BBLK + 4 * (size_t)V2_BLK + 2 * 4 * (size_t)NG * 128 + V2_BM * sizeof(float) + 16 * 8 + 16;
}

template <int NG>
static int launch_fwd_tc2(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const size_t smem = tc2_smem_bytes(NG);
    static bool configured = false;
    if (!configured) {
        DGPRF_CHECK_CUDA(cudaFuncSetAttribute(k1_fwd_tc2<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    CUtensorMap mc, ms;
    memset(&mc, 0, sizeof(mc));
    memset(&ms, 0, sizeof(ms));
    if (a.Phi != nullptr) {
        int rc = dgprf_make_tmap_3d(&mc, a.Phi, a.M, a.B, n_chains, a.F, a.phi_cs, V2_BM);
        if (rc) return rc;
        if (a.kind == DGPRF_KIND_RBF) {
            rc = dgprf_make_tmap_3d(&ms, a.Phi + a.M, a.M, a.B, n_chains, a.F, a.phi_cs, V2_BM);
            if (rc) return rc;
        }
    }
    dim3 grid(ceil_div(a.B, V2_BM), a.CS, n_chains);
    { ProfScope _ps("k1_fwd_tc2", st); k1_fwd_tc2<NG><<<grid, V2_THREADS, smem, st>>>(a, mc, ms); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// the pipelined kernel takes 64-column tiles and an input width that fits the TMEM-resident A operand
bool dgprf_fwd_tc2_supported(const FwdArgs& a) {
    return a.tile_cols == 64 && a.d <= 128 && (a.M % 4 == 0) && a.g <= 64 && (a.Phi == nullptr || (a.phi_cs % 4) == 0) &&
           getenv("DGPRF_NO_TC2") == nullptr;
}

int dgprf_launch_fwd_tc2(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const int g = a.do_gemm2 ? a.g : 1;
    if (g <= 16) return launch_fwd_tc2<16>(a, n_chains, st);
    if (g <= 32) return launch_fwd_tc2<32>(a, n_chains, st);
    return launch_fwd_tc2<64>(a, n_chains, st);
}
