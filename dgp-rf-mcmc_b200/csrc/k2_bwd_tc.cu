// K2 (tensor-core variant, DGPRF_PREC_TF32): reverse pass of one [RF layer -> GP layer] pair on tcgen05.
//
// Per (128-row tile, 64-column tile) three tf32 UMMAs share the tiles in shared memory / TMEM:
//   MMA-1  dPhi[128 x 128]   = dF[128 x g] . W_tile^T            A = dF (K-major), B = W rows (K-major)
//   epi-1  dP = Phi_c*dPhi_s - Phi_s*dPhi_c | dPhi*scale*[Phi>0]  Phi tile TMA-LOADED from the saved features
//   MMA-2  gW_tile[128 x g] += Phi_tile^T . dF                    A = Phi tile as MN-major: tf32 MN-major operands must
//                                                                 use the 128B swizzle with a 32-byte atom, which is
//                                                                 exactly what TMA writes with SWIZZLE_128B_ATOM_32B,
//                                                                 so the loaded tile is consumed without a transpose;
//                                                                 B = a second copy of the dF tile in that layout;
//                                                                 accumulated in TMEM over the row tiles of the CTA
//   MMA-3  T[128 x d_prev]   = dP[128 x 64] . z_tile^T            A = dP (K-major, written by epi-1), B = z rows
//   epi-3  dF_prev slab     (+)= exp(log_inv_ls)*T + mean*rowsum(dP)
// Decomposition, slab protocol and formulas are those of the SIMT kernel (k2_bwd_simt.cu); W-only mode
// (hyper-parameter gradients stay on the SIMT kernel).
#include <stdio.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "tc_common.cuh"

constexpr int BT_BM = 128;            // batch rows per tile
constexpr int BT_BN = 64;             // feature columns per tile -> 128 features (cos | sin) = UMMA M of MMA-2
constexpr int BT_THREADS = 256;
constexpr int BT_BLK = BT_BM * 128;   // bytes of a [128 x 32 tf32] block
constexpr uint32_t BT_TMEM_COLS = 256;   // D1: 128 | D2: <= 64 | D3: <= 64


__device__ float* g_bwd_dbg = nullptr;     // debug dump target (DGPRF_BWD_TC_DEBUG=1)

template <int NG>      // padded n_gp: 32 | 64  (K blocks of MMA-1 = NG/32, N of MMA-2 = NG)
__global__ void __launch_bounds__(BT_THREADS, 1)
k2_bwd_tc(const BwdArgs a, const __grid_constant__ CUtensorMap map_cos, const __grid_constant__ CUtensorMap map_sin) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int KGB = NG / 32;
    // 1024-byte alignment by offsetting the shared array itself: the pointer keeps its shared-memory address
    // space (a round trip through uintptr_t makes every access a generic LD/ST)
    uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sPhi = sm;                              // 4 blocks: cos 0,1 | sin 2,3
    uint8_t* sdF = sPhi + 4 * BT_BLK;                // KGB blocks [128 rows x 32 j]  K-major, 16-byte-atom swizzle
    uint8_t* sdF2 = sdF + KGB * BT_BLK;              // the same tile, MN-major 32-byte-atom swizzle (B of MMA-2)
    uint8_t* sW = sdF2 + KGB * BT_BLK;               // KGB blocks [128 feature rows x 32 j]
    uint8_t* sdP = sW + KGB * BT_BLK;                // 2 blocks [128 rows x 32 feature cols]
    uint8_t* sZ = sdP + 2 * BT_BLK;                  // 2 blocks [64 q rows x 32 feature cols]
    float* R_s = reinterpret_cast<float*>(sZ + 2 * 64 * 128);   // [2][128] row sums of dP per column half
    float* s_s = R_s + 2 * BT_BM;                    // [64] exp(log_inv_ls[q]), q < d_prev
    float* m_s = s_s + 64;                           // [64] mean[q]
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(m_s + 64);
    uint64_t* bar_tma = bar_mma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.z, cs = blockIdx.y, rs = blockIdx.x;
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float* z = a.z + chain * a.z_cs;
    const float* W = a.W + chain * a.w_cs;
    const float arc_scale = 1.41421356237f * __expf(__ldg(a.log_amp + chain * a.h_cs)) * rsqrtf((float)a.M);
    const int NQ = a.d_prev > 0 ? ((a.d_prev + 15) & ~15) : 0;      // UMMA N of MMA-3

    if (warp == 0) tc::tmem_alloc(tmem_slot, BT_TMEM_COLS);
    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_tma, 1);
        tc::mbar_fence_init();
    }
    if (tid < 64) {
        const bool ok = tid < a.d_prev;
        s_s[tid] = ok ? expf(__ldg(a.log_inv_ls + chain * a.h_cs + tid)) : 0.f;
        m_s[tid] = (ok && a.has_mean) ? __ldg(a.mean + chain * a.h_cs + tid) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + 128, tmem_d3 = tmem_base + 192;
    uint32_t ph_mma = 0, ph_tma = 0;

    const uint32_t IDESC1 = tc::make_idesc_tf32(BT_BM, 2 * BT_BN);
    const uint32_t IDESC2 = tc::make_idesc_tf32_mn(BT_BM, NG);
    const uint32_t IDESC3 = tc::make_idesc_tf32(BT_BM, NQ > 0 ? NQ : 16);
    const uint32_t phi_bytes = (rbf ? 4u : 2u) * BT_BLK;

    const int n_ct = (a.M + BT_BN - 1) / BT_BN, n_rt = (a.B + BT_BM - 1) / BT_BM;
    for (int ct = cs; ct < n_ct; ct += a.CS) {
        const int c0 = ct * BT_BN;
        const bool first_ct = ct == cs;
        // ---- per column tile: W rows (B of MMA-1) and z rows (B of MMA-3), tf32, K-major ----
        for (int e = tid; e < 2 * BT_BN * 32 * KGB; e += BT_THREADS) {
            const int n = e / (32 * KGB), j = e % (32 * KGB);
            const int col = c0 + (n & (BT_BN - 1));
            const int64_t frow = (n >= BT_BN ? a.M : 0) + col;
            const bool ok = col < a.M && j < a.g && (rbf || n < BT_BN);
            const float v = ok ? tc::to_tf32(__ldg(W + frow * a.g + j)) : 0.f;
            *reinterpret_cast<float*>(sW + (j >> 5) * BT_BLK + tc::sw128_off(n, j & 31)) = v;
        }
        if (NQ > 0) {
            for (int e = tid; e < NQ * BT_BN; e += BT_THREADS) {
                const int q = e / BT_BN, k = e % BT_BN;
                const float v = (q < a.d_prev && c0 + k < a.M) ? tc::to_tf32(__ldg(z + (int64_t)q * a.M + c0 + k)) : 0.f;
                *reinterpret_cast<float*>(sZ + (k >> 5) * (64 * 128) + tc::sw128_off(q, k & 31)) = v;
            }
        }
        bool d2_started = false;

        for (int rt = rs; rt < n_rt; rt += a.RS) {
            const int row0 = rt * BT_BM;
            // ---- Phi tile: TMA load of the saved features (OOB rows / columns arrive as zeros) ----
            if (warp == 0 && tc::elect_one()) {
                tc::mbar_expect_tx(bar_tma, phi_bytes);
                for (int b = 0; b < 2; ++b) {
                    tc::tma_load_3d(&map_cos, tc::smem_u32(sPhi + b * BT_BLK), bar_tma, c0 + 32 * b, row0, chain);
                    if (rbf) tc::tma_load_3d(&map_sin, tc::smem_u32(sPhi + (2 + b) * BT_BLK), bar_tma, c0 + 32 * b, row0, chain);
                }
            }
            // ---- dF tile (A of MMA-1, B of MMA-2) ----
            {
                constexpr int NE = BT_BM * 32 * KGB / BT_THREADS;
                float v[NE];
#pragma unroll
                for (int u = 0; u < NE; ++u) {
                    const int e = tid + u * BT_THREADS;
                    const int r = e / (32 * KGB), j = e % (32 * KGB);
                    const int64_t row = row0 + r;
                    v[u] = (row < a.B && j < a.g) ? slab_load(a.dF, chain, row, j) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < NE; ++u) {
                    const int e = tid + u * BT_THREADS;
                    const int r = e / (32 * KGB), j = e % (32 * KGB);
                    const float t = tc::to_tf32(v[u]);
                    *reinterpret_cast<float*>(sdF + (j >> 5) * BT_BLK + tc::sw128_off(r, j & 31)) = t;
                    *reinterpret_cast<float*>(sdF2 + (j >> 5) * BT_BLK + tc::sw128b32_off(r, j & 31)) = t;
                }
            }
            tc::fence_async_smem();
            __syncthreads();
            if (warp == 0 && tc::elect_one()) {               // MMA-1: dPhi = dF W^T
                tc::tc_fence_after();
                for (int kb = 0; kb < KGB; ++kb) {
                    const int kleft = a.g - 32 * kb;
                    const int ksteps = kleft >= 32 ? 4 : (kleft > 0 ? (kleft + 7) / 8 : 0);
                    for (int k4 = 0; k4 < ksteps; ++k4)
                        tc::umma_tf32(tmem_d1, tc::make_desc_sw128(tc::smem_u32(sdF + kb * BT_BLK) + k4 * 32),
                                      tc::make_desc_sw128(tc::smem_u32(sW + kb * BT_BLK) + k4 * 32), IDESC1, (kb | k4) != 0);
                }
                tc::umma_commit(bar_mma);
            }
            tc::mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc::mbar_wait(bar_tma, ph_tma); ph_tma ^= 1;
            tc::tc_fence_after();

            if (g_bwd_dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && ct == cs && rt == rs && warp == 0) {
                float d1[16];
                tc::tmem_ld16(tmem_d1, d1);
                tc::tmem_ld_wait();
                if (lane < 4) {
                    for (int c = 0; c < 8; ++c) {
                        g_bwd_dbg[lane * 8 + c] = *reinterpret_cast<float*>(sPhi + tc::sw128b32_off(lane, c));       // Phi_c[r][c]
                        g_bwd_dbg[32 + lane * 8 + c] = *reinterpret_cast<float*>(sdF + tc::sw128_off(lane, c));      // dF[r][j]
                        g_bwd_dbg[64 + lane * 8 + c] = *reinterpret_cast<float*>(sW + tc::sw128_off(lane, c));       // W[n][j]
                        g_bwd_dbg[96 + lane * 8 + c] = d1[c];                                                        // dPhi_c[r][c]
                    }
                }
            }
            // ---- epilogue 1: dP from dPhi (TMEM) and the Phi tile (smem) -> dP tile (A of MMA-3) ----
            {
                const int lq = warp & 3, hh = warp >> 2;             // TMEM lane quarter, 32-column half
                const int r = 32 * lq + lane;
                float rsum = 0.f;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    float dc[16], ds[16];
                    const uint32_t lane_addr = (uint32_t)(32 * lq) << 16;
                    tc::tmem_ld16(tmem_d1 + lane_addr + 32 * hh + 16 * pass, dc);
                    if (rbf) tc::tmem_ld16(tmem_d1 + lane_addr + BT_BN + 32 * hh + 16 * pass, ds);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int cc = 4 * pass + c4;                // 16-byte chunk inside the 32-wide block
                        const float4 pc = *reinterpret_cast<const float4*>(sPhi + hh * BT_BLK + tc::sw128b32_chunk(r, cc));
                        float4 o;
                        if (rbf) {
                            const float4 ps = *reinterpret_cast<const float4*>(sPhi + (2 + hh) * BT_BLK + tc::sw128b32_chunk(r, cc));
                            o.x = pc.x * ds[4 * c4 + 0] - ps.x * dc[4 * c4 + 0];
                            o.y = pc.y * ds[4 * c4 + 1] - ps.y * dc[4 * c4 + 1];
                            o.z = pc.z * ds[4 * c4 + 2] - ps.z * dc[4 * c4 + 2];
                            o.w = pc.w * ds[4 * c4 + 3] - ps.w * dc[4 * c4 + 3];
                        } else {
                            o.x = pc.x > 0.f ? dc[4 * c4 + 0] * arc_scale : 0.f;
                            o.y = pc.y > 0.f ? dc[4 * c4 + 1] * arc_scale : 0.f;
                            o.z = pc.z > 0.f ? dc[4 * c4 + 2] * arc_scale : 0.f;
                            o.w = pc.w > 0.f ? dc[4 * c4 + 3] * arc_scale : 0.f;
                        }
                        rsum += (o.x + o.y) + (o.z + o.w);
                        o.x = tc::to_tf32(o.x); o.y = tc::to_tf32(o.y); o.z = tc::to_tf32(o.z); o.w = tc::to_tf32(o.w);
                        *reinterpret_cast<float4*>(sdP + hh * BT_BLK + tc::sw128_chunk(r, cc)) = o;
                    }
                }
                R_s[hh * BT_BM + r] = rsum;
            }
            tc::tc_fence_before();
            tc::fence_async_smem();
            __syncthreads();
            if (warp == 0 && tc::elect_one()) {
                tc::tc_fence_after();
                // MMA-2: gW_tile += Phi_tile^T dF  (both operands MN-major; K = 128 batch rows, 8 per instruction)
                for (int k8 = 0; k8 < BT_BM / 8; ++k8)
                    tc::umma_tf32(tmem_d2, tc::make_desc_mn_b32(tc::smem_u32(sPhi) + k8 * 1024, BT_BLK, 512),
                                  tc::make_desc_mn_b32(tc::smem_u32(sdF2) + k8 * 1024, BT_BLK, 512), IDESC2,
                                  (d2_started || k8 != 0) ? 1u : 0u);
                // MMA-3: T = dP z_tile^T
                if (NQ > 0)
                    for (int b = 0; b < 2; ++b)
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::umma_tf32(tmem_d3, tc::make_desc_sw128(tc::smem_u32(sdP + b * BT_BLK) + k4 * 32),
                                          tc::make_desc_sw128(tc::smem_u32(sZ + b * (64 * 128)) + k4 * 32), IDESC3, (b | k4) != 0);
                tc::umma_commit(bar_mma);
            }
            d2_started = true;
            tc::mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc::tc_fence_after();

            if (g_bwd_dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && ct == cs && rt == rs && warp == 0) {
                float d2[16], d3[16];
                tc::tmem_ld16(tmem_d2, d2);
                tc::tmem_ld16(tmem_d3, d3);
                tc::tmem_ld_wait();
                if (lane < 4) {
                    for (int c = 0; c < 8; ++c) {
                        g_bwd_dbg[128 + lane * 8 + c] = d2[c];                                                       // gW[f][j]
                        g_bwd_dbg[160 + lane * 8 + c] = d3[c];                                                       // T[r][q]
                        g_bwd_dbg[192 + lane * 8 + c] = *reinterpret_cast<float*>(sdP + tc::sw128_off(lane, c));     // dP[r][c]
                    }
                }
            }
            // ---- epilogue 3: dF_prev slab (+)= s*T + mean*R ----
            if (NQ > 0 && warp < 4 && a.Dpart != nullptr) {
                const int r = 32 * warp + lane;
                const int64_t row = row0 + r;
                const float Rr = R_s[r] + R_s[BT_BM + r];
                for (int qc = 0; qc < NQ / 16; ++qc) {
                    float t[16];
                    tc::tmem_ld16(tmem_d3 + ((uint32_t)(32 * warp) << 16) + 16 * qc, t);
                    tc::tmem_ld_wait();
                    if (row < a.B) {
                        float* dst = a.Dpart + chain * a.d_cs + ((int64_t)cs * a.B + row) * a.d_prev;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int q = 16 * qc + i;
                            if (q < a.d_prev) {
                                float v = s_s[q] * t[i];
                                if (a.has_mean) v = fmaf(m_s[q], Rr, v);
                                dst[q] = first_ct ? v : dst[q] + v;
                            }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncthreads();          // tiles and D1/D3 are free for the next row tile
        }

        // ---- epilogue 2: this column tile's gW rows -> row-split slab rs ----
        if (warp < 4) {
            const int fl = 32 * warp + lane;                       // feature row inside [cos 64 | sin 64]
            const int col = c0 + (fl & (BT_BN - 1));
            const bool live = col < a.M && (rbf || fl < BT_BN);
            const int64_t frow = (fl >= BT_BN ? a.M : 0) + col;
            float* dst = a.gWpart + chain * a.gw_cs + (int64_t)rs * a.gw_ss + frow * a.g;
#pragma unroll
            for (int c16 = 0; c16 < NG / 16; ++c16) {
                float v[16];
                tc::tmem_ld16(tmem_d2 + ((uint32_t)(32 * warp) << 16) + 16 * c16, v);
                tc::tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (16 * c16 + j < a.g) dst[16 * c16 + j] = d2_started ? v[j] : 0.f;
                }
            }
        }
        tc::tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tc::tmem_dealloc(tmem_base, BT_TMEM_COLS);
}

static size_t tc_bwd_smem_bytes(int NG) {
    const int KGB = NG / 32;
    return 1024 + 4 * (size_t)BT_BLK + 3 * (size_t)KGB * BT_BLK + 2 * (size_t)BT_BLK + 2 * 64 * 128 +
           sizeof(float) * (2 * BT_BM + 128) + 32;
}

template <int NG>
static int launch_bwd_tc(const BwdArgs& a, int n_chains, cudaStream_t st) {
    const size_t smem = tc_bwd_smem_bytes(NG);
    { const int rc_s = dgprf_ensure_smem((const void*)k2_bwd_tc<NG>, (size_t)smem); if (rc_s) return rc_s; }
    CUtensorMap mc, ms;
    memset(&mc, 0, sizeof(mc));
    memset(&ms, 0, sizeof(ms));
    int rc = dgprf_make_tmap_3d(&mc, a.Phi, a.M, a.B, n_chains, a.F, a.phi_cs, BT_BM, true);
    if (rc) return rc;
    if (a.kind == DGPRF_KIND_RBF) {
        rc = dgprf_make_tmap_3d(&ms, a.Phi + a.M, a.M, a.B, n_chains, a.F, a.phi_cs, BT_BM, true);
        if (rc) return rc;
    }
    dim3 grid(a.RS, a.CS, n_chains);
    static float* dbg = nullptr;
    if (getenv("DGPRF_BWD_TC_DEBUG") && !dbg) {
        cudaMalloc(&dbg, 256 * sizeof(float));
        cudaMemcpyToSymbol(g_bwd_dbg, &dbg, sizeof(dbg));
    }
    { ProfScope _ps("k2_bwd_tc", st); k2_bwd_tc<NG><<<grid, BT_THREADS, smem, st>>>(a, mc, ms); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    if (dbg) {
        float h[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[7] = {"Phi_c", "dF", "W", "dPhi_c", "gW(D2)", "T(D3)", "dP"};
        fprintf(stderr, "bwd_tc debug: kind=%d M=%d g=%d d_prev=%d\n", a.kind, a.M, a.g, a.d_prev);
        for (int k = 0; k < 7; ++k) {
            fprintf(stderr, "  %-8s", names[k]);
            for (int i = 0; i < 16; ++i) fprintf(stderr, " % .4e", h[32 * k + i]);
            fprintf(stderr, "\n");
        }
    }
    return DGPRF_OK;
}

// W-only backward of shapes the tensor-core kernel takes; everything else stays on the SIMT kernel
bool dgprf_bwd_tc_supported(const BwdArgs& a) {
    return !a.hyper && (a.M % 4 == 0) && a.g <= 64 && a.d_prev <= 64 && (a.phi_cs % 4) == 0;
}

int dgprf_launch_bwd_tc(const BwdArgs& a, int n_chains, cudaStream_t st) {
    return a.g <= 32 ? launch_bwd_tc<32>(a, n_chains, st) : launch_bwd_tc<64>(a, n_chains, st);
}
