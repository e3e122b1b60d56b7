// K2 (tensor-core, DGPRF_PREC_TF32): reverse pass of one [RF layer -> GP layer] pair with every accumulation
// kept on chip.  Three UMMAs per (128-row tile, 64-column tile),
//   MMA-1  dPhi = dF . W_tile^T      MMA-2  gW_tile += Phi_tile^T . dF      MMA-3  T += dP . z_tile^T
// and the CTA (row split rs, column split cs) walks its ROW tiles in the outer loop and its column tiles in the
// inner loop:
//   * all gW tiles of the CTA's column tiles stay resident in TENSOR MEMORY for the whole kernel (up to 10 tiles of
//     32 columns) and are written once, as row-split slab rs, at the end;
//   * T = dP z^T accumulates in TMEM over the column tiles of a row tile, so the dF_prev slab is written once per
//     row tile instead of being read-modified-written per column tile;
//   * the saved-feature tile (64 KB, the HBM stream that bounds this kernel) arrives through a 2-stage TMA ring;
//     the stage is released as soon as MMA-2 and the dP epilogue have consumed it, so the load of tile t+2 is in
//     flight during tile t and t+1;
//   * the W rows and z rows of a column tile are TMA loads too (W from a zero-padded copy [F][32] made by a prep
//     kernel because a TMA row stride must be a multiple of 16 bytes; z directly from the model's spectral draws).
// n_gp <= 32.  W-only mode: d_prev <= 64.  Hyper mode (stochastic-EM / full-Bayes gradients): T is formed for ALL input
// columns (width <= 128, two passes of 64 z rows through the same z tile), and the raw T and R = rowsum(dP) are written as
// the partial slabs the hyper reduction (k_hyper_partial / k_hyper_final) consumes.  Other shapes stay on the fp32
// SIMT kernel.  Column splits come from a small cost model (dgprf_bwd_tc2_pick_cs); debug: DGPRF_BWD_TIMELINE=<call number>
// prints per-role clock stamps of one CTA, DGPRF_BWD_PF=<n> adds TMA L2 prefetches n tiles ahead of the ring (no gain measured).
#include <stdio.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "tc_common.cuh"

constexpr int B2_BM = 128, B2_BN = 64, B2_NG = 32;
// Epilogue warps: 8 (a warp owns 32 rows x 32 tile columns) or 16 (32 rows x 16 columns; -DB2_EW=16).  Measured at configs[4]
// scale: 16 warps are no faster (65 536 rows: 528-543 us per layer against 513-527; 8 192 rows: 86-88 against 88-93) -- the
// per-tile chain is the waits (Phi arrival, dPhi, dP consumed) and shared-memory bandwidth, not the warps' instruction count.
#ifndef B2_EW
#define B2_EW 8
#endif
constexpr int B2_EPI_WARPS = B2_EW;
constexpr int B2_CQ = B2_EPI_WARPS / 4;        // column groups of a tile (2 x 32 | 4 x 16 columns)
constexpr int B2_CW = B2_BN / B2_CQ;           // tile columns per epilogue thread
constexpr int B2_NCH = B2_CW / 4;              // 16-byte chunks per epilogue thread and half (cos | sin)
constexpr int B2_THREADS = (B2_EPI_WARPS + 2) * 32;     // epilogue warps + the MMA issuer warp + the TMA producer warp
constexpr int B2_BLK = B2_BM * 128;            // bytes of a [128 x 32 tf32] block
constexpr int B2_HDR = 2048;                   // R_s | s_s | m_s | mbarriers | TMEM slot
constexpr int B2_MAX_LOC = 10;                 // resident gW tiles: 10 x 32 TMEM columns
constexpr uint32_t B2_TMEM_COLS = 512;         // D1 128 | D3 64 | D2 320

namespace tc {
__device__ __forceinline__ void mbar_arrive_b2(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tc

// (Measured and not kept, round 2: a third single-thread role issuing MMA-2 the moment a saved-feature tile lands instead of
// inside the main issuer's loop -- 535-547 us per configs[4] layer against 507-523; sixteen epilogue warps; an L2 prefetch
// ahead of the ring; the next row tile's dF loaded a tile early.  profiles/r02_summary.md section F.)
// Roles: warps 0-7 = epilogue (dF staging, dP, dF_prev, gW write-out); warp 8 = MMA issuer (one elected thread issues every
// UMMA); warp 9 = TMA producer (one elected thread issues every load).  Two single-thread roles because each mbarrier
// wait costs a few hundred cycles even when it is already complete: one thread doing all nine waits of a tile took ~4000
// cycles per tile and bounded the kernel.  The sides talk through mbarriers only:
//   issuer -> epilogue   barA  MMA-1(k) done (dPhi in D1)        barB  MMA-3(k) done (dP / z tiles free, T final at a
//                        row end)                                 barC  MMA-2(k) done (Phi stage and dF tile consumed)
//   epilogue -> issuer   e_read  every warp has copied its Phi rows of tile k to registers (stage may be refilled)
//                        d1_free every warp holds its dPhi values of tile k in registers (MMA-1(k+1) may overwrite D1)
//                        dp_full every warp has written dP(k)
//                        df_full the dF tile of the next row tile is staged
__global__ void __launch_bounds__(B2_THREADS, 1)
k2_bwd_tc2(const BwdArgs a, const int pf_dist, long long* const tl, const __grid_constant__ CUtensorMap map_cos, const __grid_constant__ CUtensorMap map_sin,
           const __grid_constant__ CUtensorMap map_wp, const __grid_constant__ CUtensorMap map_z) {
    dgprf_pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_s = reinterpret_cast<float*>(sm) + 2 * B2_BM;     // [64] exp(log_inv_ls[q]), q < d_prev  (the first 1 KB is unused)
    float* m_s = s_s + 64;                                     // [64] mean[q]
    uint64_t* bars = reinterpret_cast<uint64_t*>(m_s + 64);
    uint64_t* phi_full = bars + 0;    // [2] TMA complete_tx
    uint64_t* w_full = bars + 2;      // TMA complete_tx
    uint64_t* z_full = bars + 3;      // TMA complete_tx
    uint64_t* barA = bars + 4;
    uint64_t* barB = bars + 5;
    uint64_t* barC = bars + 6;
    uint64_t* e_read = bars + 7;      // count 8 (stage 0; stage 1 is bars + 11)
    uint64_t* dp_full = bars + 8;     // count 8
    uint64_t* df_full = bars + 9;     // count 8
    uint64_t* d1_free = bars + 10;    // count 8
    uint64_t* e_read1 = bars + 11;    // count 8
    uint64_t* barZ = bars + 12;       // first z pass of MMA-3 done (hyper mode with more than 64 input columns)
    uint64_t* barC1 = bars + 13;      // barC of the odd tiles (one barrier per Phi stage, like e_read)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    uint8_t* sPhi = sm + B2_HDR;                     // 2 stages x 4 blocks: cos 0,1 | sin 2,3   (32-byte-atom swizzle)
    uint8_t* sdF = sPhi + 2 * 4 * B2_BLK;            // [128 rows x 32 j]  K-major, 16-byte-atom swizzle (A of MMA-1)
    uint8_t* sdF2 = sdF + B2_BLK;                    // the same tile, MN-major 32-byte-atom swizzle (B of MMA-2)
    uint8_t* sdP = sdF2 + B2_BLK;                    // 2 blocks [128 rows x 32 feature cols]          (A of MMA-3)
    float* R_s = reinterpret_cast<float*>(sdP);      // [B2_CQ][128] row sums of dP per column group: ALIASES the dP tile, only
                                                     // touched at a row end between "MMA-3 of the last tile retired" and the next dP
    uint8_t* sW = sdP + 2 * B2_BLK;                  // [128 feature rows x 32 j]                      (B of MMA-1)
    uint8_t* sZ = sW + B2_BLK;                       // 2 blocks [64 q rows x 32 feature cols]         (B of MMA-3)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.z, cs = blockIdx.y, rs = blockIdx.x;
    // debug timeline (DGPRF_BWD_TIMELINE=1): clock stamps of CTA 0, first 16 tiles, 12 events per tile
#define TLB(t, ev) do { if (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (t) < 16) tl[(t) * 12 + (ev)] = clock64(); } while (0)
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float arc_scale = 1.41421356237f * __expf(__ldg(a.log_amp + chain * a.h_cs)) * rsqrtf((float)a.M);
    const bool hyper = a.hyper != 0;
    // input columns T is formed for: all of them in hyper mode, the previous layer's outputs otherwise
    const int NQ = hyper ? ((a.d + 15) & ~15) : (a.d_prev > 0 ? ((a.d_prev + 15) & ~15) : 0);
    const int nq0 = NQ < 64 ? NQ : 64, nq1 = NQ - nq0;               // UMMA N of the (up to two) MMA-3 passes
    const int zc = a.z_cs != 0 ? chain : 0;

    const int n_ct = (a.M + B2_BN - 1) / B2_BN, n_rt = (a.B + B2_BM - 1) / B2_BM;
    const int n_loc = cs < n_ct ? (n_ct - cs + a.CS - 1) / a.CS : 0;          // column tiles of this CTA
    const int n_rloc = rs < n_rt ? (n_rt - rs + a.RS - 1) / a.RS : 0;         // row tiles of this CTA
    const int T = n_loc * n_rloc;

    if (warp == B2_EPI_WARPS) tc::tmem_alloc(tmem_slot, B2_TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) tc::mbar_init(bars + i, 1);
        for (int i = 7; i < 12; ++i) tc::mbar_init(bars + i, B2_EPI_WARPS);
        tc::mbar_init(barZ, 1);
        tc::mbar_init(barC1, 1);
        tc::mbar_fence_init();
    }
    if (tid < 64) {
        const bool ok = tid < a.d_prev;
        s_s[tid] = ok ? expf(__ldg(a.log_inv_ls + chain * a.h_cs + tid)) : 0.f;
        m_s[tid] = (ok && a.has_mean) ? __ldg(a.mean + chain * a.h_cs + tid) : 0.f;
    }
    if (!rbf) {       // arc-cosine: the sin halves of the Phi stages and of the W tile are never loaded; keep them finite
        for (int e = tid; e < (2 * 4 * B2_BLK) / 16; e += B2_THREADS) reinterpret_cast<float4*>(sPhi)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = tid; e < B2_BLK / 16; e += B2_THREADS) reinterpret_cast<float4*>(sW)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        tc::fence_async_smem();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_d1 = tmem_base, tm_d3 = tmem_base + 128, tm_d2 = tmem_base + (hyper ? 256 : 192);

    // tile k of this CTA = (row tile k / n_loc, column tile k % n_loc)
    if (warp == B2_EPI_WARPS) {
        // ============================================ MMA ISSUER ============================================
        if (T > 0 && tc::elect_one()) {
            constexpr uint32_t IDESC1 = tc::make_idesc_tf32(B2_BM, 2 * B2_BN);
            constexpr uint32_t IDESC2 = tc::make_idesc_tf32_mn(B2_BM, B2_NG);
            const uint32_t IDESC3a = tc::make_idesc_tf32(B2_BM, nq0 > 0 ? nq0 : 16);
            const uint32_t IDESC3b = tc::make_idesc_tf32(B2_BM, nq1 > 0 ? nq1 : 16);
            const uint64_t d_dF = tc::make_desc_sw128(tc::smem_u32(sdF));
            const uint64_t d_W = tc::make_desc_sw128(tc::smem_u32(sW));
            const uint64_t d_dP = tc::make_desc_sw128(tc::smem_u32(sdP));
            const uint64_t d_Z = tc::make_desc_sw128(tc::smem_u32(sZ));
            const uint64_t d_dF2 = tc::make_desc_mn_b32(tc::smem_u32(sdF2), B2_BLK, 512);
            const uint64_t d_Phi = tc::make_desc_mn_b32(tc::smem_u32(sPhi), B2_BLK, 512);
            const int k1steps = (a.g + 7) / 8;
            // MMA-1: dPhi = dF W_tile^T -> D1 (straight-line issue; the unused k-steps are predicated off)
            auto mma1 = [&](int k) {
                tc::mbar_wait(w_full, k & 1);
                tc::tc_fence_after();
                tc::umma_tf32(tm_d1, d_dF, d_W, IDESC1, 0u);
                if (k1steps > 1) tc::umma_tf32(tm_d1, d_dF + 2, d_W + 2, IDESC1, 1u);
                if (k1steps > 2) tc::umma_tf32(tm_d1, d_dF + 4, d_W + 4, IDESC1, 1u);
                if (k1steps > 3) tc::umma_tf32(tm_d1, d_dF + 6, d_W + 6, IDESC1, 1u);
                tc::umma_commit(barA);
            };
            // MMA-2: gW tile il += Phi_tile^T dF (K = the 128 batch rows, 8 per instruction), accumulated over row tiles
            auto mma2 = [&](int k) {
                const int s = k & 1, il = k % n_loc, rl = k / n_loc;
                tc::mbar_wait(phi_full + s, (k >> 1) & 1);
                tc::tc_fence_after();
                const uint64_t dphi = d_Phi + (uint32_t)((s * 4 * B2_BLK) >> 4);
                const uint32_t dcol = tm_d2 + il * B2_NG;
#pragma unroll
                for (int k8 = 0; k8 < B2_BM / 8; ++k8)
                    tc::umma_tf32(dcol, dphi + 64 * k8, d_dF2 + 64 * k8, IDESC2, k8 != 0 ? 1u : (rl != 0 ? 1u : 0u));
                tc::umma_commit(s ? barC1 : barC);
            };
            int zph = 0;                                            // z_full phases consumed
            tc::mbar_wait(df_full, 0);
            mma1(0);
            int rows_done = 0;                                      // df_full phases consumed so far - 1
            for (int k = 0; k < T; ++k) {
                const int il = k % n_loc;
                const bool row_end = il == n_loc - 1;
                mma2(k);
                TLB(k, 6);
                // inside a row tile, MMA-1 of the next tile goes out as soon as the epilogue holds dPhi(k) in registers,
                // i.e. under the epilogue's math / stores; at a row end it has to wait for the next dF tile
                tc::mbar_wait(d1_free, k & 1);
                tc::tc_fence_after();
                if (!row_end && k + 1 < T) mma1(k + 1);
                TLB(k, 9);
                tc::mbar_wait(dp_full, k & 1);                      // dP(k) written
                TLB(k, 10);
                tc::tc_fence_after();
                if (NQ > 0) {
                    tc::mbar_wait(z_full, zph++ & 1);
                    tc::tc_fence_after();
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::umma_tf32(tm_d3, d_dP + (uint32_t)((b * B2_BLK + 32 * k4) >> 4),
                                          d_Z + (uint32_t)((b * 64 * 128 + 32 * k4) >> 4), IDESC3a, (b | k4) != 0 ? 1u : (il != 0 ? 1u : 0u));
                    if (nq1 > 0) {                                  // second pass: z rows 64.. through the same tile (producer reloads it)
                        tc::umma_commit(barZ);
                        tc::mbar_wait(z_full, zph++ & 1);
                        tc::tc_fence_after();
#pragma unroll
                        for (int b = 0; b < 2; ++b)
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::umma_tf32(tm_d3 + 64, d_dP + (uint32_t)((b * B2_BLK + 32 * k4) >> 4),
                                              d_Z + (uint32_t)((b * 64 * 128 + 32 * k4) >> 4), IDESC3b, (b | k4) != 0 ? 1u : (il != 0 ? 1u : 0u));
                    }
                }
                tc::umma_commit(barB);
                TLB(k, 11);
                if (row_end && k + 1 < T) {                         // the next row tile's dF must be staged first
                    ++rows_done;
                    tc::mbar_wait(df_full, rows_done & 1);
                    tc::tc_fence_after();
                    mma1(k + 1);
                }
            }
        }
        __syncwarp();
    } else if (warp == B2_EPI_WARPS + 1) {
        // ============================================ TMA PRODUCER ============================================
        // Every wait below is on a barrier whose NEXT phase needs a load this thread issues after the wait, so none of
        // them can run two phases ahead of it.
        if (T > 0 && tc::elect_one()) {
            const uint32_t phi_bytes = (rbf ? 4u : 2u) * B2_BLK;
            auto load_phi = [&](int k) {
                const int s = k & 1, c0 = (cs + (k % n_loc) * a.CS) * B2_BN, row0 = (rs + (k / n_loc) * a.RS) * B2_BM;
                tc::mbar_expect_tx(phi_full + s, phi_bytes);
                if (a.phi_blocked) {                                // tile-blocked layout (k1_fwd_tc2.cu): four contiguous 16 KB blocks
                    const int nb = rbf ? 4 : 2;
                    const int blk0 = ((rs + (k / n_loc) * a.RS) * n_ct + (cs + (k % n_loc) * a.CS)) * nb;
                    for (int b = 0; b < 2; ++b) {
                        tc::tma_load_3d(&map_cos, tc::smem_u32(sPhi + (s * 4 + b) * B2_BLK), phi_full + s, 0, (blk0 + b) * B2_BM, chain);
                        if (rbf) tc::tma_load_3d(&map_cos, tc::smem_u32(sPhi + (s * 4 + 2 + b) * B2_BLK), phi_full + s, 0, (blk0 + 2 + b) * B2_BM, chain);
                    }
                } else
                for (int b = 0; b < 2; ++b) {
                    tc::tma_load_3d(&map_cos, tc::smem_u32(sPhi + (s * 4 + b) * B2_BLK), phi_full + s, c0 + 32 * b, row0, chain);
                    if (rbf) tc::tma_load_3d(&map_sin, tc::smem_u32(sPhi + (s * 4 + 2 + b) * B2_BLK), phi_full + s, c0 + 32 * b, row0, chain);
                }
            };
            // DGPRF_BWD_PF=<n> (A/B switch, off by default): L2 prefetch of the saved-feature tile n tiles ahead of the ring.  The idea:
            // the two 64 KB stages keep ~64 KB in flight per SM and a tile takes ~4 k cycles to arrive, while the same load from
            // L2 takes < 1 k (profiles/r02_tma_load_rate.txt: 72-120 B / clk / SM for tiled and bulk loads alike).  Measured at
            // configs[4] scale: monotonically SLOWER, 507 -> 554 / 537 / 568 / 575 / 626 us per layer for n = 1 / 2 / 3 / 4 / 6
            // (profiles/r02_bwd_prefetch.txt) -- the prefetches queue in front of the W / z / Phi loads the next tile waits for.
            auto prefetch_phi = [&](int k) {
                if (k >= T) return;
                if (a.phi_blocked) {
                    const int nb = rbf ? 4 : 2;
                    const int blk0 = ((rs + (k / n_loc) * a.RS) * n_ct + (cs + (k % n_loc) * a.CS)) * nb;
                    for (int b = 0; b < nb; ++b) tc::tma_prefetch_3d(&map_cos, 0, (blk0 + b) * B2_BM, chain);
                } else {
                    const int c0 = (cs + (k % n_loc) * a.CS) * B2_BN, row0 = (rs + (k / n_loc) * a.RS) * B2_BM;
                    for (int b = 0; b < 2; ++b) {
                        tc::tma_prefetch_3d(&map_cos, c0 + 32 * b, row0, chain);
                        if (rbf) tc::tma_prefetch_3d(&map_sin, c0 + 32 * b, row0, chain);
                    }
                }
            };
            auto load_w = [&](int k) {
                const int c0 = (cs + (k % n_loc) * a.CS) * B2_BN;
                tc::mbar_expect_tx(w_full, (rbf ? 2u : 1u) * (B2_BN * 128));
                tc::tma_load_3d(&map_wp, tc::smem_u32(sW), w_full, 0, c0, chain);
                if (rbf) tc::tma_load_3d(&map_wp, tc::smem_u32(sW + B2_BN * 128), w_full, 0, a.M + c0, chain);
            };
            auto load_z = [&](int k, int pass) {                    // z rows [64 pass, 64 pass + nq0) of the column tile
                const int c0 = (cs + (k % n_loc) * a.CS) * B2_BN;
                tc::mbar_expect_tx(z_full, 2u * nq0 * 128);
                for (int b = 0; b < 2; ++b) tc::tma_load_3d(&map_z, tc::smem_u32(sZ + b * (64 * 128)), z_full, c0 + 32 * b, 64 * pass, zc);
            };
            load_phi(0);
            if (T > 1) load_phi(1);
            for (int i = 0; i < pf_dist; ++i) prefetch_phi(2 + i);
            load_w(0);
            if (NQ > 0) load_z(0, 0);
            for (int k = 0; k < T; ++k) {
                tc::mbar_wait(barA, k & 1);                         // MMA-1(k) done: the W tile is free
                TLB(k, 7);
                if (k + 1 < T) load_w(k + 1);
                if (NQ > 0 && nq1 > 0) {                            // hyper mode, > 64 input columns: second z pass of tile k
                    tc::mbar_wait(barZ, k & 1);
                    load_z(k, 1);
                }
                if (k + 2 < T) {                                    // refill stage k & 1 once MMA-2(k) and the epilogue's reads are done
                    tc::mbar_wait((k & 1) ? e_read1 : e_read, (k >> 1) & 1);
                    tc::mbar_wait((k & 1) ? barC1 : barC, (k >> 1) & 1);
                    load_phi(k + 2);
                    if (pf_dist > 0) prefetch_phi(k + 2 + pf_dist);
                }
                TLB(k, 8);
                if (NQ > 0 && k + 1 < T) {                          // MMA-3(k) done: the z tile is free for tile k+1
                    tc::mbar_wait(barB, k & 1);
                    load_z(k + 1, 0);
                }
            }
        }
        __syncwarp();
    } else {
        // ============================================ EPILOGUE ============================================
        const int lq = warp & 3, hh = warp >> 2;             // TMEM lane quarter, column group of the tile (B2_CW columns)
        const int pblk = (hh * B2_CW) >> 5, pch = ((hh * B2_CW) & 31) >> 2;    // 32-column Phi / dP block and first 16-byte chunk in it
        const int r = 32 * lq + lane;
        float rsum = 0.f;
        constexpr int ET = B2_EPI_WARPS * 32;
        // dF tile of a row tile (A of MMA-1, B of MMA-2): the partial slabs are summed in slab order (the order of
        // slab_load), one slab per pass so that a thread keeps 16 independent loads in flight
        auto stage_dF = [&](int row0) {
            constexpr int NE = B2_BM * 32 / ET;
            float v[NE];
            const float* base = a.dF.ptr + chain * a.dF.cs;
#pragma unroll
            for (int u = 0; u < NE; ++u) v[u] = 0.f;
            for (int sl = 0; sl < a.dF.n_slabs; ++sl) {
                const float* p = base + sl * a.dF.ss;
#pragma unroll
                for (int u = 0; u < NE; ++u) {
                    const int e = tid + u * ET;
                    const int rr = e >> 5, j = e & 31;
                    const int64_t row = row0 + rr;
                    const float x = (row < a.B && j < a.g) ? __ldg(p + row * a.dF.ld + j) : 0.f;
                    v[u] = sl == 0 ? x : v[u] + x;
                }
            }
#pragma unroll
            for (int u = 0; u < NE; ++u) {
                const int e = tid + u * ET;
                const int rr = e >> 5, j = e & 31;
                const float t = tc::to_tf32(v[u]);
                *reinterpret_cast<float*>(sdF + tc::sw128_off(rr, j)) = t;
                *reinterpret_cast<float*>(sdF2 + tc::sw128b32_off(rr, j)) = t;
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_b2(df_full);
        };
        if (T > 0) stage_dF(rs * B2_BM);

        for (int k = 0; k < T; ++k) {
            const int il = k % n_loc, rl = k / n_loc;
            const int row0 = (rs + rl * a.RS) * B2_BM;
            const int s = k & 1;
            const bool row_end = il == n_loc - 1;
            tc::mbar_wait(phi_full + s, (k >> 1) & 1);        // the TMA-written tile is read by every thread below
            if (tid == 0) TLB(k, 0);
            // ---- the thread's Phi values go to registers first so the ring stage can be refilled right away.  One
            //      release barrier per stage: its next phase needs the stage's next TMA load, which the issuer only
            //      requests after consuming this phase, so it can never run two phases ahead of the issuer's wait ----
            const uint8_t* ph = sPhi + s * 4 * B2_BLK;
            float4 pcv[B2_NCH], psv[B2_NCH];
            // In the 32-byte-atom layout the 16-byte chunk cc of row r sits at unit (cc/2) ^ (r & 3), half cc & 1: at a fixed
            // cc the 32 rows of a warp only touch 4 of the 8 16-byte bank groups.  Rows with (r >> 2) & 1 read the two
            // halves of each unit in the opposite order (then swap the registers back), which covers all 8 groups.
            const int hsw = (r >> 2) & 1;
#pragma unroll
            for (int cc = 0; cc < B2_NCH; ++cc) {
                pcv[cc] = *reinterpret_cast<const float4*>(ph + pblk * B2_BLK + tc::sw128b32_chunk(r, pch + (cc ^ hsw)));
                if (rbf) psv[cc] = *reinterpret_cast<const float4*>(ph + (2 + pblk) * B2_BLK + tc::sw128b32_chunk(r, pch + (cc ^ hsw)));
            }
            if (hsw) {
#pragma unroll
                for (int j = 0; j < B2_NCH; j += 2) {
                    const float4 t = pcv[j]; pcv[j] = pcv[j + 1]; pcv[j + 1] = t;
                    if (rbf) { const float4 u = psv[j]; psv[j] = psv[j + 1]; psv[j + 1] = u; }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_b2(s ? e_read1 : e_read);
            if (tid == 0) TLB(k, 1);
            tc::mbar_wait(barA, k & 1);                       // dPhi of tile k is in D1
            if (tid == 0) TLB(k, 2);
            tc::tc_fence_after();
            // ---- dPhi of the thread's row x 32 columns -> registers, then D1 is free for MMA-1 of the next tile ----
            float dcv[B2_CW], dsv[B2_CW];
            {
                const uint32_t lane_addr = (uint32_t)(32 * lq) << 16;
#pragma unroll
                for (int c16 = 0; c16 < B2_CW / 16; ++c16) {
                    tc::tmem_ld16(tm_d1 + lane_addr + B2_CW * hh + 16 * c16, dcv + 16 * c16);
                    if (rbf) tc::tmem_ld16(tm_d1 + lane_addr + B2_BN + B2_CW * hh + 16 * c16, dsv + 16 * c16);
                }
                tc::tmem_ld_wait();
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_b2(d1_free);
            if (tid == 0) TLB(k, 3);
            if (k > 0) tc::mbar_wait(barB, (k - 1) & 1);      // MMA-3(k-1) done: the dP tile may be overwritten
            if (tid == 0) TLB(k, 4);
            // ---- dP -> dP tile (A of MMA-3), row sums ----
#pragma unroll
            for (int cc = 0; cc < B2_NCH; ++cc) {             // 16-byte chunks of this thread inside its 32-wide block
                const float4 pc = pcv[cc];
                const float* dc = dcv + 4 * cc;
                float4 o;
                if (rbf) {
                    const float4 ps = psv[cc];
                    const float* ds = dsv + 4 * cc;
                    o.x = pc.x * ds[0] - ps.x * dc[0];
                    o.y = pc.y * ds[1] - ps.y * dc[1];
                    o.z = pc.z * ds[2] - ps.z * dc[2];
                    o.w = pc.w * ds[3] - ps.w * dc[3];
                } else {
                    o.x = pc.x > 0.f ? dc[0] * arc_scale : 0.f;
                    o.y = pc.y > 0.f ? dc[1] * arc_scale : 0.f;
                    o.z = pc.z > 0.f ? dc[2] * arc_scale : 0.f;
                    o.w = pc.w > 0.f ? dc[3] * arc_scale : 0.f;
                }
                rsum += (o.x + o.y) + (o.z + o.w);
                o.x = tc::tf32_rn_bits(o.x); o.y = tc::tf32_rn_bits(o.y); o.z = tc::tf32_rn_bits(o.z); o.w = tc::tf32_rn_bits(o.w);
                *reinterpret_cast<float4*>(sdP + pblk * B2_BLK + tc::sw128_chunk(r, pch + cc)) = o;
            }
            tc::tc_fence_before();
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_b2(dp_full);
            if (tid == 0) TLB(k, 5);
            if (row_end) {
                // ---- end of the row tile: dF_prev slab = s*T + mean*R, written once ----
                tc::mbar_wait(barB, k & 1);                    // T = sum over the row tile's column tiles is final (and the dP tile is idle)
                tc::tc_fence_after();
                R_s[hh * B2_BM + r] = rsum;
                rsum = 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(B2_EPI_WARPS * 32) : "memory");  // R_s complete
                if (warp < 4) {
                    const int64_t row = row0 + r;                          // warp < 4: r = 32 * warp + lane
                    float Rr = R_s[r];
#pragma unroll
                    for (int c = 1; c < B2_CQ; ++c) Rr += R_s[c * B2_BM + r];
                    if (hyper && row < a.B) a.Rpart[chain * a.r_cs + (int64_t)cs * a.B + row] = Rr;
                    for (int qc = 0; qc < NQ / 16; ++qc) {
                        float t[16];
                        tc::tmem_ld16(tm_d3 + ((uint32_t)(32 * warp) << 16) + 16 * qc, t);
                        tc::tmem_ld_wait();
                        if (row < a.B) {
                            float* dst = a.Dpart != nullptr ? a.Dpart + chain * a.d_cs + ((int64_t)cs * a.B + row) * a.d_prev : nullptr;
                            float* tdst = hyper ? a.Tpart + chain * a.t_cs + ((int64_t)cs * a.B + row) * a.d : nullptr;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int q = 16 * qc + i;
                                if (dst != nullptr && q < a.d_prev) {
                                    float v = s_s[q] * t[i];
                                    if (a.has_mean) v = fmaf(m_s[q], Rr, v);
                                    dst[q] = v;
                                }
                                if (tdst != nullptr && q < a.d) tdst[q] = t[i];     // raw T for the hyper-gradient reduction
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(B2_EPI_WARPS * 32) : "memory");  // R_s and D3 are read: the next row tile may reuse them
                if (a.Dsum != nullptr) {
                    // the LAST column split of this row tile to get here adds the CS slabs of dU/dF_{l-1} (slab order, the order of
                    // k_sum_slabs: bit-identical) into the dense dF the layer below reads -- a ticket per (chain, row tile)
                    // instead of a ~9-17 us launch between two layers; the ticket resets itself for the next launch
                    uint32_t* last_s = reinterpret_cast<uint32_t*>(bars + 16);
                    const int rt = rs + rl * a.RS;
                    if (tid == 0) {
                        __threadfence();                 // the slab rows (written before the barrier above) are ordered before the ticket
                        unsigned int* ctr = a.sum_ctr + (int64_t)chain * n_rt + rt;
                        const unsigned int old = atomicAdd(ctr, 1u);
                        const bool last = old + 1 == (unsigned int)(a.CS < n_ct ? a.CS : n_ct);
                        if (last) atomicExch(ctr, 0u);
                        *last_s = last ? 1u : 0u;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(B2_EPI_WARPS * 32) : "memory");
                    if (*last_s != 0u) {
                        __threadfence();
                        const int rows = min(B2_BM, a.B - row0);
                        const int64_t ss = (int64_t)a.B * a.d_prev;
                        const float* src = a.Dpart + chain * a.d_cs + (int64_t)row0 * a.d_prev;
                        float* dst = a.Dsum + chain * a.dsum_cs + (int64_t)row0 * a.d_prev;
                        const int n = rows * a.d_prev;
                        int e_lo = 0;
                        if (((ss | a.d_cs | a.dsum_cs) & 3) == 0 && (((int64_t)row0 * a.d_prev) & 3) == 0) {      // 128-bit lanes
                            const int nv = n >> 2;
                            for (int e = tid; e < nv; e += ET) {
                                float4 acc = __ldcg(reinterpret_cast<const float4*>(src) + e);
                                for (int sl = 1; sl < a.CS; ++sl) {
                                    const float4 t = __ldcg(reinterpret_cast<const float4*>(src + sl * ss) + e);
                                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                                }
                                reinterpret_cast<float4*>(dst)[e] = acc;
                            }
                            e_lo = nv << 2;
                        }
                        for (int e = e_lo + tid; e < n; e += ET) {
                            float acc = __ldcg(src + e);
                            for (int sl = 1; sl < a.CS; ++sl) acc += __ldcg(src + sl * ss + e);
                            dst[e] = acc;
                        }
                    }
                    // (bars + 16 is rewritten only after the next row end's first bar.sync: every reader is past it by then)
                }
                if (k + 1 < T) {
                    // its dF tile may be staged once MMA-2(k), the last reader of the old one, is done
                    tc::mbar_wait((k & 1) ? barC1 : barC, (k >> 1) & 1);
                    stage_dF((rs + (rl + 1) * a.RS) * B2_BM);
                }
            }
        }
        if (T > 0) tc::mbar_wait(((T - 1) & 1) ? barC1 : barC, ((T - 1) >> 1) & 1);   // the last MMA-2: every gW tile is final
        tc::tc_fence_after();
        // ---- the CTA's gW tiles -> row-split slab rs (zeros if the CTA had no row tile).  A tile goes TMEM -> shared
        //      memory (the dP tile is idle now) -> global, so that the slab is written with coalesced stores: the
        //      [64 features x g] blocks of the cos and of the sin half are contiguous in W's layout ----
        {
            float* stg = reinterpret_cast<float*>(sdP);                    // [128][33]
            const int fl = 32 * (warp & 3) + lane;                         // feature row inside [cos 64 | sin 64]
            for (int i = 0; i < n_loc; ++i) {
                if (warp < 4) {
#pragma unroll
                    for (int c16 = 0; c16 < B2_NG / 16; ++c16) {
                        float v[16];
                        tc::tmem_ld16(tm_d2 + i * B2_NG + ((uint32_t)(32 * warp) << 16) + 16 * c16, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) stg[fl * 33 + 16 * c16 + j] = T > 0 ? v[j] : 0.f;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(B2_EPI_WARPS * 32) : "memory");
                const int c0 = (cs + i * a.CS) * B2_BN;
                const int ncol = min(B2_BN, a.M - c0);                     // live feature columns of this tile
                float* base = a.gWpart + chain * a.gw_cs + (int64_t)rs * a.gw_ss;
                for (int half = 0; half < (rbf ? 2 : 1); ++half) {
                    float* dst = base + ((int64_t)(half ? a.M : 0) + c0) * a.g;
                    for (int e = tid; e < ncol * a.g; e += ET) {
                        const int f = e / a.g, j = e - f * a.g;
                        dst[e] = stg[(half * B2_BN + f) * 33 + j];
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(B2_EPI_WARPS * 32) : "memory");
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == B2_EPI_WARPS) tc::tmem_dealloc(tmem_base, B2_TMEM_COLS);
}

// W [F, g] -> wp [F][32]: zero-padded, tf32-rounded rows (a TMA row stride must be a multiple of 16 bytes)
__global__ void __launch_bounds__(256)
k_prep_bwd_tc2(const float* __restrict__ W, int64_t w_cs, int F, int g, float* __restrict__ wp) {
    const int chain = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)F * B2_NG) return;
    const int64_t f = i >> 5;
    const int j = (int)(i & 31);
    wp[(int64_t)chain * F * B2_NG + i] = j < g ? tc::to_tf32(__ldg(W + chain * w_cs + f * g + j)) : 0.f;
}

static constexpr size_t kB2Smem = 1024 + B2_HDR + 8 * (size_t)B2_BLK + 2 * (size_t)B2_BLK + 2 * (size_t)B2_BLK +
                                  (size_t)B2_BLK + 2 * 64 * 128;

bool dgprf_bwd_tc2_shape_ok(int M, int g, int d, int d_prev, int CS, int hyper) {
    // a row split beyond the last row tile is fine (its CTA writes a zero gW slab); every column split must own a
    // tile because it owns a dF_prev slab.  Hyper mode keeps T for all d <= 128 input columns in TMEM (128 columns),
    // which leaves room for 8 resident gW tiles instead of 10.
    const int n_ct = ceil_div(M, B2_BN);
    const int max_loc = hyper ? 8 : B2_MAX_LOC;
    return (M % 4) == 0 && g <= B2_NG && d_prev <= 64 && (!hyper || d <= 128) && CS <= n_ct && ceil_div(n_ct, CS) <= max_loc &&
           getenv("DGPRF_NO_TC2") == nullptr;
}
int64_t dgprf_bwd_tc2_wp_floats(int F) { return (int64_t)F * B2_NG; }

// Column splits for the pipelined backward (0: shape not supported).  A CTA (row split, column split, chain) pays a fixed
// set-up (TMEM allocation, barrier init, first dF tile, pipeline fill; about three tile times) and then walks
// ceil(n_ct / CS) x ceil(n_rt / RS) tiles, so with few row tiles per chain a small CS (more tiles per CTA, fewer CTAs)
// beats the maximal split: pick the CS with the smallest  waves x (3 + tiles per CTA).
int dgprf_bwd_tc2_pick_cs(int B, int M, int g, int d, int d_prev, int RS, int n_chains, int hyper) {
    const int n_ct = ceil_div(M, B2_BN), n_rt = ceil_div(B, B2_BM);
    const int rs_act = RS < n_rt ? RS : n_rt;
    const int n_rloc = ceil_div(n_rt, rs_act);
    int best = 0;
    int64_t best_cost = 0;
    for (int cs = 1; cs <= kMaxCS && cs <= n_ct; ++cs) {
        if (!dgprf_bwd_tc2_shape_ok(M, g, d, d_prev, cs, hyper)) continue;
        const int64_t ctas = (int64_t)rs_act * cs * n_chains;
        const int64_t cost = ceil_div(ctas, (int64_t)148) * (3 + (int64_t)ceil_div(n_ct, cs) * n_rloc);
        if (best == 0 || cost < best_cost || (cost == best_cost && cs > best)) { best = cs; best_cost = cost; }
    }
    return best;
}

bool dgprf_bwd_tc2_supported(const BwdArgs& a) {
    return a.wp != nullptr && (a.phi_cs % 4) == 0 && dgprf_bwd_tc2_shape_ok(a.M, a.g, a.d, a.d_prev, a.CS, a.hyper);
}

int dgprf_launch_bwd_tc2(const BwdArgs& a, int n_chains, cudaStream_t st) {
    static_assert(kB2Smem <= 232448, "shared memory budget");
    { const int rc_s = dgprf_ensure_smem((const void*)k2_bwd_tc2, (size_t)kB2Smem); if (rc_s) return rc_s; }
    if (!a.prepped) {
        ProfScope _ps("k_prep_bwd_tc2", st);
        k_prep_bwd_tc2<<<dim3(ceil_div(a.F * B2_NG, 256), n_chains), 256, 0, st>>>(a.W, a.w_cs, a.F, a.g, a.wp);
        DGPRF_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap mc, ms, mw, mz;
    memset(&ms, 0, sizeof(ms));
    memset(&mz, 0, sizeof(mz));
    int rc;
    if (a.phi_blocked) {
        const int64_t n = dgprf_phi_blocked_floats(a.B, a.M, a.kind);
        DGPRF_REQUIRE(a.phi_cs >= n, "blocked saved-feature buffer too small: %lld < %lld floats", (long long)a.phi_cs, (long long)n);
        rc = dgprf_make_tmap_3d(&mc, a.Phi, 32, (uint64_t)(n / 32), n_chains, 32, a.phi_cs, B2_BM, true);
        if (rc) return rc;
    } else {
        rc = dgprf_make_tmap_3d(&mc, a.Phi, a.M, a.B, n_chains, a.F, a.phi_cs, B2_BM, true);
        if (rc) return rc;
        if (a.kind == DGPRF_KIND_RBF) {
            rc = dgprf_make_tmap_3d(&ms, a.Phi + a.M, a.M, a.B, n_chains, a.F, a.phi_cs, B2_BM, true);
            if (rc) return rc;
        }
    }
    rc = dgprf_make_tmap_3d(&mw, a.wp, B2_NG, a.F, n_chains, B2_NG, (uint64_t)a.F * B2_NG, B2_BN);
    if (rc) return rc;
    const int NQ = a.hyper ? ((a.d + 15) & ~15) : (a.d_prev > 0 ? ((a.d_prev + 15) & ~15) : 0);
    if (NQ > 0) {
        rc = dgprf_make_tmap_3d(&mz, a.z, a.M, a.d, a.z_cs != 0 ? n_chains : 1, a.M, a.z_cs, NQ < 64 ? NQ : 64);
        if (rc) return rc;
    }
    dim3 grid(a.RS, a.CS, n_chains);
    static long long* tl = nullptr;                       // debug timeline (DGPRF_BWD_TIMELINE=<call number>)
    static int tl_calls = 0;
    if (getenv("DGPRF_BWD_TIMELINE") && !tl) cudaMalloc(&tl, 16 * 12 * sizeof(long long));
    if (tl) cudaMemsetAsync(tl, 0, 16 * 12 * sizeof(long long), st);
    { ProfScope _ps("k2_bwd_tc2", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k2_bwd_tc2, grid, dim3(B2_THREADS), kB2Smem, st, a, getenv("DGPRF_BWD_PF") ? atoi(getenv("DGPRF_BWD_PF")) : 0, tl, mc, ms, mw, mz)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    if (tl && ++tl_calls == atoi(getenv("DGPRF_BWD_TIMELINE"))) {
        long long h[16 * 12];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, tl, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        fprintf(stderr, "bwd2 timeline (cycles since the epilogue saw tile 0)\n tile | epi: phi-ready Phi->regs dPhi-ready D1->regs dP-free dP-stored | issuer: mma2-issued mma1-done loads-issued mma1(k+1)-issued dP-seen mma3-issued\n");
        for (int t = 0; t < 16; ++t) {
            fprintf(stderr, " %3d |", t);
            for (int e = 0; e < 12; ++e) fprintf(stderr, " %7lld%s", h[t * 12 + e] ? h[t * 12 + e] - t0 : -1LL, e == 5 ? " |" : "");
            fprintf(stderr, "\n");
        }
    }
    return DGPRF_OK;
}
