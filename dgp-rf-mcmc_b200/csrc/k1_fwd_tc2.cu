// K1 (tensor-core, DGPRF_PREC_TF32): warp-specialised, pipelined fused [RF layer -> GP layer] forward (any layer with at
// least four 128 x 64 tiles; input width <= 128, or any width in the WIDE variant).  Same arithmetic as k1_fwd_simt.cu; the
// execution model:
//
//   * the A operand of GEMM #1, (in * exp(log_inv_ls)) split into tf32 hi + lo, is written ONCE per CTA into
//     TENSOR MEMORY (tcgen05.st) and consumed from there by every column tile (tcgen05.mma with A in TMEM),
//     so shared memory only holds the streamed operands;
//   * roles:  warps 0-15 epilogue   tcgen05.ld P -> sincos/relu -> Phi tile (smem, UMMA/TMA swizzle)
//             warp  16   MMA-1      one elected thread issues GEMM #1 (3xTF32 phases) of the tiles, running ahead
//             warp  17   producer   one thread issues the TMA loads of the next z tile (tf32 hi/lo) and W tile
//             warp  18   store      TMA store of the Phi tile (saved features)
//             warp  19   MMA-2      one elected thread issues GEMM #2 (Phi.W) of each tile as soon as its Phi is in smem
//     (two issuing threads: the tensor pipe runs about as fast as one thread can feed it, so with a single issuer
//      the wait/poll code between the two GEMMs of a tile left it idle a third of the time)
//     connected by mbarriers; P (D1) is double-buffered in TMEM and the z tile is a 2-stage ring (when it
//     fits), so GEMM #1 of tile t+1 and the loads of tile t+2 run under the epilogue of tile t.
//   * WIDE variant (input width > 128, e.g. the MNIST-shaped configs): the A operand no longer fits in tensor memory, so
//     GEMM #1 becomes a classic K loop in SS mode: a TMA ring of k-blocks [A hi | A lo | Omega hi | Omega lo] where A is the
//     raw layer input and Omega = exp(log_inv_ls) * z + mean (both split into tf32 hi/lo by prep kernels), so neither the
//     per-row scaling nor the mean bias is needed in the kernel; epilogue, GEMM #2 and the Phi store are unchanged;
//   * the streamed operands are pre-laid for TMA once per step for all layers (k_prep_layers.cu; the prep kernels in this
//     file serve stand-alone launches): z^T split into tf32 hi/lo, K-major, zero-padded to 128 K columns ([2][M][128]) and
//     W^T rounded to tf32 ([NG][F]); debug: DGPRF_TC2_TIMELINE=<call number> prints per-role clock stamps of one CTA.
#include <stdio.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "tc_common.cuh"

constexpr int V2_BM = 128, V2_BN = 64;
constexpr int V2_EPI_WARPS = 16;                 // 4 per TMEM lane quarter, 16 tile columns each
constexpr int V2_THREADS = (V2_EPI_WARPS + 4) * 32;                           // 640
constexpr int V2_HDR = 1024;                      // bias row + mbarriers + TMEM slot
constexpr int V2_BLK = V2_BM * 128;               // [128 x 32 tf32] block
constexpr int V2_BBLK = V2_BN * 128;              // [64 x 32 tf32] block
constexpr int V2_KB = 4;                          // K blocks of GEMM #1 (input width <= 128)
constexpr uint32_t V2_TMEM_COLS = 512;            // A_hi 128 | A_lo 128 | D1 2x64 | D2 64   (WIDE: D1 2x64 | D2 64 only)
constexpr int V2_RING = 2 * V2_BLK + 2 * V2_BBLK; // WIDE k-block slot: A hi 16K | A lo 16K | Omega hi 8K | Omega lo 8K

namespace tc {
// 32-byte store (sm_100+): one full sector per lane
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
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

// Straight-line MMA issue (one elected thread).  Every descriptor is base + compile-time offset and the accumulate
// flags are constants, so the SASS is a run of UTCHMMA with independent 64-bit adds: with runtime k loops the
// uniform-datapath loop / descriptor arithmetic costs ~65 cycles per k-step and bounds the whole kernel.
// All NKB x 4 k-steps are issued: the A operand and the z tile are zero-padded beyond the input width.
template <int NKB>
__device__ __forceinline__ void issue_gemm1(uint32_t dcol, uint32_t tm_ahi, uint32_t tm_alo, uint64_t db, uint32_t idesc) {
    constexpr uint32_t lo_off = (uint32_t)(NKB * V2_BBLK) >> 4;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acol = 32 * kb + 8 * k4;
            const uint64_t dbh = db + (uint32_t)((kb * V2_BBLK + k4 * 32) >> 4);
            tc::umma_tf32_ts(dcol, tm_alo + acol, dbh, idesc, (kb | k4) != 0 ? 1u : 0u);
            tc::umma_tf32_ts(dcol, tm_ahi + acol, dbh + lo_off, idesc, 1u);
            tc::umma_tf32_ts(dcol, tm_ahi + acol, dbh, idesc, 1u);
        }
}
template <int NB2, int NG>
__device__ __forceinline__ void issue_gemm2(uint32_t tm_d2, uint64_t dphi, uint64_t dw, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
    for (int b = 0; b < NB2; ++b)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
            tc::umma_tf32(tm_d2, dphi + (uint32_t)((b * V2_BLK + k4 * 32) >> 4), dw + (uint32_t)((b * NG * 128 + k4 * 32) >> 4), idesc,
                          (b | k4) != 0 ? 1u : acc_first);
}

// GEMM #2 of one Phi sub-tile (32 P columns: [cos block | sin block]); dw_h = descriptor of W^T block h of the tile's stage
template <int NBLK, int NG>
__device__ __forceinline__ void issue_gemm2_sub(uint32_t tm_d2, uint64_t dphi, uint64_t dw_h, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
    for (int b = 0; b < NBLK; ++b)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
            tc::umma_tf32(tm_d2, dphi + (uint32_t)((b * V2_BLK + k4 * 32) >> 4), dw_h + (uint32_t)((2 * b * NG * 128 + k4 * 32) >> 4), idesc,
                          (b | k4) != 0 ? 1u : acc_first);
}

template <int NG, bool WIDE>
__global__ void __launch_bounds__(V2_THREADS, 1)
k1_fwd_tc2(const FwdArgs a, const int NS1, const int NSW_, long long* const tl, const __grid_constant__ CUtensorMap map_cos,
           const __grid_constant__ CUtensorMap map_sin, const __grid_constant__ CUtensorMap map_zt, const __grid_constant__ CUtensorMap map_wt,
           const __grid_constant__ CUtensorMap map_at) {
    dgprf_pdl_sync();
    // NSW_ bit 8 (DGPRF_TC2_DIRECT_STORE=1, an experiment kept for A/B runs): the epilogue warps store the saved features
    // straight from registers (st.global.v4, 64 B per row and half) instead of the store warp's TMA store of the shared
    // tile.  Measured at configs[4] layer scale: 1.39 ms against 0.65 ms -- a warp-wide store of 32 rows 32 KB apart is 32
    // separate requests per instruction and partial-sector writes; the coalesced TMA store stays the default.
    // NSW_ bit 9: TWO Phi tiles in shared memory (the launcher then gives the z ring one stage instead of two): the TMA store
    // of tile t reads its tile while the epilogue writes tile t+1 into the other one.  With one tile the period of a tile is
    // "store reads 64 KB" + "epilogue writes 64 KB" back to back (profiles/r02_summary.md section B).
    const int NSW = NSW_ & 0xff;
    const bool direct_store = (NSW_ & 0x100) != 0 && !a.phi_blocked;      // (the experiment writes the row-major layout)
    // NSW_ bit 11 (DGPRF_TC2_REG_STORE=1, an experiment kept for A/B runs; tile-blocked layout only): the epilogue warps store
    // the saved features from registers with 256-bit stores.  In the blocked layout a thread's 16 columns are 64 contiguous
    // bytes of a 128-byte row and a warp's 32 rows are consecutive, so a warp covers a dense 4 KB region per block and the
    // Phi tile in shared memory is free as soon as GEMM #2 retires.  Measured at configs[4] layer scale: 0.78 ms against
    // 0.635 ms with the TMA store -- every store instruction is still 32 separate 32-byte requests and the epilogue warps,
    // which are the critical path, stall on the LSU queue; the asynchronous TMA store stays the default.
    const bool reg_store = (NSW_ & 0x800) != 0 && a.phi_blocked && a.Phi != nullptr;
    const int NPHI = (NSW_ & 0x200) ? 2 : 1;
    // NSW_ bit 10 (non-WIDE): the z operand streams through a ring of NS1 k-block slots [hi 8 KB | lo 8 KB] instead of whole
    // 64 KB tiles, which is what frees the shared memory for the second Phi tile
    const bool KRING = !WIDE && (NSW_ & 0x400) != 0;
    // NSW_ bit 12 (with the k-block ring; the default when the features are stored): the Phi tile is produced, consumed and
    // stored as two SUB-TILES of 32 P columns ([cos block | sin block], 32 KB) through a ring of THREE sub-tile buffers.  The
    // TMA store of a sub-tile then drains while the epilogue writes the next two (a whole second tile would not fit next to
    // a z ring deep enough to keep GEMM #1 ahead: 96 KB of Phi buffers + six k-block slots instead of 128 KB + four).
    const bool SUB3 = KRING && (NSW_ & 0x1000) != 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    float* bias_s = reinterpret_cast<float*>(sm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + V2_BM * sizeof(float));
    uint8_t* sPhi = sm + V2_HDR;                         // 4 blocks [128 x 32]: cos 0,1 | sin 2,3
    uint8_t* sW = sPhi + (SUB3 ? 6 : NPHI * 4) * V2_BLK; // NSW stages x 4 blocks [NG x 32]
    uint8_t* sB1 = sW + NSW * 4 * NG * 128;              // NS1 stages x [hi blocks 0..n_kb) | lo blocks 0..n_kb)] of [64 x 32]
                                                         // WIDE: NS1 k-block slots of V2_RING bytes
    uint64_t* b1_full = bars + 22;    // [6] TMA complete_tx      -> MMA
    uint64_t* b1_empty = bars + 28;   // [6] MMA commit           -> producer
    uint64_t* d1_full = bars + 6;     // [2] MMA commit           -> epilogue
    uint64_t* d1_empty = bars + 8;    // [2] epilogue warps       -> MMA
    uint64_t* phi_full = SUB3 ? bars + 0 : bars + 18;    // [2 | 3] epilogue warps           -> MMA, store
    uint64_t* phi_empty = SUB3 ? bars + 3 : bars + 20;   // [2 | 3] MMA commit + store warp  -> epilogue      (count 2)
    uint64_t* w_full = bars + 12;     // [2] TMA complete_tx      -> MMA
    uint64_t* w_empty = bars + 14;    // [2] MMA commit           -> producer
    uint64_t* d2_full = bars + 16;    // MMA commit               -> final epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.z, cs = blockIdx.y, row0 = blockIdx.x * V2_BM;
    // debug timeline (DGPRF_TC2_TIMELINE=1): clock stamps of CTA 0, first 16 tiles, 12 events per tile
    const unsigned tl_rb = (unsigned)NSW_ >> 16;         // row block whose CTA is stamped (DGPRF_TC2_TIMELINE_RB, default 0)
#define TL(t, ev) do { if (tl != nullptr && blockIdx.x == tl_rb && blockIdx.y == 0 && blockIdx.z == 0 && (t) < 16) tl[(t) * 12 + (ev)] = clock64(); } while (0)
    // whole-CTA phase stamps (thread 0): start | barriers + TMEM ready | input tile in smem | A operand in TMEM | last tile's epilogue | end
#define TLC(ev) do { if (tl != nullptr && tid == 0 && blockIdx.x == tl_rb && blockIdx.y == 0 && blockIdx.z == 0) tl[16 * 12 + (ev)] = clock64(); } while (0)
    TLC(0);
    const float* X = a.X + chain * a.x_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float scale = (rbf ? 1.f : 1.41421356237f) * __expf(__ldg(a.log_amp + chain * a.h_cs)) * rsqrtf((float)a.M);
    const int n_ct = (a.M + V2_BN - 1) / V2_BN;
    // column split cs owns the CONTIGUOUS tiles [cs * per, cs * per + n_my): consecutive tiles of a CTA then store
    // adjacent 256-byte pieces of the same Phi rows
    const int per = (n_ct + a.CS - 1) / a.CS;
    const int ct0 = cs * per;
    const int n_my = ct0 < n_ct ? min(per, n_ct - ct0) : 0;
    const int n_kb = (a.d + 31) / 32;
    const int nb2 = rbf ? 4 : 2;
    const uint32_t b1_stage = WIDE ? (uint32_t)V2_RING : (KRING ? 2u * V2_BBLK : 2u * n_kb * V2_BBLK);

    if (warp == V2_EPI_WARPS) tc::tmem_alloc(tmem_slot, V2_TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) {
            tc::mbar_init(b1_full + i, 1);
            tc::mbar_init(b1_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(d1_full + i, 1);
            tc::mbar_init(d1_empty + i, V2_EPI_WARPS);
            tc::mbar_init(w_full + i, 1);
            tc::mbar_init(w_empty + i, 1);
        }
        for (int i = 0; i < (SUB3 ? 3 : 2); ++i) {
            tc::mbar_init(phi_full + i, V2_EPI_WARPS);
            tc::mbar_init(phi_empty + i, 2);
        }
        tc::mbar_init(d2_full, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    TLC(1);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_ahi = tmem_base, tm_alo = tmem_base + 128, tm_d1 = tmem_base + (WIDE ? 0 : 256), tm_d2 = tmem_base + (WIDE ? 128 : 384);

    // ---- A operand -> TMEM.  The 128-row input tile is first read coalesced into shared memory (it borrows the
    //      Phi / W regions, idle until the pipeline starts), then each epilogue thread owns one row x 64 K columns.
    if (!WIDE) {
        float* sIn = reinterpret_cast<float*>(sPhi);          // [128][129]
        float* sSq = reinterpret_cast<float*>(sB1);           // [128] exp(log_inv_ls) | [128] mean
        float* sMean = sSq + 128;
        if (tid < 128) {
            sSq[tid] = tid < a.d ? expf(__ldg(ls + tid)) : 0.f;
            sMean[tid] = (mean != nullptr && tid < a.d) ? __ldg(mean + tid) : 0.f;
        }
        // The whole tile in ONE round of loads (26 per thread in flight; the partial slabs of the previous layer are added
        // slab by slab in slab order, the order of slab_load): under the other CTAs' store streams a dependent round trip to
        // L2 / HBM costs ~3.4 k cycles, and four rounds of 8 loads were 7 % of a CTA's lifetime at configs[4] layer scale
        constexpr int PB = (V2_BM * 128 + V2_THREADS - 1) / V2_THREADS;
        const float* fp = a.Fprev.ptr + chain * a.Fprev.cs;
        for (int e0 = tid; e0 < V2_BM * 128; e0 += V2_THREADS * PB) {
            float v[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int e = e0 + u * V2_THREADS;
                const int r = e >> 7, q = e & 127;
                const int64_t row = row0 + r;
                v[u] = (e < V2_BM * 128 && row < a.B && q >= a.d_prev && q < a.d) ? __ldg(X + row * a.ldx + (q - a.d_prev)) : 0.f;
            }
            for (int sl = 0; sl < a.Fprev.n_slabs; ++sl) {
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int e = e0 + u * V2_THREADS;
                    const int r = e >> 7, q = e & 127;
                    const int64_t row = row0 + r;
                    if (e < V2_BM * 128 && row < a.B && q < a.d_prev) v[u] += __ldg(fp + sl * a.Fprev.ss + row * a.Fprev.ld + q);
                }
            }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int e = e0 + u * V2_THREADS;
                if (e < V2_BM * 128) sIn[(e >> 7) * 129 + (e & 127)] = v[u];
            }
        }
        __syncthreads();
        TLC(2);
        if (warp < V2_EPI_WARPS) {
            const int lq = warp & 3, kq = warp >> 2;              // TMEM lane quarter, 32-column K quarter
            const int r = 32 * lq + lane;
            float* bpart = reinterpret_cast<float*>(sW) + 1024;  // [4][128] partial biases (sW is idle until the pipeline starts)
            float bsum = 0.f;
#pragma unroll
            for (int c16 = 0; c16 < 2; ++c16) {
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int q = 32 * kq + 16 * c16 + i;
                    const float v = sIn[r * 129 + q];
                    bsum = fmaf(v, sMean[q], bsum);
                    const float x = v * sSq[q];
                    hi[i] = tc::to_tf32(x);
                    lo[i] = tc::to_tf32(x - hi[i]);
                }
                const uint32_t col = 32 * kq + 16 * c16;
                tc::tmem_st16(tm_ahi + ((uint32_t)(32 * lq) << 16) + col, hi);
                tc::tmem_st16(tm_alo + ((uint32_t)(32 * lq) << 16) + col, lo);
            }
            tc::tmem_st_wait();
            if (a.has_mean) {                     // bias_r = sum_q in[r][q] mean[q]: the four K quarters of a row, fixed order
                bpart[kq * V2_BM + r] = bsum;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (kq == 0) bias_s[r] = ((bpart[r] + bpart[V2_BM + r]) + bpart[2 * V2_BM + r]) + bpart[3 * V2_BM + r];
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    TLC(3);

    constexpr uint32_t IDESC1 = tc::make_idesc_tf32(V2_BM, V2_BN);
    constexpr uint32_t IDESC2 = tc::make_idesc_tf32(V2_BM, NG);

    if (warp < V2_EPI_WARPS) {
        // ===================================== EPILOGUE =====================================
        const int lq = warp & 3, cq = warp >> 2;               // TMEM lane quarter, 16-column quarter of the tile
        const int r = 32 * lq + lane;
        const float bias = (!WIDE && a.has_mean) ? bias_s[r] : 0.f;     // WIDE: the mean is folded into Omega
        const int pb = cq >> 1, pc = (cq & 1) * 4;             // 32-column Phi block and first 16-byte chunk inside it
        for (int t = 0; t < n_my; ++t) {
            const int c0 = (ct0 + t) * V2_BN;
            const int buf = t & 1;
            tc::mbar_wait(d1_full + buf, (t >> 1) & 1);
            if (tid == 0) TL(t, 0);
            tc::tc_fence_after();
            if (SUB3) {
                // this warp's 8 columns of each 32-column half of the P tile -> registers, then D1 is free
                float ph[2][8];
                tc::tmem_ld8(tm_d1 + 64 * buf + ((uint32_t)(32 * lq) << 16) + 8 * cq, ph[0]);
                tc::tmem_ld8(tm_d1 + 64 * buf + ((uint32_t)(32 * lq) << 16) + 32 + 8 * cq, ph[1]);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(d1_empty + buf);
                if (tid == 0) TL(t, 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float c8[8], s8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool live = (c0 + 32 * h + 8 * cq + i) < a.M;
                        const float x = ph[h][i] + bias;
                        if (rbf) {
                            float sn, cs;
                            sincos_cw(x, &sn, &cs);
                            c8[i] = live ? scale * cs : 0.f;
                            s8[i] = live ? scale * sn : 0.f;
                        } else {
                            c8[i] = live ? scale * fmaxf(x, 0.f) : 0.f;
                        }
                    }
                    if (tid == 0 && h == 1) TL(t, 2);
                    const int u = 2 * t + h, q = u % 3;
                    uint8_t* sub = sPhi + q * 2 * V2_BLK;             // [cos block | sin block] of sub-tile u
                    tc::mbar_wait(phi_empty + q, ((u / 3) & 1) ^ 1);  // GEMM #2 and the store of the sub-tile last held here are done
                    if (tid == 0 && h == 1) TL(t, 3);
#pragma unroll
                    for (int c4 = 0; c4 < 2; ++c4) {
                        *reinterpret_cast<float4*>(sub + tc::sw128_chunk(r, 2 * cq + c4)) =
                            make_float4(c8[4 * c4], c8[4 * c4 + 1], c8[4 * c4 + 2], c8[4 * c4 + 3]);
                        if (rbf)
                            *reinterpret_cast<float4*>(sub + V2_BLK + tc::sw128_chunk(r, 2 * cq + c4)) =
                                make_float4(s8[4 * c4], s8[4 * c4 + 1], s8[4 * c4 + 2], s8[4 * c4 + 3]);
                    }
                    tc::fence_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(phi_full + q);
                }
                if (tid == 0) TL(t, 4);
                continue;
            }
            float p[16];
            tc::tmem_ld16(tm_d1 + 64 * buf + ((uint32_t)(32 * lq) << 16) + 16 * cq, p);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(d1_empty + buf);          // P is in registers: the buffer is free
            if (tid == 0) TL(t, 1);
            float f1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bool live = (c0 + 16 * cq + i) < a.M;
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
            if (tid == 0) TL(t, 2);
            if (direct_store && a.Phi != nullptr && row0 + r < a.B) {
                // saved features from registers: 64 contiguous bytes per row and half (cos | sin); the four column-quarter
                // warps of a lane quarter fill 256 contiguous bytes of the row between them (merged in L2)
                float* dst = a.Phi + chain * a.phi_cs + (int64_t)(row0 + r) * a.F + c0 + 16 * cq;
                if (c0 + 16 * cq + 16 <= a.M) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(p[4 * c4], p[4 * c4 + 1], p[4 * c4 + 2], p[4 * c4 + 3]);
                        if (rbf)
                            *reinterpret_cast<float4*>(dst + a.M + 4 * c4) =
                                make_float4(f1[4 * c4], f1[4 * c4 + 1], f1[4 * c4 + 2], f1[4 * c4 + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + 16 * cq + i < a.M) {
                            dst[i] = p[i];
                            if (rbf) dst[a.M + i] = f1[i];
                        }
                }
            }
            if (reg_store) {
                const int nb = rbf ? 4 : 2;
                const int64_t blk0 = ((int64_t)blockIdx.x * n_ct + (ct0 + t)) * nb;
                float* dst = a.Phi + chain * a.phi_cs + ((blk0 + pb) * V2_BM + r) * 32 + (cq & 1) * 16;
                tc::st_global_v8(dst, p);
                tc::st_global_v8(dst + 8, p + 8);
                if (rbf) {
                    tc::st_global_v8(dst + 2 * V2_BM * 32, f1);
                    tc::st_global_v8(dst + 2 * V2_BM * 32 + 8, f1 + 8);
                }
            }
            const int pbuf = t % NPHI;
            uint8_t* sPhiT = sPhi + pbuf * 4 * V2_BLK;
            tc::mbar_wait(phi_empty + pbuf, ((t / NPHI) & 1) ^ 1);   // GEMM #2 and the store of the tile last held here are done
            if (tid == 0) TL(t, 3);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                *reinterpret_cast<float4*>(sPhiT + pb * V2_BLK + tc::sw128_chunk(r, pc + c4)) =
                    make_float4(p[4 * c4], p[4 * c4 + 1], p[4 * c4 + 2], p[4 * c4 + 3]);
                if (rbf)
                    *reinterpret_cast<float4*>(sPhiT + (2 + pb) * V2_BLK + tc::sw128_chunk(r, pc + c4)) =
                        make_float4(f1[4 * c4], f1[4 * c4 + 1], f1[4 * c4 + 2], f1[4 * c4 + 3]);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(phi_full + pbuf);
            if (tid == 0) TL(t, 4);
        }
        TLC(4);
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
        // ---- the LAST column split of a row block to get here adds the CS slabs of the block (slab order, the order of
        //      k_sum_slabs: bit-identical) into the dense input of the next layer -- a ticket per (chain, row block) instead
        //      of a ~7-10 us launch between two layers.  The ticket resets itself for the next launch. ----
        if (a.do_gemm2 && a.Fsum != nullptr && warp < 4) {
            uint32_t* last_s = reinterpret_cast<uint32_t*>(bars + 40);
            asm volatile("bar.sync 2, 128;" ::: "memory");         // the slab rows of all four warps are written ...
            if (tid == 0) {
                __threadfence();                                   // ... and ordered before the ticket (fences are cumulative: one suffices)
                unsigned int* ctr = a.sum_ctr + (int64_t)chain * gridDim.x + blockIdx.x;
                const unsigned int old = atomicAdd(ctr, 1u);
                const bool last = old + 1 == (unsigned int)a.CS;
                if (last) atomicExch(ctr, 0u);
                *last_s = last ? 1u : 0u;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (*last_s != 0u) {
                __threadfence();
                const int rows = min(V2_BM, a.B - row0);
                const int64_t ss = (int64_t)a.B * a.g;
                const float* src = a.Fpart + chain * a.fpart_cs + (int64_t)row0 * a.g;
                float* dst = a.Fsum + chain * a.fsum_cs + (int64_t)row0 * a.g;
                const int n = rows * a.g;
                int e_lo = 0;
                if (((ss | a.fpart_cs | a.fsum_cs) & 3) == 0 && (((int64_t)row0 * a.g) & 3) == 0) {      // 128-bit lanes
                    const int nv = n >> 2;
                    for (int e = tid; e < nv; e += 128) {
                        float4 acc = __ldcg(reinterpret_cast<const float4*>(src) + e);
                        for (int sl = 1; sl < a.CS; ++sl) {
                            const float4 t = __ldcg(reinterpret_cast<const float4*>(src + sl * ss) + e);
                            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                        }
                        reinterpret_cast<float4*>(dst)[e] = acc;
                    }
                    e_lo = nv << 2;
                }
                for (int e = e_lo + tid; e < n; e += 128) {
                    float acc = __ldcg(src + e);
                    for (int sl = 1; sl < a.CS; ++sl) acc += __ldcg(src + sl * ss + e);
                    dst[e] = acc;
                }
            }
        }
    } else if (warp == V2_EPI_WARPS) {
        // ===================================== MMA ISSUER =====================================
        // The whole warp walks the loop converged (waits included); one elected lane issues.  Descriptors are
        // built once: advancing along K or to another block only adds to the 14-bit start-address field.
        const uint64_t dB1 = tc::make_desc_sw128(tc::smem_u32(sB1));
        if (WIDE) {
            // K loop in SS mode: k-block `it` (counted over all tiles of the CTA) lives in ring slot it % NS1
            int it = 0;
            for (int t = 0; t < n_my; ++t) {
                const int buf = t & 1;
                tc::mbar_wait(d1_empty + buf, ((t >> 1) & 1) ^ 1);
                if (lane == 0) TL(t, 5);
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int slot = it % NS1;
                    tc::mbar_wait(b1_full + slot, (it / NS1) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint64_t dah = dB1 + ((slot * b1_stage) >> 4), dal = dah + (V2_BLK >> 4);
                        const uint64_t doh = dah + ((2 * V2_BLK) >> 4), dol = doh + (V2_BBLK >> 4);
                        const uint32_t dcol = tm_d1 + 64 * buf;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            tc::umma_tf32(dcol, dal + 2 * k4, doh + 2 * k4, IDESC1, k4 != 0 ? 1u : (kb != 0 ? 1u : 0u));
                            tc::umma_tf32(dcol, dah + 2 * k4, dol + 2 * k4, IDESC1, 1u);
                            tc::umma_tf32(dcol, dah + 2 * k4, doh + 2 * k4, IDESC1, 1u);
                        }
                        tc::umma_commit(b1_empty + slot);              // k-block consumed
                        if (kb == n_kb - 1) tc::umma_commit(d1_full + buf);   // P ready
                    }
                    __syncwarp();
                }
                if (lane == 0) TL(t, 6);
            }
        } else if (KRING) {
            // A in tensor memory, z k-blocks from the ring: 12 UMMAs per k-block, the slot is released as soon as they retire.
            // The k-block loop is unrolled with a compile-time kb (A columns and descriptor offsets are immediates, the ring
            // slot is one add per k-block) and the slot / phase are carried, not recomputed: with a runtime kb the issue of a
            // tile's 48 UMMAs took 3.7 k cycles (uniform-datapath arithmetic in front of every instruction) and bounded the
            // kernel once the sub-tile ring had taken the store off the critical path.
            int slot = 0;
            uint32_t ph = 0;
            for (int t = 0; t < n_my; ++t) {
                const int buf = t & 1;
                tc::mbar_wait(d1_empty + buf, ((t >> 1) & 1) ^ 1);
                if (lane == 0) TL(t, 5);
                const uint32_t dcol = tm_d1 + 64 * buf;
#pragma unroll
                for (int kb = 0; kb < V2_KB; ++kb) {
                    if (kb < n_kb) {
                        tc::mbar_wait(b1_full + slot, ph);
                        tc::tc_fence_after();
                        if (tc::elect_one()) {
                            const uint64_t dzh = dB1 + (uint32_t)(slot * (int)((2 * V2_BBLK) >> 4)), dzl = dzh + (V2_BBLK >> 4);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint32_t acol = 32 * kb + 8 * k4;
                                tc::umma_tf32_ts(dcol, tm_alo + acol, dzh + 2 * k4, IDESC1, (kb | k4) != 0 ? 1u : 0u);
                                tc::umma_tf32_ts(dcol, tm_ahi + acol, dzl + 2 * k4, IDESC1, 1u);
                                tc::umma_tf32_ts(dcol, tm_ahi + acol, dzh + 2 * k4, IDESC1, 1u);
                            }
                            tc::umma_commit(b1_empty + slot);              // k-block consumed
                            if (kb == n_kb - 1) tc::umma_commit(d1_full + buf);   // P ready
                        }
                        __syncwarp();
                        if (++slot == NS1) { slot = 0; ph ^= 1u; }
                    }
                }
                if (lane == 0) TL(t, 6);
            }
        } else
        for (int t = 0; t < n_my; ++t) {
            const int buf = t & 1;
            const int s1 = t % NS1;
            tc::mbar_wait(d1_empty + buf, ((t >> 1) & 1) ^ 1);
            tc::mbar_wait(b1_full + s1, (t / NS1) & 1);
            if (lane == 0) TL(t, 5);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint64_t db = dB1 + ((s1 * b1_stage) >> 4);
                const uint32_t dcol = tm_d1 + 64 * buf;
                switch (n_kb) {
                    case 1: issue_gemm1<1>(dcol, tm_ahi, tm_alo, db, IDESC1); break;
                    case 2: issue_gemm1<2>(dcol, tm_ahi, tm_alo, db, IDESC1); break;
                    case 3: issue_gemm1<3>(dcol, tm_ahi, tm_alo, db, IDESC1); break;
                    default: issue_gemm1<4>(dcol, tm_ahi, tm_alo, db, IDESC1); break;
                }
                TL(t, 11);
                tc::umma_commit(b1_empty + s1);        // z tile consumed
                tc::umma_commit(d1_full + buf);        // P ready
            }
            __syncwarp();
            if (lane == 0) TL(t, 6);
        }
    } else if (warp == V2_EPI_WARPS + 3) {
        // ===================================== MMA-2 ISSUER =====================================
        const uint64_t dPhi = tc::make_desc_sw128(tc::smem_u32(sPhi));
        const uint64_t dW = tc::make_desc_sw128(tc::smem_u32(sW));
        if (SUB3) {
            // sub-tile u = 2 t + h: A = its [cos | sin] blocks, B = W^T blocks h (cos rows) and 2 + h (sin rows) of tile t's stage
            for (int u = 0; u < 2 * n_my; ++u) {
                const int t = u >> 1, h = u & 1, q = u % 3, ws = t % NSW;
                tc::mbar_wait(phi_full + q, (u / 3) & 1);
                if (a.do_gemm2 && h == 0) tc::mbar_wait(w_full + ws, (t / NSW) & 1);
                if (lane == 0 && h == 0) TL(t, 7);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    if (a.do_gemm2) {
                        // straight-line issue (compile-time block / k offsets, constant accumulate flags; see issue_gemm1)
                        const uint64_t dw_h = dW + (uint32_t)(((ws * 4 + h) * NG * 128) >> 4);
                        const uint64_t dphi = dPhi + (uint32_t)((q * 2 * V2_BLK) >> 4);
                        if (rbf) issue_gemm2_sub<2, NG>(tm_d2, dphi, dw_h, IDESC2, u != 0 ? 1u : 0u);
                        else issue_gemm2_sub<1, NG>(tm_d2, dphi, dw_h, IDESC2, u != 0 ? 1u : 0u);
                        if (h == 1) tc::umma_commit(w_empty + ws);
                    }
                    tc::umma_commit(phi_empty + q);        // sub-tile consumed by the tensor core (1 of 2 arrivals)
                    if (u == 2 * n_my - 1) tc::umma_commit(d2_full);
                }
                __syncwarp();
                if (lane == 0 && h == 1) TL(t, 8);
            }
        } else
        for (int u = 0; u < n_my; ++u) {
            const int ws = u % NSW;
            const int pbuf = u % NPHI;
            tc::mbar_wait(phi_full + pbuf, (u / NPHI) & 1);
            if (a.do_gemm2) tc::mbar_wait(w_full + ws, (u / NSW) & 1);
            if (lane == 0) TL(u, 7);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                if (a.do_gemm2) {
                    const uint64_t dw = dW + (uint32_t)((ws * 4 * NG * 128) >> 4);
                    const uint64_t dphi = dPhi + (uint32_t)((pbuf * 4 * V2_BLK) >> 4);
                    if (rbf) issue_gemm2<4, NG>(tm_d2, dphi, dw, IDESC2, u != 0 ? 1u : 0u);
                    else issue_gemm2<2, NG>(tm_d2, dphi, dw, IDESC2, u != 0 ? 1u : 0u);
                    tc::umma_commit(w_empty + ws);
                }
                tc::umma_commit(phi_empty + pbuf);     // Phi tile consumed by the tensor core (1 of 2 arrivals)
                if (u == n_my - 1) tc::umma_commit(d2_full);
            }
            __syncwarp();
            if (lane == 0) TL(u, 8);
        }
    } else if (warp == V2_EPI_WARPS + 1) {
        // ===================================== TMA PRODUCER =====================================
        if (tc::elect_one()) {
            const int zc = a.zt_cs != 0 ? chain : 0;
            int it = 0;
            for (int t = 0; t < n_my; ++t) {
                const int c0 = (ct0 + t) * V2_BN;
                if (WIDE) {
                    // ---- k-blocks of this tile: [A hi | A lo] rows row0.. of the input, [Omega hi | Omega lo] feature rows c0.. ----
                    for (int kb = 0; kb < n_kb; ++kb, ++it) {
                        const int slot = it % NS1;
                        tc::mbar_wait(b1_empty + slot, ((it / NS1) & 1) ^ 1);
                        if (kb == 0) TL(t, 9);
                        tc::mbar_expect_tx(b1_full + slot, b1_stage);
                        const uint32_t sl = tc::smem_u32(sB1) + slot * b1_stage;
                        tc::tma_load_3d(&map_at, sl, b1_full + slot, 32 * kb, row0, 2 * chain);
                        tc::tma_load_3d(&map_at, sl + V2_BLK, b1_full + slot, 32 * kb, row0, 2 * chain + 1);
                        tc::tma_load_3d(&map_zt, sl + 2 * V2_BLK, b1_full + slot, 32 * kb, c0, 2 * chain);
                        tc::tma_load_3d(&map_zt, sl + 2 * V2_BLK + V2_BBLK, b1_full + slot, 32 * kb, c0, 2 * chain + 1);
                    }
                } else if (KRING) {
                    for (int kb = 0; kb < n_kb; ++kb, ++it) {
                        const int slot = it % NS1;
                        tc::mbar_wait(b1_empty + slot, ((it / NS1) & 1) ^ 1);
                        if (kb == 0) TL(t, 9);
                        tc::mbar_expect_tx(b1_full + slot, b1_stage);
                        const uint32_t sl = tc::smem_u32(sB1) + slot * b1_stage;
                        tc::tma_load_3d(&map_zt, sl, b1_full + slot, 32 * kb, c0, 2 * zc);
                        tc::tma_load_3d(&map_zt, sl + V2_BBLK, b1_full + slot, 32 * kb, c0, 2 * zc + 1);
                    }
                } else {
                // ---- z tile (B of GEMM #1): rows = feature columns, K-major, tf32 hi blocks then lo blocks ----
                const int s1 = t % NS1;
                tc::mbar_wait(b1_empty + s1, ((t / NS1) & 1) ^ 1);
                TL(t, 9);
                tc::mbar_expect_tx(b1_full + s1, b1_stage);
                const uint32_t b1 = tc::smem_u32(sB1) + s1 * b1_stage;
                for (int kb = 0; kb < n_kb; ++kb) {
                    tc::tma_load_3d(&map_zt, b1 + kb * V2_BBLK, b1_full + s1, 32 * kb, c0, 2 * zc);
                    tc::tma_load_3d(&map_zt, b1 + (n_kb + kb) * V2_BBLK, b1_full + s1, 32 * kb, c0, 2 * zc + 1);
                }
                }
                // ---- W^T tile (B of GEMM #2) into ring stage t & 1 ----
                if (a.do_gemm2) {
                    const int ws = t % NSW;
                    tc::mbar_wait(w_empty + ws, ((t / NSW) & 1) ^ 1);
                    TL(t, 10);
                    tc::mbar_expect_tx(w_full + ws, (uint32_t)nb2 * NG * 128);
                    for (int b = 0; b < nb2; ++b)
                        tc::tma_load_3d(&map_wt, tc::smem_u32(sW + (ws * 4 + b) * (NG * 128)), w_full + ws,
                                        (b >= 2 ? a.M : 0) + c0 + 32 * (b & 1), 0, chain);
                }
            }
        }
    } else {
        // ===================================== STORE WARP =====================================
        if (tc::elect_one()) {
            const bool storing = a.Phi != nullptr && !direct_store && !reg_store;
            if (SUB3) {
                for (int u = 0; u < 2 * n_my; ++u) {
                    const int t = u >> 1, h = u & 1, q = u % 3;
                    const int c0 = (ct0 + t) * V2_BN;
                    tc::mbar_wait(phi_full + q, (u / 3) & 1);
                    if (storing) {
                        const uint8_t* sub = sPhi + q * 2 * V2_BLK;
                        if (a.phi_blocked) {
                            const int nb = rbf ? 4 : 2;
                            const int blk0 = ((int)blockIdx.x * n_ct + (ct0 + t)) * nb;
                            tc::tma_store_3d(&map_cos, tc::smem_u32(sub), 0, (blk0 + h) * V2_BM, chain);
                            if (rbf) tc::tma_store_3d(&map_cos, tc::smem_u32(sub + V2_BLK), 0, (blk0 + 2 + h) * V2_BM, chain);
                        } else if (c0 + 32 * h < a.M) {
                            tc::tma_store_3d(&map_cos, tc::smem_u32(sub), c0 + 32 * h, row0, chain);
                            if (rbf) tc::tma_store_3d(&map_sin, tc::smem_u32(sub + V2_BLK), c0 + 32 * h, row0, chain);
                        }
                        tc::tma_commit();
                        if (u > 0) {
                            tc::tma_wait_read1();                  // the store of sub-tile u-1 has read its buffer; u's is in flight
                            tc::mbar_arrive(phi_empty + ((u - 1) % 3));
                        }
                    } else {
                        tc::mbar_arrive(phi_empty + q);
                    }
                }
                tc::tma_wait0();
            } else
            for (int t = 0; t < n_my; ++t) {
                const int c0 = (ct0 + t) * V2_BN;
                const int pbuf = t % NPHI;
                tc::mbar_wait(phi_full + pbuf, (t / NPHI) & 1);
                if (storing) {
                    const uint8_t* sPhiT = sPhi + pbuf * 4 * V2_BLK;
                    if (a.phi_blocked) {
                        // tile-blocked layout: block (row block, column tile, blk) starts at row ((rb * n_ct + ct) * nb + blk) * 128 of a
                        // [rows, 32] matrix; dead columns of a ragged last tile are stored as the zeros the epilogue wrote
                        const int nb = rbf ? 4 : 2;
                        const int blk0 = ((int)blockIdx.x * n_ct + (ct0 + t)) * nb;
                        for (int b = 0; b < 2; ++b) {
                            tc::tma_store_3d(&map_cos, tc::smem_u32(sPhiT + b * V2_BLK), 0, (blk0 + b) * V2_BM, chain);
                            if (rbf) tc::tma_store_3d(&map_cos, tc::smem_u32(sPhiT + (2 + b) * V2_BLK), 0, (blk0 + 2 + b) * V2_BM, chain);
                        }
                    } else
                    for (int b = 0; b < 2; ++b) {
                        if (c0 + 32 * b >= a.M) break;
                        tc::tma_store_3d(&map_cos, tc::smem_u32(sPhiT + b * V2_BLK), c0 + 32 * b, row0, chain);
                        if (rbf) tc::tma_store_3d(&map_sin, tc::smem_u32(sPhiT + (2 + b) * V2_BLK), c0 + 32 * b, row0, chain);
                    }
                    tc::tma_commit();
                    if (NPHI == 1) {
                        tc::tma_wait_read0();
                        tc::mbar_arrive(phi_empty + pbuf);         // 2 of 2 arrivals: the Phi tile may be overwritten
                    } else if (t > 0) {
                        tc::tma_wait_read1();                      // the store of tile t-1 has read its tile; tile t's is in flight
                        tc::mbar_arrive(phi_empty + ((t - 1) % NPHI));
                    }
                } else {
                    tc::mbar_arrive(phi_empty + pbuf);
                }
            }
            if (!SUB3) tc::tma_wait0();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    TLC(5);
    if (warp == V2_EPI_WARPS) tc::tmem_dealloc(tmem_base, V2_TMEM_COLS);
}

// ---- operand prep: z [d, M] -> zt [2][M][128] (tf32 hi | lo, K-major, zero padded); W [F, g] -> wt [NG][F] (tf32) ----
__global__ void __launch_bounds__(256)
k_prep_tc2(const float* __restrict__ z, int64_t z_cs, int d, int M, float* __restrict__ zt, int n_zt_tiles,
           const float* __restrict__ W, int64_t w_cs, int F, int g, int NG, float* __restrict__ wt, int64_t wt_cs) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int chain = blockIdx.y;
    int b = blockIdx.x;
    if (b < n_zt_tiles) {
        if (z_cs == 0 && chain > 0) return;                           // shared spectral draws: one copy
        const int m0 = (b >> 2) * 32, k0 = (b & 3) * 32;
        const float* zz = z + chain * z_cs;
        for (int i = ty; i < 32; i += 8)
            tile[i][tx] = (k0 + i < d && m0 + tx < M) ? __ldg(zz + (int64_t)(k0 + i) * M + m0 + tx) : 0.f;
        __syncthreads();
        float* hi = zt + (int64_t)chain * 2 * M * 128;
        float* lo = hi + (int64_t)M * 128;
        for (int i = ty; i < 32; i += 8)
            if (m0 + i < M) {
                const float v = tile[tx][i];
                const float h = tc::to_tf32(v);
                hi[(int64_t)(m0 + i) * 128 + k0 + tx] = h;
                lo[(int64_t)(m0 + i) * 128 + k0 + tx] = tc::to_tf32(v - h);
            }
        return;
    }
    b -= n_zt_tiles;
    const int nj = (NG + 31) / 32;
    const int f0 = (b / nj) * 32, j0 = (b % nj) * 32;
    const float* WW = W + chain * w_cs;
    for (int i = ty; i < 32; i += 8)
        tile[i][tx] = (f0 + i < F && j0 + tx < g) ? __ldg(WW + (int64_t)(f0 + i) * g + j0 + tx) : 0.f;
    __syncthreads();
    float* o = wt + chain * wt_cs;
    for (int i = ty; i < 32; i += 8)
        if (j0 + i < NG && f0 + tx < F) o[(int64_t)(j0 + i) * F + f0 + tx] = tc::to_tf32(tile[tx][i]);
}

// ---- WIDE operand prep ------------------------------------------------------------------------------------------
// at [2][B][Kp]: the layer input (previous layer's output slabs summed in slab order | model input), tf32 hi / lo, zero padded
__global__ void __launch_bounds__(256)
k_prep_wide_a(const FwdArgs a, int Kp, float* __restrict__ at) {
    dgprf_pdl_sync();
    // grid (row blocks of 8 rows, chains); a thread converts 4 consecutive K columns of one row per step (float4 stores)
    const int chain = blockIdx.y;
    const int64_t n = (int64_t)a.B * Kp;
    const float* X = a.X + chain * a.x_cs;
    const float* fp = a.Fprev.ptr + chain * a.Fprev.cs;
    float* hi = at + (int64_t)chain * 2 * n;
    float* lo = hi + n;
    const int k4n = Kp >> 2;                               // float4 per row
    const int r_in = threadIdx.x / 32, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + r_in; row < a.B; row += (int64_t)gridDim.x * 8) {
        for (int k4 = lane; k4 < k4n; k4 += 32) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = 4 * k4 + i;
                float x = 0.f;
                if (q < a.d_prev) {
                    x = __ldg(fp + row * a.Fprev.ld + q);
                    for (int sl = 1; sl < a.Fprev.n_slabs; ++sl) x += __ldg(fp + sl * a.Fprev.ss + row * a.Fprev.ld + q);
                } else if (q < a.d) {
                    x = __ldg(X + row * a.ldx + (q - a.d_prev));
                }
                v[i] = x;
            }
            float4 h, l;
            h.x = tc::to_tf32(v[0]); h.y = tc::to_tf32(v[1]); h.z = tc::to_tf32(v[2]); h.w = tc::to_tf32(v[3]);
            l.x = tc::to_tf32(v[0] - h.x); l.y = tc::to_tf32(v[1] - h.y); l.z = tc::to_tf32(v[2] - h.z); l.w = tc::to_tf32(v[3] - h.w);
            *reinterpret_cast<float4*>(hi + row * Kp + 4 * k4) = h;
            *reinterpret_cast<float4*>(lo + row * Kp + 4 * k4) = l;
        }
    }
}
// ot [2][M][Kp]: Omega^T = (exp(log_inv_ls) * z + mean)^T, tf32 hi / lo, K-major, zero padded (layers/rf_layers.py: Omega)
__global__ void __launch_bounds__(256)
k_prep_wide_o(const FwdArgs a, int Kp, float* __restrict__ ot) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int chain = blockIdx.z;
    const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const float* zz = a.z + chain * a.z_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    for (int i = ty; i < 32; i += 8) {
        const int q = k0 + i;
        float v = 0.f;
        if (q < a.d && m0 + tx < a.M) {
            v = expf(__ldg(ls + q)) * __ldg(zz + (int64_t)q * a.M + m0 + tx);
            if (mean != nullptr) v += __ldg(mean + q);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    float* hi = ot + (int64_t)chain * 2 * a.M * Kp;
    float* lo = hi + (int64_t)a.M * Kp;
    for (int i = ty; i < 32; i += 8)
        if (m0 + i < a.M) {
            const float v = tile[tx][i];
            const float h = tc::to_tf32(v);
            hi[(int64_t)(m0 + i) * Kp + k0 + tx] = h;
            lo[(int64_t)(m0 + i) * Kp + k0 + tx] = tc::to_tf32(v - h);
        }
}

int64_t dgprf_phi_blocked_floats(int B, int M, int kind) {
    return (int64_t)ceil_div(B, V2_BM) * ceil_div(M, V2_BN) * (kind == DGPRF_KIND_RBF ? 4 : 2) * (V2_BM * 32);
}

// tensor maps of the saved-feature store: row-major [B, F] (cos half | sin half) or the tile-blocked layout (one map)
static int make_phi_maps(const FwdArgs& a, int n_chains, CUtensorMap* mc, CUtensorMap* ms) {
    if (a.phi_blocked) {
        const int64_t n = dgprf_phi_blocked_floats(a.B, a.M, a.kind);
        DGPRF_REQUIRE(a.phi_cs >= n, "blocked saved-feature buffer too small: %lld < %lld floats", (long long)a.phi_cs, (long long)n);
        const int rc = dgprf_make_tmap_3d(mc, a.Phi, 32, (uint64_t)(n / 32), n_chains, 32, a.phi_cs, V2_BM);
        *ms = *mc;
        return rc;
    }
    int rc = dgprf_make_tmap_3d(mc, a.Phi, a.M, a.B, n_chains, a.F, a.phi_cs, V2_BM);
    if (rc) return rc;
    if (a.kind == DGPRF_KIND_RBF) rc = dgprf_make_tmap_3d(ms, a.Phi + a.M, a.M, a.B, n_chains, a.F, a.phi_cs, V2_BM);
    return rc;
}

static size_t tc2_smem_bytes(int NG, int n_kb, int ns1) {
    return 1024 + V2_HDR + 4 * (size_t)V2_BLK + 2 * 4 * (size_t)NG * 128 + (size_t)ns1 * 2 * n_kb * V2_BBLK;
}
static size_t tc2_wide_smem_bytes(int NG, int nsw, int ns) {
    return 1024 + V2_HDR + 4 * (size_t)V2_BLK + (size_t)nsw * 4 * NG * 128 + (size_t)ns * V2_RING;
}

template <int NG>
static int launch_fwd_tc2_wide(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const int Kp = (a.d + 31) & ~31;
    const int nsw = NG >= 32 ? 1 : 2;
    int ns = 3;
    while (ns > 1 && tc2_wide_smem_bytes(NG, nsw, ns) > 232448) --ns;
    const size_t smem = tc2_wide_smem_bytes(NG, nsw, ns);
    { const int rc_s = dgprf_ensure_smem((const void*)k1_fwd_tc2<NG, true>, (size_t)232448); if (rc_s) return rc_s; }
    DGPRF_REQUIRE(a.at != nullptr && a.ot != nullptr && a.wt != nullptr, "wide pipelined forward needs the prepped operand buffers");
    {
        ProfScope _ps("k_prep_tc2_wide", st);
        int blocks = ceil_div(a.B, 8);
        if (blocks > 148 * 16) blocks = 148 * 16;
        DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_prep_wide_a, dim3(blocks, n_chains), dim3(256), 0, st, a, Kp, a.at));
        if (!a.prepped) {
            k_prep_wide_o<<<dim3(ceil_div(a.M, 32), Kp / 32, n_chains), 256, 0, st>>>(a, Kp, a.ot);
            const int n_wt_tiles = ceil_div(a.F, 32) * ceil_div(NG, 32);
            k_prep_tc2<<<dim3(n_wt_tiles, n_chains), 256, 0, st>>>(a.z, a.z_cs, a.d, a.M, nullptr, 0, a.W, a.w_cs, a.F, a.g, NG, a.wt, (int64_t)NG * a.F);
        }
        DGPRF_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap mc, ms, mo, mw, ma;
    memset(&mc, 0, sizeof(mc));
    memset(&ms, 0, sizeof(ms));
    if (a.Phi != nullptr) {
        const int rc = make_phi_maps(a, n_chains, &mc, &ms);
        if (rc) return rc;
    }
    int rc = dgprf_make_tmap_3d(&mo, a.ot, Kp, a.M, 2 * n_chains, Kp, (uint64_t)a.M * Kp, V2_BN);
    if (rc) return rc;
    rc = dgprf_make_tmap_3d(&ma, a.at, Kp, a.B, 2 * n_chains, Kp, (uint64_t)a.B * Kp, V2_BM);
    if (rc) return rc;
    rc = dgprf_make_tmap_3d(&mw, a.wt, a.F, NG, n_chains, a.F, (uint64_t)NG * a.F, NG);
    if (rc) return rc;
    static long long* tl = nullptr;
    dim3 grid(ceil_div(a.B, V2_BM), a.CS, n_chains);
    { ProfScope _ps("k1_fwd_tc2_wide", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k1_fwd_tc2<NG, true>, grid, dim3(V2_THREADS), smem, st, a, ns, nsw | (getenv("DGPRF_TC2_DIRECT_STORE") ? 0x100 : 0) | (getenv("DGPRF_TC2_REG_STORE") ? 0x800 : 0), tl, mc, ms, mo, mw, ma)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

template <int NG>
static int launch_fwd_tc2(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const int n_kb = (a.d + 31) / 32;
    // DGPRF_TC2_PHI2=1 (A/B experiment): two Phi tiles + a single-stage z ring instead of one Phi tile + two z stages, so that
    // the store of a tile overlaps the next tile's epilogue.  Measured at configs[4] layer scale: 0.71-0.76 ms against 0.65 ms --
    // with one z stage GEMM #1 of tile t+1 waits for its 64 KB z tile after GEMM #1 of tile t; the second Phi tile needs the
    // z ring cut into k-block slots first.
    const bool phi2 = a.Phi != nullptr && getenv("DGPRF_TC2_PHI2") && tc2_smem_bytes(NG, n_kb, 1) + 4 * (size_t)V2_BLK <= 232448;
    // z k-block ring (4 slots of [hi 8 KB | lo 8 KB]) + two Phi tiles when the features are stored and a tile needs at most
    // three k-blocks (input width <= 96: the ring then has a slot of slack and the store of tile t overlaps the epilogue of
    // tile t+1 -- 0.62 -> 0.57 ms at B = 65536, d = 90, M = 4096); with four k-blocks per tile it loses to the two whole-tile
    // stages (0.65 -> 0.69 ms: shared memory has no room for a fifth slot).  DGPRF_TC2_KRING=<slots | 0> overrides.
    const char* e_kr = getenv("DGPRF_TC2_KRING");
    // (with the register-store experiment no TMA store sits on the tile's critical path: one Phi tile and two whole-tile z
    //  stages, as in the eval-mode forward)
    const bool regst = a.phi_blocked && getenv("DGPRF_TC2_REG_STORE");
    // Default when the features are stored: THREE Phi sub-tile buffers (96 KB) + as many k-block slots as fit (six with
    // n_gp <= 32) -- the store of a sub-tile drains under the epilogue of the next two and GEMM #1 stays 1.5 tiles ahead.
    // Tried on top of it and measured slower or equal at configs[4] layer scale (0.54 ms): W^T tiles loaded by the GEMM #2
    // warp (blocking on the retirement of GEMM #2: 0.54; one tile later: 0.64; three W stages + five z slots: 0.54), by
    // polling from the producer (0.69: the z stream falls behind with five slots), and forcing the issue order
    // G1(t+1), G2(t), G1(t+2) through a barrier between the two issuing threads (0.55-0.60).  What bounds a tile now is the
    // in-order tensor pipe: GEMM #2 of tile t retires behind the GEMM #1 work queued in front of it, and its commit is what
    // frees a Phi sub-tile buffer and a W stage.
    // DGPRF_TC2_NO_SUB3=1 gives round 2's earlier geometry back (two whole Phi tiles + four slots for <= 3 k-blocks, one Phi
    // tile + two whole-tile z stages otherwise).
    const size_t smem_fix3 = 1024 + V2_HDR + 6 * (size_t)V2_BLK + 2 * 4 * (size_t)NG * 128;
    int sub3_slots = (a.Phi != nullptr && !regst && !phi2 && !e_kr && !getenv("DGPRF_TC2_NO_SUB3") && !getenv("DGPRF_TC2_DIRECT_STORE"))
                         ? (int)((232448 - (long long)smem_fix3) / (2 * V2_BBLK)) : 0;
    if (sub3_slots > 6) sub3_slots = 6;
    const bool sub3 = sub3_slots >= 3 && sub3_slots > n_kb;
    int kring = a.Phi == nullptr ? 0 : (e_kr ? atoi(e_kr) : ((n_kb <= 3 && !regst) ? 4 : 0));
    if (kring > 6) kring = 6;
    const int nsw_k = getenv("DGPRF_TC2_NSW") ? atoi(getenv("DGPRF_TC2_NSW")) : 2;      // W^T ring stages next to the k-block ring (1 | 2)
    const size_t smem_kring = 1024 + V2_HDR + 8 * (size_t)V2_BLK + (size_t)(nsw_k == 1 ? 1 : 2) * 4 * (size_t)NG * 128 + (size_t)kring * 2 * V2_BBLK;
    if (kring < 2 || smem_kring > 232448) kring = 0;
    if (sub3) kring = sub3_slots;
    const int ns1 = kring ? kring : (phi2 ? 1 : (tc2_smem_bytes(NG, n_kb, 2) <= 232448 ? 2 : 1));
    const size_t smem = sub3 ? smem_fix3 + (size_t)kring * 2 * V2_BBLK
                             : (kring ? smem_kring : tc2_smem_bytes(NG, n_kb, ns1) + (phi2 ? 4 * (size_t)V2_BLK : 0));
    { const int rc_s = dgprf_ensure_smem((const void*)k1_fwd_tc2<NG, false>, (size_t)232448); if (rc_s) return rc_s; }
    DGPRF_REQUIRE(a.zt != nullptr && a.wt != nullptr, "pipelined forward needs the prepped operand buffers");
    if (!a.prepped) {
        const int n_zt_tiles = ceil_div(a.M, 32) * 4;
        const int n_wt_tiles = ceil_div(a.F, 32) * ceil_div(NG, 32);
        ProfScope _ps("k_prep_tc2", st);
        k_prep_tc2<<<dim3(n_zt_tiles + n_wt_tiles, n_chains), 256, 0, st>>>(a.z, a.z_cs, a.d, a.M, a.zt, n_zt_tiles,
                                                                          a.W, a.w_cs, a.F, a.g, NG, a.wt, (int64_t)NG * a.F);
        DGPRF_CHECK_CUDA(cudaGetLastError());
    }
    CUtensorMap mc, ms, mz, mw;
    memset(&mc, 0, sizeof(mc));
    memset(&ms, 0, sizeof(ms));
    if (a.Phi != nullptr) {
        const int rc = make_phi_maps(a, n_chains, &mc, &ms);
        if (rc) return rc;
    }
    int rc = dgprf_make_tmap_3d(&mz, a.zt, 128, a.M, 2 * (a.z_cs != 0 ? n_chains : 1), 128, (uint64_t)a.M * 128, V2_BN);
    if (rc) return rc;
    rc = dgprf_make_tmap_3d(&mw, a.wt, a.F, NG, n_chains, a.F, (uint64_t)NG * a.F, NG);
    if (rc) return rc;
    FwdArgs b = a;
    b.zt_cs = a.z_cs != 0 ? 1 : 0;
    dim3 grid(ceil_div(a.B, V2_BM), a.CS, n_chains);
    static long long* tl = nullptr;                       // debug timeline (DGPRF_TC2_TIMELINE=1)
    static int tl_calls = 0;
    if (getenv("DGPRF_TC2_TIMELINE") && !tl) { cudaMalloc(&tl, (16 * 12 + 8) * sizeof(long long)); }
    if (tl) cudaMemsetAsync(tl, 0, (16 * 12 + 8) * sizeof(long long), st);
    { ProfScope _ps("k1_fwd_tc2", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k1_fwd_tc2<NG, false>, grid, dim3(V2_THREADS), smem, st, b, ns1, ((kring && !sub3 && nsw_k == 1) ? 1 : 2) | (getenv("DGPRF_TC2_DIRECT_STORE") ? 0x100 : 0) | ((phi2 || kring) ? 0x200 : 0) | (kring ? 0x400 : 0) | (sub3 ? 0x1000 : 0) | (getenv("DGPRF_TC2_REG_STORE") ? 0x800 : 0) | ((tl && getenv("DGPRF_TC2_TIMELINE_RB")) ? (atoi(getenv("DGPRF_TC2_TIMELINE_RB")) << 16) : 0), tl, mc, ms, mz, mw, mw)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    if (tl && ++tl_calls == (atoi(getenv("DGPRF_TC2_TIMELINE")) > 1 ? atoi(getenv("DGPRF_TC2_TIMELINE")) : 4)) {
        long long h[16 * 12 + 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, tl, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[9];                         // producer's first stamp
        const long long* c = h + 16 * 12;
        fprintf(stderr, "tc2 CTA phases (cycles): barriers + TMEM alloc %lld | input tile -> smem %lld | A operand -> TMEM %lld | %d tiles %lld | F slab + drain %lld | total %lld (grid %u x %u)\n",
                c[1] - c[0], c[2] - c[1], c[3] - c[2], (int)ceil_div(ceil_div(a.M, V2_BN), a.CS), c[4] - c[3], c[5] - c[4], c[5] - c[0], grid.x, grid.y);
        fprintf(stderr, "tc2 timeline (cycles since the producer started; train=%d)\n tile | epi: P-ready P-loaded sincos-done phi-free stored | mma: g1-go g1-issued g2-go g2-issued | prod: z-slot w-slot | g1-mmas-issued(before commits)\n", a.Phi != nullptr);
        for (int t = 0; t < 16; ++t) {
            fprintf(stderr, " %3d |", t);
            for (int e = 0; e < 12; ++e) fprintf(stderr, " %7lld%s", h[t * 12 + e] ? h[t * 12 + e] - t0 : -1LL, (e == 4 || e == 8 || e == 10) ? " |" : "");
            fprintf(stderr, "\n");
        }
    }
    return DGPRF_OK;
}

// The pipelined kernel takes 64-column tiles and an input width that fits the TMEM-resident A operand.  Returns the
// number of column splits (CTAs per row block; each walks M/64/splits column tiles) or 0 when the SIMT
// kernel should run instead: the per-CTA prologue and pipeline fill only pay off over >= 4 tiles per CTA.
// Tile-width hint for a layer: 32 when 64-wide column tiles would leave most of the 148 SMs without a CTA (small
// minibatches: the pipelined kernel then runs with as little as one tile per CTA), else 64.
int dgprf_tc_tile_cols(int B, int M, int n_chains) {
    const int64_t ctas64 = (int64_t)ceil_div(B, V2_BM) * (ceil_div(M, 64) < kMaxCS ? ceil_div(M, 64) : kMaxCS) * n_chains;
    return ctas64 < 120 ? 32 : 64;
}

int dgprf_fwd_tc2_col_splits(int tile_cols, int B, int d, int M, int g, int n_chains) {
    if ((M % 4) != 0 || g > 64 || getenv("DGPRF_NO_TC2") != nullptr) return 0;
    const int n_ct = ceil_div(M, V2_BN);
    const int64_t rb = (int64_t)ceil_div(B, V2_BM) * n_chains;
    // small grids (tile_cols == 32) still take this kernel, one tile per CTA if need be (measured faster than the fp32 SIMT
    // kernel from about 4 tiles in total); below that the SIMT kernel runs
    if (tile_cols != 64 && (rb * n_ct < 4 || getenv("DGPRF_NO_TC2_SMALL") != nullptr)) return 0;
    int cs = (int)((4 * 148 + rb - 1) / rb);                   // >= 4 waves of CTAs when the problem allows
    if (cs > kMaxCS) cs = kMaxCS;
    if (cs > n_ct) cs = n_ct;
    if (cs < 1) cs = 1;
    if (d > 128) {
        // WIDE variant: every tile is a K loop of d/32 k-blocks, so one tile per CTA already amortises the start-up;
        // the operand buffers hold the input / Omega split in two, which bounds the size this path is taken for
        if (getenv("DGPRF_NO_TC2_WIDE") != nullptr || (int64_t)B * ((d + 31) & ~31) * n_chains > ((int64_t)1 << 28)) return 0;
        return cs;
    }
    const int min_tiles = getenv("DGPRF_TC2_MIN_TILES") ? atoi(getenv("DGPRF_TC2_MIN_TILES")) : (tile_cols != 64 ? 1 : 4);
    while (cs > 1 && n_ct / cs < min_tiles) --cs;
    if (n_ct / cs < min_tiles) return 0;
    if (tile_cols == 64) {
        // Every CTA pays a fixed ~5 tile times before and after its tiles (measured at configs[4] layer scale, per-CTA stamps
        // of DGPRF_TC2_TIMELINE: input tile -> shared memory 8.1 k cycles, A operand -> tensor memory 2.9 k, barriers + TMEM
        // allocation 1.3 k, F slab + store drain 6.1 k; a tile is 4.35 k) and CTAs run one per SM in waves, so "as many CTAs
        // as possible" loses when the row blocks are few: 8 192 rows ran 64 x 8 CTAs of 8 tiles in four waves (127-137 us)
        // where 64 x 2 CTAs of 32 tiles take one.  Pick the split with the smallest  waves x (5 + tiles per CTA).
        if (const char* e = getenv("DGPRF_TC2_CS")) {
            const int f = atoi(e);
            if (f >= 1 && f <= kMaxCS && f <= n_ct) return f;
        }
        int best = cs;
        int64_t best_cost = -1;
        for (int c = 1; c <= cs; ++c) {
            const int64_t cost = ceil_div(rb * c, (int64_t)148) * (5 + ceil_div(n_ct, c));
            if (best_cost < 0 || cost < best_cost) { best = c; best_cost = cost; }
        }
        cs = best;
    }
    return cs;
}
bool dgprf_fwd_tc2_supported(const FwdArgs& a) {
    return a.wt != nullptr && (a.d <= 128 ? a.zt != nullptr : (a.at != nullptr && a.ot != nullptr)) &&
           (a.Phi == nullptr || (a.phi_cs % 4) == 0);
}
int64_t dgprf_fwd_tc2_at_floats(int B, int d) { return d > 128 ? 2 * (int64_t)B * ((d + 31) & ~31) : 0; }
int64_t dgprf_fwd_tc2_ot_floats(int M, int d) { return d > 128 ? 2 * (int64_t)M * ((d + 31) & ~31) : 0; }
// floats of the prepped operand buffers: zt per spectral-draw copy, wt per chain
int64_t dgprf_fwd_tc2_zt_floats(int M) { return 2 * (int64_t)M * 128; }
int64_t dgprf_fwd_tc2_wt_floats(int F, int g) { return (int64_t)(g <= 16 ? 16 : g <= 32 ? 32 : 64) * F; }

int dgprf_launch_fwd_tc2(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const int g = a.g;
    if (a.d > 128) {
        if (g <= 16) return launch_fwd_tc2_wide<16>(a, n_chains, st);
        if (g <= 32) return launch_fwd_tc2_wide<32>(a, n_chains, st);
        return launch_fwd_tc2_wide<64>(a, n_chains, st);
    }
    if (g <= 16) return launch_fwd_tc2<16>(a, n_chains, st);
    if (g <= 32) return launch_fwd_tc2<32>(a, n_chains, st);
    return launch_fwd_tc2<64>(a, n_chains, st);
}
