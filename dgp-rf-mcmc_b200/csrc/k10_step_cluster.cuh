// K10 device code and per-instantiation launchers, included by k10_step_cluster.cu (planner + dispatch) and by the
// k10_inst_cl*.cu translation units that instantiate the kernel for one cluster size each (compiled in parallel).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "update_core.cuh"

#ifndef K10_VAR
#define K10_VAR 0
#endif

#define K10_GAUSS(a) ((a).likelihood == DGPRF_LIK_GAUSSIAN)
#define K10_MEAN(y) ((y).has_mean)
#define K10_XS(a) ((a).x_in_smem)
#define K10_LPV8(a) ((a).upd_lpv == 8)

constexpr int kT = 512;          // threads per CTA
constexpr int kW = kT / 32;      // warps
constexpr int kNJ = 4;           // most 8-wide tiles of a GP output / previous-layer width (g, d_prev + 1 <= 32)
constexpr int kFS = 33;          // row stride of the fp32 [RT][<=32] matrix f_s

struct ClLayer {
    int32_t kind, d_prev, d_x, M, g, has_mean;
    int32_t cols;                // P columns per CTA (multiple of 8)
    int32_t ldp;                 // row stride of the saved Phi tile [RT][ldp]
    int32_t phi_off;             // float offset of the tile in the phi region
    const float* z; int64_t z_cs;
    const float* log_inv_ls; const float* log_amp; const float* mean;   // + chain*h_cs
    const float* W;                                                     // + chain*w_cs
    int64_t off_W;                                                      // into a gradient slab
};

struct ClArgs {
    int32_t n_layers, likelihood, B, d_in, d_out, CL, lda, dmax, ncs;
    int32_t x_in_smem;                       // the [RT, d_in] input rows are staged in shared memory (small d_in); else read from L2
    int64_t h_cs, w_cs;
    const float* X; int64_t x_cs;
    const float* Y; int64_t y_cs;
    const float* lik_log_var;
    float* gwpart; int64_t gw_cs, gw_ss;     // [C][n_tiles][w_len]
    float* ll_part; int64_t ll_cs;           // [C][n_tiles]
    float inv_B;
    int32_t fuse_update, upd_lpv;
    unsigned int* bar;                       // [2] {arrival count, generation} (zero-initialised workspace words)
    float* u_out;                            // [C] sum_i ll_i (nullable; fused path only)
    long long* timing;                       // debug phase stamps of CTA 0 (nullable)
    UpdArgs upd;
    ClLayer layer[DGPRF_MAX_LAYERS];
};

namespace {

// x = hi + lo, both rounded to nearest tf32 (ties away): adding half an ulp (0x1000) to the bit pattern and dropping the
// low 13 bits -- which the tensor core does by itself, so lo is only biased, not masked.  Four instructions;
// cvt.rna.tf32.f32 is a four-instruction emulation with an Inf/NaN guard on sm_100a (nine per hi / lo split).  Rounding
// (not truncation) matters: truncated lo terms all err towards zero and the bias adds up coherently over a K = 1024 sum.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
// D += A(16x8, row) * B(8x8, col), tf32 in, fp32 accumulate.  lane = 4*g + t:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);   b0 (k=t, n=g)  b1 (k=t+4, n=g);
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32: (ah + al)(bh + bl) without the al*bl term, small terms first
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma_tf32(c, al, bh0, bh1);
    mma_tf32(c, ah, bl0, bl1);
    mma_tf32(c, ah, bh0, bh1);
}
// accumulator fragment (cols 2t, 2t+1) -> A fragment (K slots t, t+4), split hi / lo
__device__ __forceinline__ void acc_to_a(const float (&v)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    split_tf32(v[0], ah[0], al[0]);
    split_tf32(v[2], ah[1], al[1]);
    split_tf32(v[1], ah[2], al[2]);
    split_tf32(v[3], ah[3], al[3]);
}

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Grid-wide barrier for the cooperative launch (all CTAs co-resident), ONE atomic per CTA: `ticket` only ever grows, a CTA
// that draws ticket t waits until the counter reaches the end of t's round, (t / n + 1) * n -- the last arriver's own atomic
// is the release, there is no second "generation" word to bump (one L2 round trip less on the critical path than K9's
// barrier) and nothing to reset (comparisons are on the signed difference, so the 32-bit counter may wrap).
// Bounded spin: a lost arrival traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier_cl(unsigned int* ticket, unsigned int n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        const unsigned int target = (t / n_ctas + 1u) * n_ctas;
        if (t + 1u != target) {
            const long long t0 = clock64();
            while ((int)(*reinterpret_cast<volatile unsigned int*>(ticket) - target) < 0)
                if (clock64() - t0 > 4000000000LL) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- cluster exchange: pushed partials, counted by a transaction barrier ------------------------------------------
// Every CTA owns receive slots rx[2][CL][RT*ncs] (ping-pong x source rank) and two mbarriers.  A sender writes its partial
// straight into the peers' slots with st.async (a remote shared-memory store that completes `bytes` on the RECEIVER's
// mbarrier), so the receiver needs no cluster-wide barrier, no gpu-scope fence and no L1 invalidate: it arms its barrier
// with the byte count it expects and waits on its own shared memory.  (barrier.cluster.arrive.release compiles to
// MEMBAR.ALL.GPU + ERRBAR and the wait to CCTL.IVALL: ~3000 cycles per exchange in the pull version.)
// Ping-pong safety: a peer sends exchange n only after it has consumed exchange n-1, which it received from me after I
// had finished reading exchange n-2 out of the same slots.
struct Xchg {
    float* red;              // [kW][RT][ncs] per-warp partials
    float* rx;               // [2][CL][RT*ncs]
    uint64_t* bar;           // [2]
    int ncs, CL, rank;
    uint32_t n;              // exchanges so far
};
__device__ __forceinline__ uint32_t map_peer(const void* local, uint32_t rank) {
    const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    return ra;
}
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();           // a lost partial traps instead of hanging the GPU
    }
}

// Cross-warp + cross-CTA reduction of a per-warp accumulator set acc[MT][kNJ][4] (rows of the cluster's row tile x nt*8
// columns).  Warp partials -> `red`, fixed-order sum over the warps = this CTA's partial, pushed to every CTA of the
// cluster; then every CTA sums the CL partials in rank order (deterministic, identical on all CTAs) and hands element
// (r, c) to `consume` on the thread that owns it.  The caller synchronises the CTA after its own follow-up work.
template <int MT, int NJM, int CL, typename Consume>
__device__ __forceinline__ void reduce_gather(const float (&acc)[MT][NJM][4], int nt, Xchg& x, Consume&& consume) {
    constexpr int RT = 16 * MT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ncs = x.ncs;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NJM; ++j)
            if (j < nt) {
                float* p = x.red + ((warp * RT + mt * 16 + g) * ncs + j * 8 + 2 * t);
                *reinterpret_cast<float2*>(p) = make_float2(acc[mt][j][0], acc[mt][j][1]);
                *reinterpret_cast<float2*>(p + 8 * ncs) = make_float2(acc[mt][j][2], acc[mt][j][3]);
            }
    __syncthreads();
    const int nc = nt * 8;
    const int ph = x.n & 1;
    float* rxp = x.rx + ph * (CL * RT * ncs);
    uint64_t* bar = x.bar + ph;
    if (CL > 1 && tid == 0) mbar_expect_tx(bar, (uint32_t)((CL - 1) * RT * nc * sizeof(float)));
    uint32_t peer_rx[CL], peer_bar[CL];
    if (CL > 1) {
#pragma unroll
        for (int k = 0; k < CL; ++k) {
            peer_rx[k] = map_peer(rxp + x.rank * RT * ncs, (uint32_t)k);     // my slot in CTA k's receive buffer
            peer_bar[k] = map_peer(bar, (uint32_t)k);
        }
    }
    for (int e = tid; e < RT * nc; e += kT) {
        const int r = e / nc, c = e - r * nc;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kW; ++w) s += x.red[(w * RT + r) * ncs + c];
        if (CL == 1) {
            consume(r, c, s);
        } else {
            const int o = r * ncs + c;
            rxp[x.rank * RT * ncs + o] = s;
#pragma unroll
            for (int k = 0; k < CL; ++k)
                if (k != x.rank) st_async_f32(peer_rx[k] + 4u * o, s, peer_bar[k]);
        }
    }
    if (CL > 1) {
        mbar_wait_parity(bar, (x.n >> 1) & 1);
        for (int e = tid; e < RT * nc; e += kT) {
            const int r = e / nc, c = e - r * nc;
            const int o = r * ncs + c;
            float pv[CL];
#pragma unroll
            for (int k = 0; k < CL; ++k) pv[k] = rxp[k * RT * ncs + o];
            float s = pv[0];
#pragma unroll
            for (int k = 1; k < CL; ++k) s += pv[k];         // rank order
            consume(r, c, s);
        }
    }
    ++x.n;
}

// NJM: most 8-wide tiles of any GP output / previous-layer width of the model (1 | 2 | 4): every loop over those tiles is
// unrolled to NJM with warp-uniform guards, and the unrolled-but-skipped iterations are not free (code size, guards):
// configs[1] (n_gp = 9 -> two tiles) runs 8 % faster with NJM = 2 than with 4.
// KIND: 0 every layer RBF, 2 decided per layer at run time (arc-cosine or mixed models): the feature-map branches of
// the epilogue, the second (sine) block of GEMM #2 / dPhi / gW and the dP formula become compile-time.
// CLT: the cluster size (1 | 2 | 4 | 8): the peer loops of the exchange unroll exactly, single-CTA clusters lose the exchange.
template <int MT, int NJM, int KIND, int CLT>
__global__ void __launch_bounds__(kT, 1) k10_step_cluster(const __grid_constant__ ClArgs a, const __grid_constant__ SegTable tab) {
    constexpr int RT = 16 * MT;
    constexpr int LDT = RT + 4;                          // row stride of the transposed dF operand
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int CL = CLT;
    const int L = a.n_layers;
    const int rank = CL > 1 ? (int)cluster_rank() : 0;
    const int tile = blockIdx.x / CL, chain = blockIdx.y;
    const int row0 = tile * RT;

    float* x_s   = sm;                                   // [RT][d_in]
    float* f_s   = x_s + (K10_XS(a) ? ((RT * a.d_in + 3) & ~3) : 0);       // [RT][kFS]   F_l / dF_l / raw T_l (fp32)
    float* bias  = f_s + ((RT * kFS + 3) & ~3);          // [RT]        in . mean (forward); per-row log-likelihood
    float* s_all = bias + RT;                            // [L][dmax]   exp(log_inv_ls)
    float* m_all = s_all + L * a.dmax;                   // [L][dmax]   mean
    float* a_hi  = m_all + L * a.dmax;                   // [RT][lda]   A operand of GEMM #1 (in*s) / of dPhi (dF), tf32 hi
    float* a_lo  = a_hi + RT * a.lda;                    //             ... lo
    float* t_hi  = a_lo + RT * a.lda;                    // [32][LDT]   dF^T (A operand of gW^T = dF^T Phi), hi
    float* t_lo  = t_hi + 32 * LDT;                      //             ... lo
    float* red   = t_lo + 32 * LDT;                      // [kW][RT][ncs] per-warp partials
    float* rx    = red + kW * RT * a.ncs;                // [2][CL][RT][ncs] receive slots of the cluster exchange (CL > 1)
    uint64_t* xbar = reinterpret_cast<uint64_t*>(rx + (CL > 1 ? 2 * CL * RT * a.ncs : 0));      // [2] transaction barriers (+ pad)
    float* phi_all = reinterpret_cast<float*>(xbar + 2) ;// per layer [RT][ldp]

    const float* X = a.X + chain * a.x_cs;
    const float* Y = a.Y + chain * a.y_cs;
    Xchg xc;
    xc.red = red; xc.rx = rx; xc.bar = xbar; xc.ncs = a.ncs; xc.CL = CL; xc.rank = rank; xc.n = 0;
    if (CL > 1) {
        if (tid == 0) {
            mbar_init(xbar, 1);
            mbar_init(xbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();                              // every CTA's barriers exist before the first remote store
    }
    int tsi = 0;
#define K10_STAMP() do { if (a.timing && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.timing[tsi] = clock64(); ++tsi; } while (0)
    K10_STAMP();

    // ---- set-up.  The global loads this CTA waits for first (its input rows, the length-scales) are issued BEFORE the L2
    //      warm-up loops and parked in registers, so their round trip runs under the prefetch instructions
    constexpr int kXR = 8;
    const int nx = RT * a.d_in;
    const bool x_regs = K10_XS(a) && nx <= kXR * kT;
    float xr[kXR];
    if (x_regs) {
#pragma unroll
        for (int u = 0; u < kXR; ++u) {
            const int e = tid + u * kT;
            xr[u] = 0.f;
            if (e < nx) {                                             // rows of the tile are contiguous in X: element e of the tile
                const int r = e / a.d_in;
                if (row0 + r < a.B) xr[u] = __ldg(X + (int64_t)row0 * a.d_in + e);
            }
        }
    }
    const bool hyp_regs = L * a.dmax <= kT;              // one (layer, input column) per thread: length-scale and mean
    bool hyp_ok = false;
    float ls_r = 0.f, mean_r = 0.f;
    if (hyp_regs && tid < L * a.dmax) {
        const int l = tid / a.dmax, q = tid - l * a.dmax;
        const ClLayer& y = a.layer[l];
        hyp_ok = q < y.d_prev + y.d_x;
        if (hyp_ok) {
            ls_r = __ldg(y.log_inv_ls + chain * a.h_cs + q);
            if (K10_MEAN(y)) mean_r = __ldg(y.mean + chain * a.h_cs + q);
        }
    }
    // ---- L2 warm-up: this CTA's slices of z and W of every layer (first touch after an L2 flush is an HBM round trip)
    for (int l = 0; l < L; ++l) {
        const ClLayer& y = a.layer[l];
        const int d = y.d_prev + y.d_x;
        const int c_lo = rank * y.cols, c_hi = min(y.M, c_lo + y.cols);
        if (c_hi <= c_lo) continue;
        const float* z = y.z + chain * y.z_cs;
        const float* W = y.W + chain * a.w_cs;
        const int lpr = (c_hi - c_lo + 31) >> 5;         // 128-byte lines per z row slice
        for (int e = tid; e < d * lpr; e += kT) {
            const int q = e / lpr, i = e - q * lpr;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(z + (int64_t)q * y.M + c_lo + 32 * i));
        }
        const int nblk = y.kind == DGPRF_KIND_RBF ? 2 : 1;
        const int lw = ((c_hi - c_lo) * y.g + 31) >> 5;
        for (int e = tid; e < nblk * lw; e += kT) {
            const int b = e / lw, i = e - b * lw;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(W + ((int64_t)b * y.M + c_lo) * y.g + 32 * i));
        }
    }
    {   // the targets of this row tile are first needed by the likelihood seed, far down the dependent chain: pull them into L2 now
        const int ycols = K10_GAUSS(a) ? a.d_out : 1;
        const int nline = (RT * ycols + 31) >> 5;
        if (tid < nline && row0 < a.B) asm volatile("prefetch.global.L2 [%0];" ::"l"(Y + (int64_t)row0 * ycols + 32 * tid));
    }
    if (x_regs) {
#pragma unroll
        for (int u = 0; u < kXR; ++u)
            if (tid + u * kT < nx) x_s[tid + u * kT] = xr[u];
    } else if (K10_XS(a))
        for (int e = tid; e < RT * a.d_in; e += kT) {
            const int r = e / a.d_in, q = e - r * a.d_in;
            x_s[e] = (row0 + r) < a.B ? __ldg(X + (int64_t)(row0 + r) * a.d_in + q) : 0.f;
        }
    // wide inputs (the MNIST-shaped configs) are not staged: every layer reads its X columns from L2
    auto x_at = [&](int r, int q) -> float {
        if (K10_XS(a)) return x_s[r * a.d_in + q];
        return (row0 + r) < a.B ? __ldg(X + (int64_t)(row0 + r) * a.d_in + q) : 0.f;
    };
    if (hyp_regs) {
        if (tid < L * a.dmax) { s_all[tid] = hyp_ok ? expf(ls_r) : 0.f; m_all[tid] = mean_r; }
    } else
        for (int e = tid; e < L * a.dmax; e += kT) {
            const int l = e / a.dmax, q = e - l * a.dmax;
            const ClLayer& y = a.layer[l];
            const bool ok = q < y.d_prev + y.d_x;
            s_all[e] = ok ? expf(__ldg(y.log_inv_ls + chain * a.h_cs + q)) : 0.f;
            m_all[e] = (ok && K10_MEAN(y)) ? __ldg(y.mean + chain * a.h_cs + q) : 0.f;
        }
    __syncthreads();

    // A operand of a layer: (in * s) split hi / lo, K zero-padded to a multiple of 8; in = [F_{l-1}, X].  The X part (and
    // the padding) of layer l is written here, the F part by the consumer of the previous layer's exchange.
    auto build_x_part = [&](int l) {
        const ClLayer& y = a.layer[l];
        const int d = y.d_prev + y.d_x, Kp = (d + 7) & ~7, nx = Kp - y.d_prev;
        const float* s_s = s_all + l * a.dmax;
        for (int e = tid; e < RT * nx; e += kT) {
            const int r = e / nx, q = y.d_prev + (e - r * nx);
            const float v = q < d ? x_at(r, q - y.d_prev) * s_s[q] : 0.f;
            uint32_t hi, lo;
            split_tf32(v, hi, lo);
            a_hi[r * a.lda + q] = __uint_as_float(hi);
            a_lo[r * a.lda + q] = __uint_as_float(lo);
        }
    };
    auto build_bias = [&](int l) {                       // bias_r = in_r . mean (needs the complete F_{l-1} in f_s)
        const ClLayer& y = a.layer[l];
        if (tid < RT) {
            float b = 0.f;
            if (K10_MEAN(y)) {
                const float* m_s = m_all + l * a.dmax;
                for (int q = 0; q < y.d_prev; ++q) b = fmaf(f_s[tid * kFS + q], m_s[q], b);
                for (int q = y.d_prev; q < y.d_prev + y.d_x; ++q) b = fmaf(x_at(tid, q - y.d_prev), m_s[q], b);
            }
            bias[tid] = b;
        }
    };
    build_x_part(0);
    build_bias(0);
    __syncthreads();
    K10_STAMP();

    // =========================== forward ===========================
    for (int l = 0; l < L; ++l) {
        const ClLayer& y = a.layer[l];
        const int d = y.d_prev + y.d_x, M = y.M, G = y.g;
        const bool rbf = KIND == 0 ? true : (KIND == 1 ? false : y.kind == DGPRF_KIND_RBF);
        const int Kp = (d + 7) & ~7;
        const int c_lo = rank * y.cols, c_hi = min(M, c_lo + y.cols);
        const int ntile = c_hi > c_lo ? (c_hi - c_lo + 7) >> 3 : 0;
        const float* z = y.z + chain * y.z_cs;
        const float* W = y.W + chain * a.w_cs;
        // keep the two base pointers in registers: under the 128-register cap ptxas otherwise re-derives chain * stride +
        // base (64-bit multiplies and carries) in front of EVERY operand load of the chain
#if (K10_VAR & 1)
        asm volatile("" : "+l"(z), "+l"(W));
#endif
        const float scale = (rbf ? 1.f : 1.41421356237f) * __expf(__ldg(y.log_amp + chain * a.h_cs)) * rsqrtf((float)M);
        float* phi = phi_all + y.phi_off;
        const int NJ = (G + 7) >> 3;
        const int nblk = rbf ? 2 : 1;
        float facc[MT][NJM][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < NJM; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) facc[mt][j][i] = 0.f;

        for (int ti = warp; ti < ntile; ti += kW) {
            const int n0 = c_lo + ti * 8;                 // first P column of this tile
            // W fragments of GEMM #2 for this tile (both blocks), requested before GEMM #1 so they land under it:
            // K slot t <-> feature f0 = n0 + 2t, slot t+4 <-> f0 + 1;  n <-> output column j = 8 jt + g
            // Out-of-range rows / columns are CLAMPED instead of predicated: a clamped W column only feeds output columns
            // >= G (never read), a clamped feature row meets Phi = 0.
            float wv[2][NJM][2];
#if (K10_VAR & 2)
            // The loads are unconditional (clamped indices instead of `if (j < NJ)`): a warp-uniform condition the compiler
            // cannot prove uniform costs a BSSY / BSYNC / BRA region around every pair of loads.
            {
                const int f0 = min(n0 + 2 * t, c_hi - 1), f1 = min(n0 + 2 * t + 1, c_hi - 1);
                const int sb = rbf ? M * G : 0;                          // arc-cosine: block 1 is never used, re-read block 0
                const float* w0 = W + f0 * G;
                const float* w1 = W + f1 * G;
#pragma unroll
                for (int j = 0; j < NJM; ++j) {
                    const int jj = min(j * 8 + g, G - 1);
                    wv[0][j][0] = __ldg(w0 + jj);
                    wv[0][j][1] = __ldg(w1 + jj);
                    wv[1][j][0] = __ldg(w0 + sb + jj);
                    wv[1][j][1] = __ldg(w1 + sb + jj);
                }
            }
#else
            {
                const int f0 = min(n0 + 2 * t, c_hi - 1), f1 = min(n0 + 2 * t + 1, c_hi - 1);
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int j = 0; j < NJM; ++j) {
                        wv[b][j][0] = wv[b][j][1] = 0.f;
                        if (b < nblk && j < NJ) {
                            const int jj = min(j * 8 + g, G - 1);
                            wv[b][j][0] = __ldg(W + (b * M + f0) * G + jj);
                            wv[b][j][1] = __ldg(W + (b * M + f1) * G + jj);
                        }
                    }
            }
#endif
            // ---- GEMM #1: P tile [RT x 8] = A [RT x Kp] . z[:, n0:n0+8]
            float acc[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
            const int col = min(n0 + g, c_hi - 1);        // clamped column: its Phi is zeroed in the epilogue
            for (int k0 = 0; k0 < Kp; k0 += 32) {
                float bz[4][2];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                 // clamped row: the K padding of A is zero
                    const int q = k0 + 8 * u + t;
                    bz[u][0] = __ldg(z + min(q, d - 1) * M + col);
                    bz[u][1] = __ldg(z + min(q + 4, d - 1) * M + col);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kk = k0 + 8 * u;
                    if (kk < Kp) {
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(bz[u][0], bh0, bl0);
                        split_tf32(bz[u][1], bh1, bl1);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const float* ph = a_hi + (mt * 16 + g) * a.lda + kk + t;
                            const float* pl = a_lo + (mt * 16 + g) * a.lda + kk + t;
                            uint32_t ah[4], al[4];
                            ah[0] = __float_as_uint(ph[0]); ah[1] = __float_as_uint(ph[8 * a.lda]);
                            ah[2] = __float_as_uint(ph[4]); ah[3] = __float_as_uint(ph[8 * a.lda + 4]);
                            al[0] = __float_as_uint(pl[0]); al[1] = __float_as_uint(pl[8 * a.lda]);
                            al[2] = __float_as_uint(pl[4]); al[3] = __float_as_uint(pl[8 * a.lda + 4]);
                            mma_3x(acc[mt], ah, al, bh0, bh1, bl0, bl1);
                        }
                    }
                }
            }
            // ---- epilogue in registers: Phi = scale [cos P, sin P] | scale relu(P); saved for the backward
            const int lc = n0 - c_lo + 2 * t;
            float v[MT][4], w[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = mt * 16 + g + (i >> 1) * 8;
                    const bool ok = (n0 + 2 * t + (i & 1)) < c_hi;
                    const float p = acc[mt][i] + bias[r];
                    if (rbf) {
                        float sn, cs;
                        sincos_cw(p, &sn, &cs);
                        v[mt][i] = ok ? scale * cs : 0.f;
                        w[mt][i] = ok ? scale * sn : 0.f;
                    } else {
                        v[mt][i] = ok ? scale * fmaxf(p, 0.f) : 0.f;
                        w[mt][i] = 0.f;
                    }
                }
                float* pr = phi + (mt * 16 + g) * y.ldp + lc;
                *reinterpret_cast<float2*>(pr) = make_float2(v[mt][0], v[mt][1]);
                *reinterpret_cast<float2*>(pr + 8 * y.ldp) = make_float2(v[mt][2], v[mt][3]);
                if (rbf) {
                    *reinterpret_cast<float2*>(pr + y.cols) = make_float2(w[mt][0], w[mt][1]);
                    *reinterpret_cast<float2*>(pr + 8 * y.ldp + y.cols) = make_float2(w[mt][2], w[mt][3]);
                }
            }
            // ---- GEMM #2 partial: F[RT x G] += Phi tile . W rows (accumulator fragment re-used as the A fragment)
#pragma unroll
            for (int b = 0; b < 2; ++b)
                if (b < nblk) {
                    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (b == 0) acc_to_a(v[mt], ah[mt], al[mt]);
                        else acc_to_a(w[mt], ah[mt], al[mt]);
                    }
#pragma unroll
                    for (int j = 0; j < NJM; ++j)
                        if (j < NJ) {
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(wv[b][j][0], bh0, bl0);
                            split_tf32(wv[b][j][1], bh1, bl1);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) mma_3x(facc[mt][j], ah[mt], al[mt], bh0, bh1, bl0, bl1);
                        }
                }
        }
        K10_STAMP();
        {   // F_l = sum of the partials; the owner of element (r, c) also writes the next layer's A operand
            const bool more = l + 1 < L;
            const float* s_n = s_all + (more ? l + 1 : l) * a.dmax;
            reduce_gather<MT, NJM, CL>(facc, NJ, xc, [&](int r, int c, float v) {
                f_s[r * kFS + c] = v;
                if (more && c < G) {
                    uint32_t hi, lo;
                    split_tf32(v * s_n[c], hi, lo);
                    a_hi[r * a.lda + c] = __uint_as_float(hi);
                    a_lo[r * a.lda + c] = __uint_as_float(lo);
                }
            });
            if (more) {
                build_x_part(l + 1);
                if (K10_MEAN(a.layer[l + 1])) __syncthreads();          // the bias needs all of F_l
                build_bias(l + 1);
            }
        }
        __syncthreads();
        K10_STAMP();
    }

    // =========================== likelihood seed ===========================
    // f_s = F_{L-1}; dU/dF_{L-1} overwrites it in place (one thread per row)
    if (tid < RT) {
        const int r = tid;
        const int64_t row = row0 + r;
        float* fr = f_s + r * kFS;
        float ll = 0.f;
        const bool live = row < a.B;
        if (K10_GAUSS(a)) {
            const float llv = __ldg(a.lik_log_var + chain * a.h_cs);
            const float inv_var = expf(-llv);
            for (int j = 0; j < a.d_out; ++j) {
                const float res = live ? __ldg(Y + row * a.d_out + j) - fr[j] : 0.f;
                ll += live ? -0.5f * (DGPRF_LOG_2PI + llv + res * res * inv_var) : 0.f;
                fr[j] = -(res * inv_var) * a.inv_B;
            }
        } else {
            float mx = -INFINITY;
            for (int j = 0; j < a.d_out; ++j) mx = fmaxf(mx, fr[j]);
            float se = 0.f;
            for (int j = 0; j < a.d_out; ++j) se += expf(fr[j] - mx);
            const float lse = mx + logf(se);
            const int label = live ? (int)__ldg(Y + row) : 0;
            ll = live ? ((label >= 0 && label < a.d_out) ? fr[label] : NAN) - lse : 0.f;
            for (int j = 0; j < a.d_out; ++j) {
                const float pj = expf(fr[j] - lse);
                fr[j] = live ? (pj - (j == label ? 1.f : 0.f)) * a.inv_B : 0.f;
            }
        }
        if (RT == 16) ll += __shfl_xor_sync(0x0000ffffu, ll, 8);
        else ll += __shfl_xor_sync(0xffffffffu, ll, 16), ll += __shfl_xor_sync(0xffffffffu, ll, 8);
        ll += __shfl_xor_sync(RT == 16 ? 0x0000ffffu : 0xffffffffu, ll, 4);
        ll += __shfl_xor_sync(RT == 16 ? 0x0000ffffu : 0xffffffffu, ll, 2);
        ll += __shfl_xor_sync(RT == 16 ? 0x0000ffffu : 0xffffffffu, ll, 1);
        if (tid == 0 && rank == 0) a.ll_part[chain * a.ll_cs + tile] = ll;
    }
    __syncthreads();
    K10_STAMP();

    // =========================== backward ===========================
    // element (r, c) of dF_l into the two operand images (zero beyond g: the K / M padding of the MMAs)
    auto put_dF = [&](int r, int c, float v, int NJ, int MJ) {
        uint32_t hi, lo;
        split_tf32(v, hi, lo);
        if (c < NJ * 8) {
            a_hi[r * a.lda + c] = __uint_as_float(hi);
            a_lo[r * a.lda + c] = __uint_as_float(lo);
        }
        if (c < MJ * 16) {
            t_hi[c * LDT + r] = __uint_as_float(hi);
            t_lo[c * LDT + r] = __uint_as_float(lo);
        }
    };
    for (int l = L - 1; l >= 0; --l) {
        const ClLayer& y = a.layer[l];
        const int M = y.M, G = y.g;
        const bool rbf = KIND == 0 ? true : (KIND == 1 ? false : y.kind == DGPRF_KIND_RBF);
        const int NJ = (G + 7) >> 3, MJ = (G + 15) >> 4;
        const int nblk = rbf ? 2 : 1;
        // ---- dF_l operands (a_hi / a_lo: A of dPhi = dF W^T; t_hi / t_lo: dF^T, A of gW^T = dF^T Phi).  The top layer's
        //      come from the likelihood seed in f_s; a lower layer's were written by the consumer of the exchange above
        //      (or, with a trainable mean, from the raw T | R left in f_s: dF = s*T + mean*R needs the whole row)
        if (l == L - 1 || K10_MEAN(a.layer[l + 1])) {
            const bool raw = l < L - 1;
            const float* s_u = s_all + (l + 1) * a.dmax;
            const float* m_u = m_all + (l + 1) * a.dmax;
            for (int e = tid; e < RT * 32; e += kT) {
                const int r = e >> 5, c = e & 31;
                float v = 0.f;
                if (c < G) {
                    v = f_s[r * kFS + c];
                    if (raw) v = fmaf(m_u[c], f_s[r * kFS + G], v * s_u[c]);
                }
                put_dF(r, c, v, NJ, MJ);
            }
            __syncthreads();
        }
        K10_STAMP();

        const int c_lo = rank * y.cols, c_hi = min(M, c_lo + y.cols);
        const int ntile = c_hi > c_lo ? (c_hi - c_lo + 7) >> 3 : 0;
        const float* z = y.z + chain * y.z_cs;
        const float* W = y.W + chain * a.w_cs;
#if (K10_VAR & 1)
        asm volatile("" : "+l"(z), "+l"(W));
#endif
        const float* phi = phi_all + y.phi_off;
        float* gw = a.gwpart + chain * a.gw_cs + (int64_t)tile * a.gw_ss + y.off_W;
        const float arc_scale = 1.41421356237f * __expf(__ldg(y.log_amp + chain * a.h_cs)) * rsqrtf((float)M);
        const int nq_cols = y.d_prev + (K10_MEAN(y) ? 1 : 0);      // T columns (+ the row-sum column)
        const int NQ = l > 0 ? (nq_cols + 7) >> 3 : 0;
        float tacc[MT][NJM][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < NJM; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) tacc[mt][j][i] = 0.f;

        for (int ti = warp; ti < ntile; ti += kW) {
            const int n0 = c_lo + ti * 8;
            const int lc = n0 - c_lo;
            if (l > 0) {
                // z fragments of the T GEMM, requested first: K slot t <-> column n0 + 2t, slot t+4 <-> n0 + 2t + 1; n <-> q
                float zv[NJM][2];
#pragma unroll
                for (int j = 0; j < NJM; ++j) {
                    const int q = j * 8 + g, c0 = n0 + 2 * t;
                    float b0 = 0.f, b1 = 0.f;
                    if (j < NQ) {                             // clamped row / column: dP is zero beyond c_hi, T columns >= d_prev are never read
                        const float* zp = z + min(q, y.d_prev - 1) * M;
                        b0 = __ldg(zp + min(c0, c_hi - 1));
                        b1 = __ldg(zp + min(c0 + 1, c_hi - 1));
                        if (q == y.d_prev && K10_MEAN(y)) b0 = b1 = 1.f;      // the row-sum column R
                    }
                    zv[j][0] = b0; zv[j][1] = b1;
                }
                // ---- dPhi = dF . W^T for the tile's features: n <-> feature fb + n0 + g, k <-> j
                float dacc[2][MT][4];
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) dacc[b][mt][i] = 0.f;
                const int fcl = min(n0 + g, c_hi - 1);        // clamped feature row of W: its dP is zero (Phi = 0 there)
#pragma unroll
                for (int j = 0; j < NJM; ++j)
                    if (j < NJ) {
                        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const float* ph = a_hi + (mt * 16 + g) * a.lda + j * 8 + t;
                            const float* pl = a_lo + (mt * 16 + g) * a.lda + j * 8 + t;
                            ah[mt][0] = __float_as_uint(ph[0]); ah[mt][1] = __float_as_uint(ph[8 * a.lda]);
                            ah[mt][2] = __float_as_uint(ph[4]); ah[mt][3] = __float_as_uint(ph[8 * a.lda + 4]);
                            al[mt][0] = __float_as_uint(pl[0]); al[mt][1] = __float_as_uint(pl[8 * a.lda]);
                            al[mt][2] = __float_as_uint(pl[4]); al[mt][3] = __float_as_uint(pl[8 * a.lda + 4]);
                        }
#pragma unroll
                        for (int b = 0; b < 2; ++b)
                            if (b < nblk) {
                                const int j0 = j * 8 + t;     // clamped k: the K padding of the dF operand is zero
                                const float* wp = W + (b * M + fcl) * G;
                                const float w0 = __ldg(wp + min(j0, G - 1));
                                const float w1 = __ldg(wp + min(j0 + 4, G - 1));
                                uint32_t bh0, bl0, bh1, bl1;
                                split_tf32(w0, bh0, bl0);
                                split_tf32(w1, bh1, bl1);
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt) mma_3x(dacc[b][mt], ah[mt], al[mt], bh0, bh1, bl0, bl1);
                            }
                    }
                // ---- dP from the saved features (same fragment layout), then T += dP . z^T
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const float* pr = phi + (mt * 16 + g) * y.ldp + lc + 2 * t;
                    const float2 c01 = *reinterpret_cast<const float2*>(pr);
                    const float2 c23 = *reinterpret_cast<const float2*>(pr + 8 * y.ldp);
                    float dp[4];
                    if (rbf) {
                        const float2 s01 = *reinterpret_cast<const float2*>(pr + y.cols);
                        const float2 s23 = *reinterpret_cast<const float2*>(pr + 8 * y.ldp + y.cols);
                        dp[0] = c01.x * dacc[1][mt][0] - s01.x * dacc[0][mt][0];
                        dp[1] = c01.y * dacc[1][mt][1] - s01.y * dacc[0][mt][1];
                        dp[2] = c23.x * dacc[1][mt][2] - s23.x * dacc[0][mt][2];
                        dp[3] = c23.y * dacc[1][mt][3] - s23.y * dacc[0][mt][3];
                    } else {
                        dp[0] = c01.x > 0.f ? dacc[0][mt][0] * arc_scale : 0.f;
                        dp[1] = c01.y > 0.f ? dacc[0][mt][1] * arc_scale : 0.f;
                        dp[2] = c23.x > 0.f ? dacc[0][mt][2] * arc_scale : 0.f;
                        dp[3] = c23.y > 0.f ? dacc[0][mt][3] * arc_scale : 0.f;
                    }
                    uint32_t ah[4], al[4];
                    acc_to_a(dp, ah, al);
#pragma unroll
                    for (int j = 0; j < NJM; ++j)
                        if (j < NQ) {
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(zv[j][0], bh0, bl0);
                            split_tf32(zv[j][1], bh1, bl1);
                            mma_3x(tacc[mt][j], ah, al, bh0, bh1, bl0, bl1);
                        }
                }
            }
            // ---- gW^T [G x 8 features] = dF^T [G x RT] . Phi tile [RT x 8]: A from t_hi / t_lo, B from the saved tile
            __syncwarp();
#pragma unroll
            for (int b = 0; b < 2; ++b)
                if (b < nblk) {
                    float gacc[(NJM + 1) / 2][4];
#pragma unroll
                    for (int mj = 0; mj < (NJM + 1) / 2; ++mj)
#pragma unroll
                        for (int i = 0; i < 4; ++i) gacc[mj][i] = 0.f;
#pragma unroll
                    for (int kr = 0; kr < RT / 8; ++kr) {
                        const float* pb = phi + (kr * 8 + t) * y.ldp + b * y.cols + lc + g;
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(pb[0], bh0, bl0);
                        split_tf32(pb[4 * y.ldp], bh1, bl1);
#pragma unroll
                        for (int mj = 0; mj < (NJM + 1) / 2; ++mj)
                            if (mj < MJ) {
                                const float* ph = t_hi + (mj * 16 + g) * LDT + kr * 8 + t;
                                const float* pl = t_lo + (mj * 16 + g) * LDT + kr * 8 + t;
                                uint32_t ah[4], al[4];
                                ah[0] = __float_as_uint(ph[0]); ah[1] = __float_as_uint(ph[8 * LDT]);
                                ah[2] = __float_as_uint(ph[4]); ah[3] = __float_as_uint(ph[8 * LDT + 4]);
                                al[0] = __float_as_uint(pl[0]); al[1] = __float_as_uint(pl[8 * LDT]);
                                al[2] = __float_as_uint(pl[4]); al[3] = __float_as_uint(pl[8 * LDT + 4]);
                                mma_3x(gacc[mj], ah, al, bh0, bh1, bl0, bl1);
                            }
                    }
#pragma unroll
                    for (int mj = 0; mj < (NJM + 1) / 2; ++mj)
                        if (mj < MJ) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int j = mj * 16 + g + (i >> 1) * 8, c = n0 + 2 * t + (i & 1);
                                if (j < G && c < c_hi) gw[((int64_t)b * M + c) * G + j] = gacc[mj][i];
                            }
                        }
                }
        }
        K10_STAMP();
        if (l == 0) break;
        {   // T_l (| R_l) = sum of the partials; without a trainable mean the owner of (r, c) writes dF_{l-1} = s * T straight
            // into the operand images of the next backward step
            const ClLayer& yp = a.layer[l - 1];
            const int Gp = yp.g, NJp = (Gp + 7) >> 3, MJp = (Gp + 15) >> 4;
            const float* s_l = s_all + l * a.dmax;
            const bool direct = !K10_MEAN(y);
            reduce_gather<MT, NJM, CL>(tacc, NQ, xc, [&](int r, int c, float v) {
                if (direct) put_dF(r, c, c < Gp ? v * s_l[c] : 0.f, NJp, MJp);
                else f_s[r * kFS + c] = v;
            });
            if (direct)                                   // operand rows / columns beyond the exchanged width
                for (int e = tid; e < RT * (32 - NQ * 8); e += kT) {
                    const int r = e / (32 - NQ * 8), c = NQ * 8 + (e - r * (32 - NQ * 8));
                    put_dF(r, c, 0.f, NJp, MJp);
                }
        }
        __syncthreads();
        K10_STAMP();
    }

    // =========================== fused update (cooperative launch only) ===========================
    if (a.fuse_update) {
        // Everything of the update that does not depend on the other CTAs' gradient slabs runs BEFORE the grid barrier:
        // theta / momentum loads and the Philox draws of this CTA's first pass of vectors hide under the wait.
        const int64_t n4 = a.upd.n >> 2;
        const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;          // 128-bit vectors per CTA
        const int64_t v0 = (int64_t)blockIdx.x * per, v1 = min(n4, v0 + per);
        const float* grad = a.upd.grad + chain * a.upd.grad_cs;
        const bool lpv8 = K10_LPV8(a);
        const int sub = lpv8 ? (tid & 7) : 0;
        const int64_t vf = v0 + (lpv8 ? (tid >> 3) : tid);             // this thread's vector of the first pass
        float4 th0 = make_float4(0.f, 0.f, 0.f, 0.f), mo0 = th0, e0 = th0;
        if (sub == 0 && vf < v1) {
            th0 = *reinterpret_cast<const float4*>(a.upd.theta + chain * a.upd.cs + (vf << 2));
            mo0 = *reinterpret_cast<const float4*>(a.upd.mom + chain * a.upd.cs + (vf << 2));
            sgmcmc_draw_vec(a.upd, chain, vf, e0, mo0);
        }
        grid_barrier_cl(a.bar + 2, gridDim.x * gridDim.y);      // word 2 of the barrier block (K9 owns words 0 and 1)
        K10_STAMP();
        if (a.u_out != nullptr && blockIdx.x == 0 && tid >= kT - 32) {  // minibatch log-likelihood, fixed order: on the last warp,
            const int n_tiles = gridDim.x / CL;                         // which has no vector of the update to do in most shapes
            const int ln = tid - (kT - 32);
            float sll = 0.f;
            for (int i = ln; i < n_tiles; i += 32) sll += __ldcg(a.ll_part + chain * a.ll_cs + i);
            sll = warp_sum(sll);
            if (ln == 0) a.u_out[chain] = sll;
        }
        if (lpv8) {
            for (int64_t vb = v0; vb < v1; vb += kT / 8) {
                const int64_t v = vb + (tid >> 3);
                float4 gr = make_float4(0.f, 0.f, 0.f, 0.f), th = th0, mo = mo0, e = e0;
                if (vb != v0 && sub == 0 && v < v1) {
                    th = *reinterpret_cast<const float4*>(a.upd.theta + chain * a.upd.cs + (v << 2));
                    mo = *reinterpret_cast<const float4*>(a.upd.mom + chain * a.upd.cs + (v << 2));
                    sgmcmc_draw_vec(a.upd, chain, v, e, mo);
                }
                if (v < v1) gr = slab_sum_lane<8>(grad, a.upd.part_stride, a.upd.n_part, sub, v << 2);
                gr = shuffle_sum_lpv<8>(gr);
                if (sub == 0 && v < v1) sgmcmc_apply_vec(a.upd, tab, chain, v, gr, th, mo, e);
            }
        } else {
            for (int64_t v = vf; v < v1; v += kT) {
                float4 th = th0, mo = mo0, e = e0;
                if (v != vf) {
                    th = *reinterpret_cast<const float4*>(a.upd.theta + chain * a.upd.cs + (v << 2));
                    mo = *reinterpret_cast<const float4*>(a.upd.mom + chain * a.upd.cs + (v << 2));
                    sgmcmc_draw_vec(a.upd, chain, v, e, mo);
                }
                const float4 gr = slab_sum_lane<1>(grad, a.upd.part_stride, a.upd.n_part, 0, v << 2);
                sgmcmc_apply_vec(a.upd, tab, chain, v, gr, th, mo, e);
            }
        }
        K10_STAMP();
    }
#undef K10_STAMP
}

}  // namespace

template <int MT, int NJM, int KIND, int CLT>
static int ensure_smem(size_t smem) { return dgprf_ensure_smem((const void*)k10_step_cluster<MT, NJM, KIND, CLT>, smem); }

template <int MT, int NJM, int KIND, int CLT>
static int launch_cl(const ClArgs& a, const SegTable& tab, dim3 grid, size_t smem, bool coop, cudaStream_t st) {
    const int rc0 = ensure_smem<MT, NJM, KIND, CLT>(smem);
    if (rc0) return rc0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (a.CL > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = a.CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (coop) {
        at[na].id = cudaLaunchAttributeCooperative;
        at[na].val.cooperative = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k10_step_cluster<MT, NJM, KIND, CLT>, a, tab);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        dgprf_set_error("k10_step_cluster launch failed: %s (grid %u x %u, cluster %d, smem %zu, cooperative %d)",
                        cudaGetErrorString(e), grid.x, grid.y, a.CL, smem, (int)coop);
        return DGPRF_ECUDA;
    }
    return DGPRF_OK;
}

static int g_dbg_calls = 0;
static int dbg_calls_peek() { return g_dbg_calls; }

template <int MT, int NJM, int KIND, int CLT>
static int max_coresident(int CL, size_t smem) {
    if (ensure_smem<MT, NJM, KIND, CLT>(smem) != DGPRF_OK) return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CL * 64); cfg.blockDim = dim3(kT); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k10_step_cluster<MT, NJM, KIND, CLT>, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n * CL;
}


// one translation unit per cluster size: (MT, NJM, KIND) -> instantiation of that cluster size
#define K10_SHAPES(X, kd, cl) X(1, 1, kd, cl) X(1, 2, kd, cl) X(1, 4, kd, cl) X(2, 1, kd, cl) X(2, 2, kd, cl) X(2, 4, kd, cl)
#define K10_DEFINE_CL(cl)                                                                                                          \
    int dgprf_k10_launch_cl##cl(int MT, int NJM, int KIND, const ClArgs& a, const SegTable& tab, dim3 grid, size_t smem, bool coop, \
                                cudaStream_t st) {                                                                               \
        K10_SHAPES(K10_LAUNCH_CASE, 0, cl) K10_SHAPES(K10_LAUNCH_CASE, 2, cl)                                                      \
        dgprf_set_error("k10: no instantiation for MT=%d NJM=%d KIND=%d CL=%d", MT, NJM, KIND, cl);                               \
        return DGPRF_EINVAL;                                                                                                       \
    }                                                                                                                              \
    int dgprf_k10_coresident_cl##cl(int MT, int NJM, int KIND, size_t smem) {                                                      \
        K10_SHAPES(K10_CORES_CASE, 0, cl) K10_SHAPES(K10_CORES_CASE, 2, cl)                                                        \
        return 0;                                                                                                                  \
    }
#define K10_LAUNCH_CASE(mt, nj, kd, cl) if (MT == mt && NJM == nj && KIND == kd) return launch_cl<mt, nj, kd, cl>(a, tab, grid, smem, coop, st);
#define K10_CORES_CASE(mt, nj, kd, cl) if (MT == mt && NJM == nj && KIND == kd) return max_coresident<mt, nj, kd, cl>(cl, smem);
#define K10_DECLARE_CL(cl)                                                                                                         \
    int dgprf_k10_launch_cl##cl(int MT, int NJM, int KIND, const ClArgs& a, const SegTable& tab, dim3 grid, size_t smem, bool coop, \
                                cudaStream_t st);                                                                                \
    int dgprf_k10_coresident_cl##cl(int MT, int NJM, int KIND, size_t smem);
K10_DECLARE_CL(1) K10_DECLARE_CL(2) K10_DECLARE_CL(4) K10_DECLARE_CL(8)
