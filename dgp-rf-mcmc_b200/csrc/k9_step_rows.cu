// K9: row-fused sampling step for small minibatches (fp32).
//
// One CTA owns R = 8 batch rows and carries them through EVERY layer: forward (features kept in shared
// memory, never written to HBM), likelihood seed, and the whole reverse pass.  Nothing crosses CTAs except
// the W gradient, which each CTA writes as one partial slab (summed in fixed order by K5).  A step is then
// two launches (this kernel + the update) instead of 2L+2, and no phase waits on another CTA -- the layered
// kernels (K1/K2) are bound by exactly those dependent round trips at BASELINE configs[1] sizes.
//
// Arithmetic (same formulas as K1/K2; reference lines there):
//   fwd  P = in (s*z + mean);  Phi = amp/sqrt(M)[cos P, sin P] | sqrt(2) amp/sqrt(M) relu(P);  F = Phi W
//   lik  dU/dF seed as K3
//   bwd  dPhi = dF W^T;  gW = Phi^T dF;  dP;  T = dP z^T;  dF_prev = s*T + mean*rowsum(dP)
// Shared-memory layouts: the small per-row vectors are "transposed" (X_t[q][r], F_t[j][r], dF_t[j][r]) so a
// thread that owns a feature column reads all R rows with two broadcast float4 loads; the wide matrices are
// row-major with a padded stride (Phi[r][f], dPhi/dP[r][f]) and the staged W / z operands are K-contiguous
// (Wt[j][f], z[q][m]), so the K-reductions run on float4 loads along K.
#include <stdio.h>
#include <stdlib.h>
#include "kernels.cuh"
#include "update_core.cuh"

constexpr int kSR = 8;           // rows per CTA
constexpr int kST = 512;         // threads per CTA
constexpr int kSW = kST / 32;    // warps
constexpr int kSN = 64;          // widest small-GEMM output (n_gp, d_prev <= 64)

struct StepRowsLayer {
    int32_t kind, d_prev, d_x, M, g, has_mean;
    const float* z; int64_t z_cs;
    const float* log_inv_ls; const float* log_amp; const float* mean;   // + chain*h_cs
    const float* W;                                                     // + chain*w_cs
    int64_t off_W;                                                      // into a gradient slab
    int32_t phi_off;                                                    // float offset of Phi_l in shared memory
};

struct StepRowsArgs {
    int32_t n_layers, likelihood, B, d_in, d_out, dmax, Fmax, bs_cap, prefetch_w, zr_cap, prefetch_bwd;
    int64_t h_cs, w_cs;
    const float* X; int64_t x_cs;
    const float* Y; int64_t y_cs;
    const float* lik_log_var;
    float* gwpart; int64_t gw_cs, gw_ss;     // [C][n_groups][w_len]
    float* ll_part; int64_t ll_cs;           // [C][n_groups]
    float inv_B;
    long long* timing;                       // debug: phase timestamps of CTA 0 (nullable)
    int32_t fuse_update;                     // 1: grid barrier, then every CTA updates its slice of the W buffer
    unsigned int* bar;                       // [2] {arrival count, generation} in the workspace (zero-initialised)
    float* u_out;                            // [C] sum_i ll_i (nullable; fused path only)
    UpdArgs upd;
    StepRowsLayer layer[DGPRF_MAX_LAYERS];
};

__device__ __forceinline__ void ld8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// Cooperative async copy of n contiguous floats global -> shared (16-byte pieces when aligned).
__device__ __forceinline__ void stage_async(float* dst, const float* src, int n) {
    const int tid = threadIdx.x;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const int n4 = n >> 2;
        for (int e = tid; e < n4; e += kST) cp_async16(dst + 4 * e, src + 4 * e, true);
        for (int e = 4 * n4 + tid; e < n; e += kST) cp_async4(dst + e, src + e, true);
    } else {
        for (int e = tid; e < n; e += kST) cp_async4(dst + e, src + e, true);
    }
}
// rows [f0, f0+kc) of W [F, g] -> bs[ff][g] (natural layout, one contiguous copy); rows kc..kp are zeroed.
// A 4-byte-granular transposing copy would put ~g LDGSTS per thread in front of every other memory
// operation of the phase; the contiguous copy is F*g/4/512 (~4) 16-byte LDGSTS per thread.
__device__ __forceinline__ void stage_W_rows(float* bs, const float* W, int f0, int kc, int kp, int g) {
    stage_async(bs, W + (int64_t)f0 * g, kc * g);
    for (int e = kc * g + threadIdx.x; e < kp * g; e += kST) bs[e] = 0.f;
}
// the first N rows of z [d, M], columns [k0, k0+kc) -> bs[n][ld]; columns kc..kp zero-filled
__device__ __forceinline__ void stage_z_rows(float* bs, int ld, const float* z, int64_t M, int k0, int kc, int kp, int N) {
    for (int n = 0; n < N; ++n) {
        const float* src = z + (int64_t)n * M + k0;
        float* dst = bs + n * ld;
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            for (int e = threadIdx.x; e < (kp >> 2); e += kST) {
                if (4 * e + 3 < kc) cp_async16(dst + 4 * e, src + 4 * e, true);
                else
                    for (int i = 0; i < 4; ++i) cp_async4(dst + 4 * e + i, src + 4 * e + i, 4 * e + i < kc);
            }
        } else {
            for (int e = threadIdx.x; e < kp; e += kST) cp_async4(dst + e, src + e, e < kc);
        }
    }
}

// out_t[n][r] = sum_k A[r][k] * Bm(k, n)  for n < N (<= 64) and the R rows; A is row-major in shared memory
// (stride lda, K padded with zeros to a multiple of 4), B is staged K-contiguous through `bs` in K chunks.
// lane = (r = lane / 4, c4 = lane % 4) owns columns n = c4, c4+4, ... (NQ of them); the 16 warps split K in
// float4 steps; the cross-warp reduction is in fixed order (deterministic).
// WMODE=true: B = W [K, N] row-major (staged as is: bs[k][N]); false: B(k,n) = Bg[n*ldb + k] (z rows: bs[n][k]).
// extra_ones: out_t[N][r] = sum_k A[r][k] (row sums, for the mean term).
template <int NQ, bool WMODE>
__device__ __forceinline__ void small_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bg, int64_t ldb,
                                           int K, int N, bool extra_ones, float* __restrict__ bs, int bs_cap,
                                           bool first_chunk_staged, bool keep_newest, float* __restrict__ red, float* __restrict__ out_t) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, c4 = lane & 3;
    float acc[NQ][2], ones = 0.f;                            // two independent FMA chains per column
#pragma unroll
    for (int i = 0; i < NQ; ++i) acc[i][0] = acc[i][1] = 0.f;
    const int K4 = (K + 3) & ~3;
    const int kc_max = min(K4, ((bs_cap / N) - 4) & ~3);       // chunk length (multiple of 4); z-row stride kc_max + 4
    const int ldbs = kc_max + 4;
    for (int k0 = 0; k0 < K; k0 += kc_max) {
        const int kc = min(kc_max, K - k0), kp = min(kc_max, K4 - k0);
        if (!(first_chunk_staged && k0 == 0)) {
            if (WMODE) stage_W_rows(bs, Bg, k0, kc, kp, N);
            else stage_z_rows(bs, ldbs, Bg, ldb, k0, kc, kp, N);
            cp_async_commit();
            cp_async_wait_all();
        } else if (keep_newest) {
            cp_async_wait_but_newest();           // a prefetch for a LATER phase was committed after this operand
        } else {
            cp_async_wait_all();
        }
        __syncthreads();
        const float* Ar = A + r * lda + k0;
#pragma unroll 2
        for (int k = 4 * warp; k < kp; k += 4 * kSW) {
            const float4 a4 = *reinterpret_cast<const float4*>(Ar + k);
            if (k + 3 < kc) ones += (a4.x + a4.y) + (a4.z + a4.w);      // A may hold unrelated data in its K padding
            else ones += (k < kc ? a4.x : 0.f) + (k + 1 < kc ? a4.y : 0.f) + (k + 2 < kc ? a4.z : 0.f);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int n = c4 + 4 * i;
                if (n < N) {
                    float4 b4;
                    if (WMODE) {
                        const float* bp = bs + k * N + n;
                        b4 = make_float4(bp[0], bp[N], bp[2 * N], bp[3 * N]);
                    } else {
                        b4 = *reinterpret_cast<const float4*>(bs + n * ldbs + k);
                    }
                    acc[i][0] = fmaf(a4.x, b4.x, fmaf(a4.y, b4.y, acc[i][0]));
                    acc[i][1] = fmaf(a4.z, b4.z, fmaf(a4.w, b4.w, acc[i][1]));
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int n = c4 + 4 * i;
        if (n < N) red[(warp * (kSN + 1) + n) * kSR + r] = acc[i][0] + acc[i][1];
    }
    if (extra_ones && c4 == 0) red[(warp * (kSN + 1) + kSN) * kSR + r] = ones;
    __syncthreads();
    const int tot = (N + (extra_ones ? 1 : 0)) * kSR;
    for (int e = tid; e < tot; e += kST) {
        const int n = e / kSR, rr = e % kSR;
        const int src = n < N ? n : kSN;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kSW; ++w) s += red[(w * (kSN + 1) + src) * kSR + rr];
        out_t[n * kSR + rr] = s;
    }
    __syncthreads();
}

template <bool WMODE>
__device__ __forceinline__ void small_gemm_dispatch(const float* A, int lda, const float* Bg, int64_t ldb, int K, int N,
                                                    bool extra_ones, float* bs, int bs_cap, bool staged, bool keep_newest,
                                                    float* red, float* out_t) {
    const int nq = (N + 3) >> 2;
    if (nq <= 1) small_gemm<1, WMODE>(A, lda, Bg, ldb, K, N, extra_ones, bs, bs_cap, staged, keep_newest, red, out_t);
    else if (nq <= 3) small_gemm<3, WMODE>(A, lda, Bg, ldb, K, N, extra_ones, bs, bs_cap, staged, keep_newest, red, out_t);
    else if (nq <= 8) small_gemm<8, WMODE>(A, lda, Bg, ldb, K, N, extra_ones, bs, bs_cap, staged, keep_newest, red, out_t);
    else small_gemm<16, WMODE>(A, lda, Bg, ldb, K, N, extra_ones, bs, bs_cap, staged, keep_newest, red, out_t);
}
// Stage chunk 0 of the z-row operand of small_gemm<.., false> ahead of time (same layout computation); returns
// whether the whole operand is that one chunk (only then may the caller pass first_chunk_staged = true).
__device__ __forceinline__ bool small_gemm_prefetch_z(const float* Bg, int64_t ldb, int K, int N, float* bs, int bs_cap) {
    const int K4 = (K + 3) & ~3;
    const int kc_max = min(K4, ((bs_cap / N) - 4) & ~3);
    if (kc_max < K4) return false;
    stage_z_rows(bs, kc_max + 4, Bg, ldb, 0, K, K4, N);
    return true;
}

// out_t[n][r] = sum_k Phi[r][k] W[k][n] with the WHOLE W (natural layout, rows >= K zero-filled up to K4) staged in
// `bs` and the output width N a compile-time constant: four consecutive W rows are N contiguous float4, so a lane
// reads them with N 128-bit loads and knows at compile time which (k, n) every component is.  lane = (row pair
// rp = lane / 8, k-slot ks = lane % 8): 2 rows x N columns x 4 k per step = 8N FMAs for N + 2 shared loads
// (the generic small_gemm issues 13 loads per 12 FMAs).  Reduction: shuffles over the 8 k-slots, then the same
// fixed-order cross-warp sum as small_gemm.
template <int N>
__device__ __forceinline__ void phi_w_gemm(const float* __restrict__ A, int lda, int K, const float* __restrict__ bs,
                                           float* __restrict__ red, float* __restrict__ out_t) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rp = lane >> 3, ks = lane & 7;
    float acc0[N], acc1[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc0[n] = acc1[n] = 0.f;
    const int n4 = ((K + 3) & ~3) >> 2;                     // float4 steps along K
    const float* A0 = A + (2 * rp) * lda;
    const float* A1 = A0 + lda;
    for (int k4 = warp * 8 + ks; k4 < n4; k4 += kSW * 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(A0 + 4 * k4);
        const float4 a1 = *reinterpret_cast<const float4*>(A1 + 4 * k4);
        const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
        const float4* wp = reinterpret_cast<const float4*>(bs + (int64_t)k4 * 4 * N);
        float w[4 * N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float4 t = wp[j];
            w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int e = 0; e < 4 * N; ++e) {
            acc0[e % N] = fmaf(av0[e / N], w[e], acc0[e % N]);
            acc1[e % N] = fmaf(av1[e / N], w[e], acc1[e % N]);
        }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
            acc0[n] += __shfl_xor_sync(0xffffffffu, acc0[n], o);
            acc1[n] += __shfl_xor_sync(0xffffffffu, acc1[n], o);
        }
    }
    if (ks == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
            red[(warp * (kSN + 1) + n) * kSR + 2 * rp] = acc0[n];
            red[(warp * (kSN + 1) + n) * kSR + 2 * rp + 1] = acc1[n];
        }
    }
    __syncthreads();
    for (int e = tid; e < N * kSR; e += kST) {
        const int n = e / kSR, rr = e % kSR;
        float sum = 0.f;
#pragma unroll
        for (int wq = 0; wq < kSW; ++wq) sum += red[(wq * (kSN + 1) + n) * kSR + rr];
        out_t[n * kSR + rr] = sum;
    }
    __syncthreads();
}
// true if handled (N <= 16); the operand must already be complete in `bs` (caller has waited and synchronised)
__device__ __forceinline__ bool phi_w_gemm_dispatch(int N, const float* A, int lda, int K, const float* bs, float* red, float* out_t) {
    switch (N) {
#define DGPRF_PW(n) case n: phi_w_gemm<n>(A, lda, K, bs, red, out_t); return true;
        DGPRF_PW(1) DGPRF_PW(2) DGPRF_PW(3) DGPRF_PW(4) DGPRF_PW(5) DGPRF_PW(6) DGPRF_PW(7) DGPRF_PW(8)
        DGPRF_PW(9) DGPRF_PW(10) DGPRF_PW(11) DGPRF_PW(12) DGPRF_PW(13) DGPRF_PW(14) DGPRF_PW(15) DGPRF_PW(16)
#undef DGPRF_PW
        default: return false;
    }
}

__device__ __forceinline__ int padded_F(int F) { return ((F + 3) & ~3) + 4; }

// Grid-wide barrier for a cooperative launch (all CTAs co-resident).  Self-resetting: the last arriver zeroes
// the count and bumps the generation; `my_gen` was read before this CTA arrived.  Bounded spin: a lost
// arrival traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int my_gen, unsigned int n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int old = atomicAdd(bar, 1u);
        if (old == n_ctas - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned int*>(bar + 1) == my_gen)
                if (clock64() - t0 > 4000000000LL) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kST, 1) k9_step_rows(const __grid_constant__ StepRowsArgs a, const __grid_constant__ SegTable tab) {
    extern __shared__ __align__(16) float sm[];
    unsigned int my_gen = 0;
    if (a.fuse_update && threadIdx.x == 0) my_gen = *reinterpret_cast<volatile unsigned int*>(a.bar + 1);
    float* x_t   = sm;                                   // [d_in][R]
    float* fa_t  = x_t + a.d_in * kSR;                   // [kSN+1][R]  layer output / dF (ping)
    float* fb_t  = fa_t + (kSN + 1) * kSR;               // [kSN+1][R]  (pong)
    float* s_all = fb_t + (kSN + 1) * kSR;               // [L][dmax] exp(log_inv_ls)
    float* m_all = s_all + a.n_layers * a.dmax;          // [L][dmax] mean
    float* red   = m_all + a.n_layers * a.dmax;          // [kSW][kSN+1][R]
    float* dphi  = red + kSW * (kSN + 1) * kSR;          // [R][Fmax_p]  dPhi, then dP in place (cos half)
    float* b_s   = dphi + kSR * a.Fmax;                  // [bs_cap] staging of W / z operands
    float* zr_s  = b_s + a.bs_cap;                       // [zr_cap] z rows of the backward T GEMM (prefetched), may be empty
    float* phi_all = zr_s + a.zr_cap;                    // per layer [R][F_p]
    float* zf_s  = dphi;                                 // forward only: the first zf_rows rows of z_l (dphi is idle then)

    const int tid = threadIdx.x;
    const int chain = blockIdx.y, grp = blockIdx.x;
    const int row0 = grp * kSR;
    const float* X = a.X + chain * a.x_cs;
    const float* Y = a.Y + chain * a.y_cs;

    // Warm L2 with every layer's operands (z, W) once, spread over the grid: after an L2 flush (or the first step) each
    // later phase would otherwise pay its own HBM round trip on first touch.
    if (a.prefetch_w) {
        const int64_t gtid = (int64_t)blockIdx.x * kST + tid, gthreads = (int64_t)gridDim.x * kST;
        for (int l = 0; l < a.n_layers; ++l) {
            const StepRowsLayer& y = a.layer[l];
            const int F = y.kind == DGPRF_KIND_RBF ? 2 * y.M : y.M;
            const float* zz = y.z + chain * y.z_cs;
            const float* ww = y.W + chain * a.w_cs;
            const int64_t nz = ((int64_t)(y.d_prev + y.d_x) * y.M + 31) / 32, nw = ((int64_t)F * y.g + 31) / 32;   // 128-byte lines
            for (int64_t i = gtid; i < nz; i += gthreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(zz + 32 * i));
            for (int64_t i = gtid; i < nw; i += gthreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(ww + 32 * i));
        }
    }
    for (int e = tid; e < a.d_in * kSR; e += kST) {
        const int q = e / kSR, r = e % kSR;
        x_t[e] = (row0 + r) < a.B ? __ldg(X + (int64_t)(row0 + r) * a.d_in + q) : 0.f;
    }
    for (int e = tid; e < a.n_layers * a.dmax; e += kST) {       // exp(log_inv_ls), mean of every layer, once
        const int l = e / a.dmax, q = e % a.dmax;
        const StepRowsLayer& y = a.layer[l];
        const bool ok = q < y.d_prev + y.d_x;
        s_all[e] = ok ? expf(__ldg(y.log_inv_ls + chain * a.h_cs + q)) : 0.f;
        m_all[e] = (ok && y.has_mean) ? __ldg(y.mean + chain * a.h_cs + q) : 0.f;
    }
    float* fcur = fa_t;      // F_{l-1} (transposed)
    float* fnext = fb_t;
    // Operand prefetch (cp.async groups, oldest first): the leading rows of z_l sit in zf_s before layer l starts --
    // z_0 is requested here, z_{l+1} right after GEMM #1 of layer l, so the L2 round trip hides under the Phi.W GEMM.
    if (a.prefetch_w) {
        const StepRowsLayer& y0 = a.layer[0];
        const int nz0 = min(y0.d_prev + y0.d_x, (a.Fmax * kSR) / y0.M);
        stage_async(zf_s, y0.z + chain * y0.z_cs, nz0 * y0.M);
        cp_async_commit();
    }
    bool top_w_in_bs = false;                            // the last layer's W is still staged when the backward starts
    int tsi = 0;
#define K9_STAMP() do { if (a.timing && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.timing[tsi] = clock64(); ++tsi; } while (0)
#define K9_STAMP2() do { if (a.timing && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) a.timing[40 + l] = clock64(); } while (0)
    K9_STAMP();

    // =========================== forward ===========================
    for (int l = 0; l < a.n_layers; ++l) {
        const StepRowsLayer& y = a.layer[l];
        const int d = y.d_prev + y.d_x, M = y.M;
        const bool rbf = y.kind == DGPRF_KIND_RBF;
        const int F = rbf ? 2 * M : M, Fp = padded_F(F);
        const float* z = y.z + chain * y.z_cs;
        const float* s_s = s_all + l * a.dmax;
        const float* m_s = m_all + l * a.dmax;
        const float scale = (rbf ? 1.f : 1.41421356237f) * __expf(__ldg(y.log_amp + chain * a.h_cs)) * rsqrtf((float)M);
        float* phi = phi_all + y.phi_off;
        const float* Wl = y.W + chain * a.w_cs;
        const int F4 = (F + 3) & ~3;
        const bool w_fits = a.prefetch_w && (F4 + 4) * y.g <= a.bs_cap;  // whole W_l staged while GEMM #1 runs
        const int nzp = a.prefetch_w ? min(d, (a.Fmax * kSR) / M) : 0;   // rows of z_l prefetched into zf_s
        if (w_fits) {
            stage_W_rows(b_s, Wl, 0, F, F4, y.g);
            cp_async_commit();
        }
        for (int e = tid; e < kSR * (Fp - F); e += kST)  // zero the K padding of Phi
            phi[(e / (Fp - F)) * Fp + F + e % (Fp - F)] = 0.f;
        if (w_fits) cp_async_wait_but_newest();          // z_l has landed; W_l may still be in flight
        else cp_async_wait_all();
        __syncthreads();
        K9_STAMP();
        for (int m = tid; m < M; m += kST) {
            float p[kSR];
#pragma unroll
            for (int r = 0; r < kSR; ++r) p[r] = 0.f;
            for (int q0 = 0; q0 < d; q0 += 20) {
                float zr[20];
#pragma unroll
                for (int i = 0; i < 20; ++i)
                    zr[i] = (q0 + i) < nzp ? zf_s[(q0 + i) * M + m] : ((q0 + i) < d ? __ldg(z + (int64_t)(q0 + i) * M + m) : 0.f);
#pragma unroll
                for (int i = 0; i < 20; ++i) {
                    const int q = q0 + i;
                    if (q < d) {
                        const float om = fmaf(s_s[q], zr[i], m_s[q]);
                        float in8[kSR];
                        ld8(q < y.d_prev ? fcur + q * kSR : x_t + (q - y.d_prev) * kSR, in8);
#pragma unroll
                        for (int r = 0; r < kSR; ++r) p[r] = fmaf(in8[r], om, p[r]);
                    }
                }
            }
            K9_STAMP2();
#pragma unroll
            for (int r = 0; r < kSR; ++r) {
                if (rbf) {
                    float s, c;
                    sincos_cw(p[r], &s, &c);
                    phi[r * Fp + m] = scale * c;
                    phi[r * Fp + M + m] = scale * s;
                } else {
                    phi[r * Fp + m] = scale * fmaxf(p[r], 0.f);
                }
            }
        }
        __syncthreads();
        K9_STAMP();
        bool z_next = false;
        if (a.prefetch_w && l + 1 < a.n_layers) {        // zf_s is free again: request the next layer's z rows
            const StepRowsLayer& y2 = a.layer[l + 1];
            const int nz2 = min(y2.d_prev + y2.d_x, (a.Fmax * kSR) / y2.M);
            stage_async(zf_s, y2.z + chain * y2.z_cs, nz2 * y2.M);
            cp_async_commit();
            z_next = true;
        }
        bool done = false;
        if (w_fits && y.g <= 16) {                       // whole W_l staged: register-blocked compile-time-N path
            if (z_next) cp_async_wait_but_newest(); else cp_async_wait_all();
            __syncthreads();
            done = phi_w_gemm_dispatch(y.g, phi, Fp, F, b_s, red, fnext);
        }
        if (!done) small_gemm_dispatch<true>(phi, Fp, Wl, 0, F, y.g, false, b_s, a.bs_cap, w_fits, z_next && w_fits, red, fnext);
        top_w_in_bs = w_fits;
        K9_STAMP();
        float* t = fcur; fcur = fnext; fnext = t;
    }

    // =========================== likelihood seed ===========================
    // fcur = F_{L-1}[j][r];  dF_{L-1} -> fnext
    if (tid < kSR) {
        const int r = tid;
        const int64_t row = row0 + r;
        float ll = 0.f;
        const bool live = row < a.B;
        if (a.likelihood == DGPRF_LIK_GAUSSIAN) {
            const float llv = __ldg(a.lik_log_var + chain * a.h_cs);
            const float inv_var = expf(-llv);
            for (int j = 0; j < a.d_out; ++j) {
                const float res = live ? __ldg(Y + row * a.d_out + j) - fcur[j * kSR + r] : 0.f;
                ll += live ? -0.5f * (DGPRF_LOG_2PI + llv + res * res * inv_var) : 0.f;
                fnext[j * kSR + r] = -(res * inv_var) * a.inv_B;
            }
        } else {
            float mx = -INFINITY;
            for (int j = 0; j < a.d_out; ++j) mx = fmaxf(mx, fcur[j * kSR + r]);
            float se = 0.f;
            for (int j = 0; j < a.d_out; ++j) se += expf(fcur[j * kSR + r] - mx);
            const float lse = mx + logf(se);
            const int label = live ? (int)__ldg(Y + row) : 0;
            for (int j = 0; j < a.d_out; ++j) {
                const float pj = expf(fcur[j * kSR + r] - lse);
                fnext[j * kSR + r] = live ? (pj - (j == label ? 1.f : 0.f)) * a.inv_B : 0.f;
            }
            ll = live ? ((label >= 0 && label < a.d_out) ? fcur[label * kSR + r] : NAN) - lse : 0.f;
        }
        red[r] = ll;
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kSR; ++r) s += red[r];
        a.ll_part[chain * a.ll_cs + grp] = s;
    }
    __syncthreads();
    float* dF = fnext;       // dU/dF_l (transposed)
    float* dFp = fcur;       // next (previous layer's) dF
    K9_STAMP();

    // =========================== backward ===========================
    bool w_prefetched = false;
    for (int l = a.n_layers - 1; l >= 0; --l) {
        const StepRowsLayer& y = a.layer[l];
        const int M = y.M, g = y.g;
        const bool rbf = y.kind == DGPRF_KIND_RBF;
        const int F = rbf ? 2 * M : M, Fp = padded_F(F), Fdp = a.Fmax;
        const float* W = y.W + chain * a.w_cs;
        const float* phi = phi_all + y.phi_off;
        float* gw = a.gwpart + chain * a.gw_cs + (int64_t)grp * a.gw_ss + y.off_W;
        // (a) per feature: dPhi = dF W^T, gW = Phi^T dF.  W rows are staged K-contiguous through shared
        //     memory; the gW rows of a feature batch go through a shared tile so the slab is written with
        //     coalesced stores (a thread-per-feature store pattern has a 4*g-byte stride between lanes).
        const int fc_max = min((F + 3) & ~3, ((a.bs_cap / g) - 4) & ~3);
        const int fb_max = min(kST, (kSW * (kSN + 1) * kSR) / g);      // features per store batch (tile = `red`)
        // W_l may already be staged: the top layer's copy from the forward, or the prefetch issued during the T GEMM
        // of the layer above.  The z rows of THIS layer's T GEMM are requested now and land during phase (a).
        const bool top = l == a.n_layers - 1;
        const bool w_ready = l > 0 && fc_max >= F && (top ? top_w_in_bs : w_prefetched);
        bool z_ready = false;
        if (a.prefetch_bwd && l > 0 && a.zr_cap > 0) {
            z_ready = small_gemm_prefetch_z(y.z + chain * y.z_cs, M, M, y.d_prev, zr_s, a.zr_cap);
            if (z_ready) cp_async_commit();
        }
        for (int f0 = 0; f0 < F; f0 += fc_max) {
            const int fc = min(fc_max, F - f0);
            if (l > 0) {                                  // layer 0 needs no dPhi, hence no W
                if (w_ready) {
                    if (!top) { if (z_ready) cp_async_wait_but_newest(); else cp_async_wait_all(); }
                } else {
                    stage_W_rows(b_s, W, f0, fc, fc, g);
                    cp_async_commit();
                    cp_async_wait_all();
                }
            }
            __syncthreads();
            for (int b0 = 0; b0 < fc; b0 += fb_max) {
                const int nb = min(fb_max, fc - b0);
                const int ff = b0 + tid;
                if (tid < nb) {
                    const int f = f0 + ff;
                    float ph[kSR], dph[kSR];
#pragma unroll
                    for (int r = 0; r < kSR; ++r) { ph[r] = phi[r * Fp + f]; dph[r] = 0.f; }
#pragma unroll 3
                    for (int j = 0; j < g; ++j) {
                        const float w = l > 0 ? b_s[ff * g + j] : 0.f;
                        float df[kSR];
                        ld8(dF + j * kSR, df);
                        float gwj = 0.f;
#pragma unroll
                        for (int r = 0; r < kSR; ++r) {
                            dph[r] = fmaf(df[r], w, dph[r]);
                            gwj = fmaf(ph[r], df[r], gwj);
                        }
                        red[tid * g + j] = gwj;
                    }
                    if (l > 0) {
#pragma unroll
                        for (int r = 0; r < kSR; ++r) dphi[r * Fdp + f] = dph[r];
                    }
                }
                __syncthreads();
                float* dst = gw + (int64_t)(f0 + b0) * g;                 // nb*g contiguous floats of the slab
                for (int e = tid; e < nb * g; e += kST) dst[e] = red[e];
                __syncthreads();
            }
        }
        K9_STAMP();
        if (l == 0) break;                               // nothing upstream of the first layer in W-only mode
        // (b) per column: dP, in place over the first M columns of dphi (the T GEMM zero-fills B beyond M)
        const float arc_scale = 1.41421356237f * __expf(__ldg(y.log_amp + chain * a.h_cs)) * rsqrtf((float)M);
        for (int m = tid; m < M; m += kST) {
#pragma unroll
            for (int r = 0; r < kSR; ++r) {
                const float pc = phi[r * Fp + m], dc = dphi[r * Fdp + m];
                float o;
                if (rbf) o = pc * dphi[r * Fdp + M + m] - phi[r * Fp + M + m] * dc;
                else o = pc > 0.f ? dc * arc_scale : 0.f;
                dphi[r * Fdp + m] = o;
            }
        }
        const float* s_s = s_all + l * a.dmax;           // s, mean of THIS layer: chain rule into F_{l-1}
        const float* m_s = m_all + l * a.dmax;
        __syncthreads();
        K9_STAMP();
        // (c) T = dP z^T (first d_prev rows of z), R = rowsum(dP)
        w_prefetched = false;
        if (a.prefetch_w && z_ready && l - 1 > 0) {      // b_s is idle during this T GEMM: request W_{l-1} for the next phase (a)
            const StepRowsLayer& y1 = a.layer[l - 1];
            const int F1 = y1.kind == DGPRF_KIND_RBF ? 2 * y1.M : y1.M;
            if (min((F1 + 3) & ~3, ((a.bs_cap / y1.g) - 4) & ~3) >= F1) {
                stage_W_rows(b_s, y1.W + chain * a.w_cs, 0, F1, F1, y1.g);
                cp_async_commit();
                w_prefetched = true;
            }
        }
        small_gemm_dispatch<false>(dphi, Fdp, y.z + chain * y.z_cs, M, M, y.d_prev, y.has_mean != 0, z_ready ? zr_s : b_s,
                                   z_ready ? a.zr_cap : a.bs_cap, z_ready, z_ready && w_prefetched, red, dFp);
        for (int e = tid; e < y.d_prev * kSR; e += kST) {
            const int q = e / kSR, r = e % kSR;
            float v = s_s[q] * dFp[e];
            if (y.has_mean) v = fmaf(m_s[q], dFp[y.d_prev * kSR + r], v);
            dFp[e] = v;
        }
        __syncthreads();
        K9_STAMP();
        float* t = dF; dF = dFp; dFp = t;
    }
    K9_STAMP();
    // =========================== fused update (cooperative launch only) ===========================
    if (a.fuse_update) {
        grid_barrier(a.bar, my_gen, gridDim.x * gridDim.y);
        K9_STAMP();
        const int64_t n4 = a.upd.n >> 2;
        const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;          // 128-bit vectors per CTA
        const int64_t v0 = (int64_t)grp * per, v1 = min(n4, v0 + per);
        const float* grad = a.upd.grad + chain * a.upd.grad_cs;
        const int sub = tid & 7;
        for (int64_t vb = v0; vb < v1; vb += kST / 8) {                 // 8 lanes per vector, 64 vectors per pass
            const int64_t v = vb + (tid >> 3);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f), th = g, m = g;
            if (sub == 0 && v < v1) {
                th = *reinterpret_cast<const float4*>(a.upd.theta + chain * a.upd.cs + (v << 2));
                m = *reinterpret_cast<const float4*>(a.upd.mom + chain * a.upd.cs + (v << 2));
            }
            if (v < v1) g = slab_sum_lane<8>(grad, a.upd.part_stride, a.upd.n_part, sub, v << 2);
            g = shuffle_sum_lpv<8>(g);
            if (sub == 0 && v < v1) sgmcmc_update_vec(a.upd, tab, chain, v, g, th, m);
        }
        if (a.u_out != nullptr && grp == 0 && tid < 32) {               // minibatch log-likelihood, fixed order
            float s = 0.f;
            for (int i = tid; i < (int)gridDim.x; i += 32) s += __ldcg(a.ll_part + chain * a.ll_cs + i);
            s = warp_sum(s);
            if (tid == 0) a.u_out[chain] = s;
        }
        K9_STAMP();
    }
#undef K9_STAMP
}

// ---- host side ---------------------------------------------------------------------------------------
static int host_padded_F(int F) { return ((F + 3) & ~3) + 4; }

// staging capacity: the largest W_l / z-row block, capped at 48 KB (bigger operands go through in chunks)
static int step_rows_bs_cap(const dgprf_model* m) {
    int64_t need = 256;
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        const int64_t F = host_padded_F(y.kind == DGPRF_KIND_RBF ? 2 * y.M : y.M);      // rows carry a +4 pad
        if (F * y.g > need) need = F * y.g;
        if ((int64_t)y.d_prev * (y.M + 8) > need) need = (int64_t)y.d_prev * (y.M + 8);
    }
    if (need > 12288) need = 12288;
    if (need < 64 * 12) need = 64 * 12;              // at least 8 K-columns for the widest (N = 64) operand
    return (int)round_up(need, 4);
}

// z rows of the widest backward T GEMM: d_prev x (M rounded to 4, + 4) floats (small_gemm's single-chunk layout)
static int step_rows_zr_need(const dgprf_model* m) {
    int64_t need = 0;
    for (int l = 1; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        const int64_t n = (int64_t)y.d_prev * (((y.M + 3) & ~3) + 4);
        if (n > need) need = n;
    }
    return (int)round_up(need, 4);
}
static size_t step_rows_smem_base(const dgprf_model* m);
// the prefetch buffer is optional: it is dropped when it would push the CTA past the 227 KB limit
static int step_rows_zr_cap(const dgprf_model* m) {
    const int need = step_rows_zr_need(m);
    return step_rows_smem_base(m) + sizeof(float) * (size_t)need <= 232448 ? need : 0;
}
size_t dgprf_step_rows_smem(const dgprf_model* m) { return step_rows_smem_base(m) + sizeof(float) * (size_t)step_rows_zr_cap(m); }

static size_t step_rows_smem_base(const dgprf_model* m) {
    int64_t dmax = 1, Fmax = 1, phis = 0;
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        const int d = y.d_prev + y.d_x, F = host_padded_F(y.kind == DGPRF_KIND_RBF ? 2 * y.M : y.M);
        if (d > dmax) dmax = d;
        if (F > Fmax) Fmax = F;
        phis += F;
    }
    dmax = round_up(dmax, 4);
    return sizeof(float) * (size_t)(m->d_in * kSR + 2 * (kSN + 1) * kSR + 2 * dmax * m->n_layers + kSW * (kSN + 1) * kSR + Fmax * kSR +
                                    step_rows_bs_cap(m) + phis * kSR);      // Fmax, phis: padded strides
}

int dgprf_step_rows_groups(int B) { return ceil_div(B, kSR); }

// The row-fused kernel applies when the features of R rows of every layer fit in shared memory and the
// minibatch is small enough that one CTA per 8 rows is a sensible grid.
bool dgprf_step_rows_eligible(const dgprf_model* m, int B) {
    if (m->precision != DGPRF_PREC_FP32) return false;
    if ((int64_t)ceil_div(B, kSR) * m->n_chains > 160) return false;      // more than ~one wave: the layered kernels win
    for (int l = 0; l < m->n_layers; ++l)
        if (m->layer[l].g > kSN || m->layer[l].d_prev > kSN) return false;
    return step_rows_smem_base(m) <= 220 * 1024;
}

// upd != nullptr asks for the fused update; *fused tells the caller whether it was done (it needs a
// cooperative launch, i.e. every CTA co-resident) or whether K5 still has to run.
int dgprf_launch_step_rows(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y, int64_t y_cs, int B,
                           float* gwpart, int64_t gw_cs, int64_t gw_ss, float* ll_part, int64_t ll_cs,
                           const UpdArgs* upd, const dgprf_segment* segs, int n_seg, unsigned int* bar, float* u_out,
                           bool* fused, cudaStream_t st) {
    StepRowsArgs a;
    memset(&a, 0, sizeof(a));
    a.n_layers = m->n_layers; a.likelihood = m->likelihood; a.B = B; a.d_in = m->d_in; a.d_out = m->d_out;
    a.h_cs = m->h_cs; a.w_cs = m->w_cs;
    a.X = X; a.x_cs = x_cs; a.Y = Y; a.y_cs = y_cs;
    a.lik_log_var = m->likelihood == DGPRF_LIK_GAUSSIAN ? m->h_base + m->off_lik_log_var : nullptr;
    a.gwpart = gwpart; a.gw_cs = gw_cs; a.gw_ss = gw_ss; a.ll_part = ll_part; a.ll_cs = ll_cs;
    a.inv_B = 1.f / (float)B;
    int64_t dmax = 1, Fmax = 1, phis = 0;
    for (int l = 0; l < m->n_layers; ++l) {
        const dgprf_layer& y = m->layer[l];
        StepRowsLayer& s = a.layer[l];
        const int F = host_padded_F(y.kind == DGPRF_KIND_RBF ? 2 * y.M : y.M);
        s.kind = y.kind; s.d_prev = y.d_prev; s.d_x = y.d_x; s.M = y.M; s.g = y.g; s.has_mean = y.has_mean;
        s.z = y.z; s.z_cs = y.z_cs;
        s.log_inv_ls = m->h_base + y.off_log_inv_ls; s.log_amp = m->h_base + y.off_log_amp;
        s.mean = y.has_mean ? m->h_base + y.off_mean : nullptr;
        s.W = m->w_base + y.off_W; s.off_W = y.off_W;
        s.phi_off = (int32_t)(phis * kSR);
        phis += F;
        if (y.d_prev + y.d_x > dmax) dmax = y.d_prev + y.d_x;
        if (F > Fmax) Fmax = F;
    }
    a.dmax = (int32_t)round_up(dmax, 4); a.Fmax = (int32_t)Fmax; a.bs_cap = step_rows_bs_cap(m); a.zr_cap = step_rows_zr_cap(m);
    const size_t smem = dgprf_step_rows_smem(m);
    { const int rc_s = dgprf_ensure_smem((const void*)k9_step_rows, smem); if (rc_s) return rc_s; }
    dim3 grid(ceil_div(B, kSR), m->n_chains);
    SegTable tab;
    memset(&tab, 0, sizeof(tab));
    *fused = false;
    if (upd != nullptr && !getenv("DGPRF_NO_FUSED_UPDATE")) {
        // co-resident CTAs for THIS shared-memory size on THIS device (a process-wide cache computed for the first
        // model would let a later, larger model through to a cooperative launch that cannot fit)
        static int cache_dev = -1, cache_val = 0;
        static size_t cache_smem = 0;
        int dev = 0;
        DGPRF_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev != cache_dev || smem != cache_smem) {
            int sms = 0, per_sm = 0;
            DGPRF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            DGPRF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k9_step_rows, kST, smem));
            cache_dev = dev; cache_smem = smem; cache_val = sms * per_sm;
        }
        const int max_coresident = cache_val;
        if ((int64_t)grid.x * grid.y <= max_coresident) {
            const int rc = dgprf_build_segtable(segs, n_seg, upd->n, &tab);
            if (rc) return rc;
            a.fuse_update = 1; a.bar = bar; a.u_out = u_out; a.upd = *upd;
            *fused = true;
        }
    }
    static long long* dbg = nullptr;                     // DGPRF_K9_TIMING=1: print phase cycle counts (debug only)
    static int dbg_calls = 0;
    if (getenv("DGPRF_K9_TIMING") && !dbg) cudaMalloc(&dbg, 64 * sizeof(long long));
    a.timing = dbg;
    a.prefetch_w = getenv("DGPRF_K9_NOPREFETCH") ? 0 : 1;
    a.prefetch_bwd = (a.prefetch_w && getenv("DGPRF_K9_BWD_PREFETCH")) ? 1 : 0;
    {
        ProfScope _ps("k9_step_rows", st);
        if (a.fuse_update) {
            void* kargs[2] = {&a, &tab};
            if (cudaLaunchCooperativeKernel((const void*)k9_step_rows, grid, dim3(kST), kargs, smem, st) != cudaSuccess) {
                (void)cudaGetLastError();             // refused (e.g. the GPU is shared): run unfused, K5 follows
                a.fuse_update = 0;
                *fused = false;
                k9_step_rows<<<grid, kST, smem, st>>>(a, tab);
            }
        } else {
            k9_step_rows<<<grid, kST, smem, st>>>(a, tab);
        }
    }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    if (dbg && ++dbg_calls == 30) {
        long long h[64];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const int n = 2 + 3 * m->n_layers + 3 * (m->n_layers - 1) + 2 + (a.fuse_update ? 1 : 0);
        fprintf(stderr, "k9 qloop-done offsets:");
        for (int l = 0; l < m->n_layers; ++l) fprintf(stderr, " %lld", h[40 + l] - h[1 + 3 * l]);
        fprintf(stderr, "\nk9 phase cycles:");
        for (int i = 1; i < n; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
        fprintf(stderr, "  total %lld\n", h[n - 1] - h[0]);
    }
    return DGPRF_OK;
}
