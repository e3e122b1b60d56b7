// K1 (fp32 / SIMT variant): one fused [RF layer -> GP layer] forward.
//
//   in   = concat(sum_slabs Fprev[:, :d_prev], X[:, :d_x])              (utils.py:42)
//   P    = in @ (exp(log_inv_ls)[:,None] * z + mean)                    (layers/rf_layers.py:36-42, 82-88)
//   Phi  = amp/sqrt(M) [cos P, sin P]  |  sqrt(2) amp/sqrt(M) relu(P)   (layers/rf_layers.py:43-44, 89-90)
//   F    = Phi @ W                                                      (layers/GP_weight_layers.py:13)
//
// P never reaches HBM.  A CTA owns a 64-row tile and every CS-th 64-column tile of the random
// features; it writes Phi (optionally, for the backward) and ONE partial slab of F.  The CS
// slabs are summed, in fixed order, by whoever reads F next (deterministic, no atomics).
// Latency structure: the W tile of GEMM #2 is fetched with cp.async while GEMM #1 runs, and
// every global load of a phase is in flight before the first one is consumed.
#include "kernels.cuh"

template <int GP>
__global__ void __launch_bounds__(kThreads)
k1_fwd_simt(const FwdArgs a) {
    dgprf_pdl_sync();
    extern __shared__ __align__(16) float smem[];
    constexpr int LDI = kKC + 1;
    constexpr int LDP = 2 * kTN + 1;
    float* in_s  = smem;                       // [kTM][LDI]
    float* om_s  = in_s + kTM * LDI;           // [kKC][kTN]
    float* phi_s = om_s + kKC * kTN;           // [kTM][LDP]   (also the cross-quarter reduction buffer)
    float* w_s   = phi_s + kTM * LDP;          // [2*kTN][GP]

    const int tid = threadIdx.x;
    const int chain = blockIdx.z;
    const int cs = blockIdx.y;
    const int row0 = blockIdx.x * kTM;
    const int tx = tid & 15, ty = tid >> 4;

    const float* z = a.z + chain * a.z_cs;
    const float* X = a.X + chain * a.x_cs;
    const float* W = a.W + chain * a.w_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    const float amp = __expf(__ldg(a.log_amp + chain * a.h_cs));
    const bool rbf = a.kind == DGPRF_KIND_RBF;
    const float scale = (rbf ? 1.f : 1.41421356237f) * amp * rsqrtf((float)a.M);

    const int KT = rbf ? 2 * kTN : kTN;        // K extent of GEMM #2 per column tile
    const int r2 = tid & (kTM - 1), kq = tid >> 6;
    float acc2[GP];
#pragma unroll
    for (int j = 0; j < GP; ++j) acc2[j] = 0.f;

    const int n_ct = (a.M + kTN - 1) / kTN;
    for (int ct = cs; ct < n_ct; ct += a.CS) {
        const int c0 = ct * kTN;
        if (a.do_gemm2) {                       // W tile of this column tile: async, consumed after GEMM #1
            for (int e = tid; e < KT * GP; e += kThreads) {
                const int k = e / GP, j = e % GP;
                const int col = c0 + (k < kTN ? k : k - kTN);
                const int64_t frow = (k < kTN ? 0 : a.M) + col;
                cp_async4(w_s + e, W + frow * a.g + j, col < a.M && j < a.g);
            }
            cp_async_commit();
        }
        float p[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) p[i][j] = 0.f;

        for (int k0 = 0; k0 < a.d; k0 += kKC) {
            float vin[kTM * kKC / kThreads], vom[kKC * kTN / kThreads];
#pragma unroll
            for (int u = 0; u < kTM * kKC / kThreads; ++u) {
                const int e = tid + u * kThreads;
                const int r = e / kKC, k = e % kKC, q = k0 + k;
                const int64_t row = row0 + r;
                float v = 0.f;
                if (row < a.B && q < a.d)
                    v = q < a.d_prev ? slab_load(a.Fprev, chain, row, q)
                                     : __ldg(X + row * a.ldx + (q - a.d_prev));
                vin[u] = v;
            }
#pragma unroll
            for (int u = 0; u < kKC * kTN / kThreads; ++u) {
                const int e = tid + u * kThreads;
                const int k = e / kTN, c = e % kTN, q = k0 + k, col = c0 + c;
                float v = 0.f;
                if (q < a.d && col < a.M)
                    v = fmaf(expf(__ldg(ls + q)), __ldg(z + (int64_t)q * a.M + col), mean ? __ldg(mean + q) : 0.f);
                vom[u] = v;
            }
#pragma unroll
            for (int u = 0; u < kTM * kKC / kThreads; ++u) {
                const int e = tid + u * kThreads;
                in_s[(e / kKC) * LDI + (e % kKC)] = vin[u];
            }
#pragma unroll
            for (int u = 0; u < kKC * kTN / kThreads; ++u) om_s[tid + u * kThreads] = vom[u];
            __syncthreads();
            const int kmax = min(kKC, a.d - k0);
            for (int k = 0; k < kmax; ++k) {
                const float4 b = *reinterpret_cast<const float4*>(om_s + k * kTN + tx * 4);
                float av[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = in_s[(ty * 4 + i) * LDI + k];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p[i][0] = fmaf(av[i], b.x, p[i][0]);
                    p[i][1] = fmaf(av[i], b.y, p[i][1]);
                    p[i][2] = fmaf(av[i], b.z, p[i][2]);
                    p[i][3] = fmaf(av[i], b.w, p[i][3]);
                }
            }
            __syncthreads();
        }

        // ---- fused epilogue: activation, scale, save, stage for GEMM #2 ----
        const bool vec_ok = (a.M % 4 == 0) && (c0 + tx * 4 + 3 < a.M);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const int64_t row = row0 + r;
            float f0[4], f1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool live = (c0 + tx * 4 + j) < a.M;
                if (rbf) {
                    float s, c;
                    sincos_cw(p[i][j], &s, &c);
                    f0[j] = live ? scale * c : 0.f;
                    f1[j] = live ? scale * s : 0.f;
                } else {
                    f0[j] = live ? scale * fmaxf(p[i][j], 0.f) : 0.f;
                    f1[j] = 0.f;
                }
                phi_s[r * LDP + tx * 4 + j] = f0[j];
                if (rbf) phi_s[r * LDP + kTN + tx * 4 + j] = f1[j];
            }
            if (a.Phi != nullptr && row < a.B) {
                float* dst = a.Phi + chain * a.phi_cs + row * a.F + c0 + tx * 4;
                if (vec_ok) {
                    *reinterpret_cast<float4*>(dst) = make_float4(f0[0], f0[1], f0[2], f0[3]);
                    if (rbf) *reinterpret_cast<float4*>(dst + a.M) = make_float4(f1[0], f1[1], f1[2], f1[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + tx * 4 + j < a.M) {
                            dst[j] = f0[j];
                            if (rbf) dst[a.M + j] = f1[j];
                        }
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (a.do_gemm2) {
            const int kb = kq * (KT / 4);
            for (int kk = 0; kk < KT / 4; ++kk) {
                const float av = phi_s[r2 * LDP + kb + kk];
                const float* wr = w_s + (kb + kk) * GP;
#pragma unroll
                for (int j4 = 0; j4 < GP / 4; ++j4) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + j4 * 4);
                    acc2[j4 * 4 + 0] = fmaf(av, w.x, acc2[j4 * 4 + 0]);
                    acc2[j4 * 4 + 1] = fmaf(av, w.y, acc2[j4 * 4 + 1]);
                    acc2[j4 * 4 + 2] = fmaf(av, w.z, acc2[j4 * 4 + 2]);
                    acc2[j4 * 4 + 3] = fmaf(av, w.w, acc2[j4 * 4 + 3]);
                }
            }
        }
        __syncthreads();
    }

    if (a.do_gemm2) {
        // reduce the four K-quarters through shared memory (phi_s is free now), fixed order
        float* red = phi_s;   // [4][kTM][GP]  (GP<=32: 8192 floats <= kTM*LDP)
        constexpr int RED_ROUNDS = (4 * kTM * GP > kTM * LDP) ? 2 : 1;   // GP==64 needs two halves
        constexpr int GH = GP / RED_ROUNDS;
#pragma unroll
        for (int h = 0; h < RED_ROUNDS; ++h) {
#pragma unroll
            for (int j = 0; j < GH; ++j) red[(kq * kTM + r2) * GH + j] = acc2[h * GH + j];
            __syncthreads();
            for (int e = tid; e < kTM * GH; e += kThreads) {
                const int r = e / GH, j = e % GH, jj = h * GH + j;
                const int64_t row = row0 + r;
                if (row < a.B && jj < a.g) {
                    const float v = ((red[(0 * kTM + r) * GH + j] + red[(1 * kTM + r) * GH + j]) +
                                     red[(2 * kTM + r) * GH + j]) + red[(3 * kTM + r) * GH + j];
                    a.Fpart[chain * a.fpart_cs + ((int64_t)cs * a.B + row) * a.g + jj] = v;
                }
            }
            __syncthreads();
        }
    }
}

static size_t fwd_smem_bytes(int GP) {
    return sizeof(float) * (size_t)(kTM * (kKC + 1) + kKC * kTN + kTM * (2 * kTN + 1) + 2 * kTN * GP);
}

template <int GP>
static int launch_fwd(const FwdArgs& a, int n_chains, cudaStream_t st) {
    const size_t smem = fwd_smem_bytes(GP);
    { const int rc_s = dgprf_ensure_smem((const void*)k1_fwd_simt<GP>, (size_t)smem); if (rc_s) return rc_s; }
    dim3 grid(ceil_div(a.B, kTM), a.CS, n_chains);
    { ProfScope _ps("k1_fwd_simt", st); k1_fwd_simt<GP><<<grid, kThreads, smem, st>>>(a); }      // (no programmatic launch: measured 12 % slower at configs[4] scale in fp32)
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

int dgprf_launch_fwd_simt(const FwdArgs& a, int n_chains, cudaStream_t st) {
    switch (pad_g(a.do_gemm2 ? a.g : 1)) {
        case 4:  return launch_fwd<4>(a, n_chains, st);
        case 16: return launch_fwd<16>(a, n_chains, st);
        case 32: return launch_fwd<32>(a, n_chains, st);
        case 64: return launch_fwd<64>(a, n_chains, st);
        default: dgprf_set_error("n_gp=%d > 64 unsupported", a.g); return DGPRF_EINVAL;
    }
}
