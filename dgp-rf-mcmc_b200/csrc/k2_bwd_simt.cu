// K2 (fp32 / SIMT variant): reverse pass of one [RF layer -> GP layer] pair, reusing the
// features saved by K1 (no trigonometry in the backward).  Replaces what tf.GradientTape
// produces for models/dgp.py:194-204.
//
//   dPhi = dF W^T                                  gW   = Phi^T dF
//   RBF: dP = Phi_c * dPhi_s - Phi_s * dPhi_c      ARC: dP = dPhi * scale * [Phi > 0]
//   T = dP z^T,  R = rowsum(dP)
//   dF_prev = (exp(log_inv_ls) * T + mean * R)[:, :d_prev]
//   hyper mode keeps T (all d columns) and R for the hyper-gradient reduction (k_misc.cu).
//
// A CTA owns every CS-th column tile and, inside it, every RS-th row tile: gW accumulates in
// registers over the rows and lands in row-split slab `rs`; dF_prev/T/R accumulate over the
// column tiles into column-split slab `cs` (read-modify-write by the same thread only).
#include "kernels.cuh"

template <int GP>
__global__ void __launch_bounds__(kThreads)
k2_bwd_simt(const BwdArgs a) {
    dgprf_pdl_sync();
    extern __shared__ __align__(16) float smem[];
    constexpr int LDT = kTN + 4;               // 68: float4-aligned rows
    constexpr int LDZ = kTN + 1;               // 65
    float* dF_s  = smem;                       // [kTM][GP]
    float* wcT_s = dF_s + kTM * GP;            // [GP][kTN]   W rows of the cos (or relu) block, transposed
    float* wsT_s = wcT_s + GP * kTN;           // [GP][kTN]   W rows of the sin block, transposed
    float* phc_s = wsT_s + GP * kTN;           // [kTM][LDT]
    float* phs_s = phc_s + kTM * LDT;          // [kTM][LDT]
    float* dP_s  = phs_s + kTM * LDT;          // [kTM][LDT]
    float* z_s   = dP_s + kTM * LDT;           // [64 q][LDZ]
    float* R_s   = z_s + kTN * LDZ;            // [kTM]
    float* s_all = R_s + kTM;                  // [d]
    float* m_all = s_all + a.d;                // [d]

    const int tid = threadIdx.x;
    const int chain = blockIdx.z, cs = blockIdx.y, rs = blockIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const bool rbf = a.kind == DGPRF_KIND_RBF;

    const float* z = a.z + chain * a.z_cs;
    const float* W = a.W + chain * a.w_cs;
    const float* Phi = a.Phi + chain * a.phi_cs;
    const float* ls = a.log_inv_ls + chain * a.h_cs;
    const float* mean = a.has_mean ? a.mean + chain * a.h_cs : nullptr;
    const float amp = __expf(__ldg(a.log_amp + chain * a.h_cs));
    const float arc_scale = 1.41421356237f * amp * rsqrtf((float)a.M);
    const bool need_R = a.has_mean || a.hyper;
    const int dq = a.hyper ? a.d : a.d_prev;   // T columns that somebody consumes
    const bool z_once = dq <= kTN;              // a single T pass: stage z once per column tile

    for (int q = tid; q < a.d; q += kThreads) {
        s_all[q] = expf(__ldg(ls + q));
        m_all[q] = mean ? __ldg(mean + q) : 0.f;
    }
    // 16-byte async copies of Phi need M % 4 == 0 (rows of Phi and the sin block stay 16 B aligned)
    const bool phi_vec = (a.M % 4 == 0) && ((a.phi_cs % 4) == 0);

    const int n_ct = (a.M + kTN - 1) / kTN, n_rt = (a.B + kTM - 1) / kTM;
    // gW phase mapping: feature f of the 128 (64 cos|relu + 64 sin), half jh of the outputs
    const int gf = tid & 127, jh = tid >> 7;
    constexpr int GH = GP / 2;

    for (int ct = cs; ct < n_ct; ct += a.CS) {
        const int c0 = ct * kTN;
        const bool first_ct = ct == cs;
        __syncthreads();
        for (int e = tid; e < GP * kTN; e += kThreads) {
            const int c = e / GP, k = e % GP;      // coalesced along a W row
            const int col = c0 + c;
            const bool ok = col < a.M && k < a.g;
            cp_async4(wcT_s + k * kTN + c, W + (int64_t)col * a.g + k, ok);
            cp_async4(wsT_s + k * kTN + c, W + (int64_t)(a.M + col) * a.g + k, ok && rbf);
        }
        if (z_once && dq > 0) {                    // the z tile depends on the column tile only
            for (int e = tid; e < kTN * kTN; e += kThreads) {
                const int qq = e / kTN, c = e % kTN;
                cp_async4(z_s + qq * LDZ + c, z + (int64_t)qq * a.M + c0 + c, qq < a.d && (c0 + c) < a.M);
            }
        }
        cp_async_commit();
        float accg[GH];
#pragma unroll
        for (int j = 0; j < GH; ++j) accg[j] = 0.f;

        for (int rt = rs; rt < n_rt; rt += a.RS) {
            const int row0 = rt * kTM;
            __syncthreads();
            if (phi_vec) {
                for (int e = tid; e < kTM * (kTN / 4); e += kThreads) {
                    const int r = e / (kTN / 4), c = (e % (kTN / 4)) * 4;
                    const int64_t row = row0 + r;
                    const bool ok = row < a.B && (c0 + c) < a.M;
                    cp_async16(phc_s + r * LDT + c, Phi + row * a.F + c0 + c, ok);
                    cp_async16(phs_s + r * LDT + c, Phi + row * a.F + a.M + c0 + c, ok && rbf);
                }
            } else {
                for (int e = tid; e < kTM * kTN; e += kThreads) {
                    const int r = e / kTN, c = e % kTN;
                    const int64_t row = row0 + r;
                    const bool ok = row < a.B && (c0 + c) < a.M;
                    cp_async4(phc_s + r * LDT + c, Phi + row * a.F + c0 + c, ok);
                    cp_async4(phs_s + r * LDT + c, Phi + row * a.F + a.M + c0 + c, ok && rbf);
                }
            }
            cp_async_commit();
            {
                float v[kTM * GP / kThreads];
#pragma unroll
                for (int u = 0; u < kTM * GP / kThreads; ++u) {
                    const int e = tid + u * kThreads;
                    const int r = e / GP, k = e % GP;
                    const int64_t row = row0 + r;
                    v[u] = (row < a.B && k < a.g) ? slab_load(a.dF, chain, row, k) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < kTM * GP / kThreads; ++u) dF_s[tid + u * kThreads] = v[u];
            }
            cp_async_wait_all();
            __syncthreads();

            // ---- dPhi = dF W^T on a 4x4 micro-tile, then dP ----
            {
                float dc[4][4], ds[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { dc[i][j] = 0.f; ds[i][j] = 0.f; }
                for (int k = 0; k < a.g; ++k) {
                    const float4 wc = *reinterpret_cast<const float4*>(wcT_s + k * kTN + tx * 4);
                    const float4 ws = *reinterpret_cast<const float4*>(wsT_s + k * kTN + tx * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float df = dF_s[(ty * 4 + i) * GP + k];
                        dc[i][0] = fmaf(df, wc.x, dc[i][0]); dc[i][1] = fmaf(df, wc.y, dc[i][1]);
                        dc[i][2] = fmaf(df, wc.z, dc[i][2]); dc[i][3] = fmaf(df, wc.w, dc[i][3]);
                        if (rbf) {
                            ds[i][0] = fmaf(df, ws.x, ds[i][0]); ds[i][1] = fmaf(df, ws.y, ds[i][1]);
                            ds[i][2] = fmaf(df, ws.z, ds[i][2]); ds[i][3] = fmaf(df, ws.w, ds[i][3]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = ty * 4 + i;
                    const float4 pc = *reinterpret_cast<const float4*>(phc_s + r * LDT + tx * 4);
                    float4 o;
                    if (rbf) {
                        const float4 ps = *reinterpret_cast<const float4*>(phs_s + r * LDT + tx * 4);
                        o.x = pc.x * ds[i][0] - ps.x * dc[i][0];
                        o.y = pc.y * ds[i][1] - ps.y * dc[i][1];
                        o.z = pc.z * ds[i][2] - ps.z * dc[i][2];
                        o.w = pc.w * ds[i][3] - ps.w * dc[i][3];
                    } else {
                        o.x = pc.x > 0.f ? dc[i][0] * arc_scale : 0.f;
                        o.y = pc.y > 0.f ? dc[i][1] * arc_scale : 0.f;
                        o.z = pc.z > 0.f ? dc[i][2] * arc_scale : 0.f;
                        o.w = pc.w > 0.f ? dc[i][3] * arc_scale : 0.f;
                    }
                    *reinterpret_cast<float4*>(dP_s + r * LDT + tx * 4) = o;
                }
            }

            // ---- gW += Phi^T dF : thread = (feature gf, output half jh) ----
            if (rbf || gf < kTN) {
                const float* ph = gf < kTN ? phc_s + gf : phs_s + (gf - kTN);
#pragma unroll 4
                for (int i = 0; i < kTM; ++i) {
                    const float av = ph[i * LDT];
                    const float* dfr = dF_s + i * GP + jh * GH;
#pragma unroll
                    for (int j = 0; j < GH; ++j) accg[j] = fmaf(av, dfr[j], accg[j]);
                }
            }
            __syncthreads();   // dP_s complete

            if (need_R) {
                if (tid < kTM) {
                    float r = 0.f;
                    for (int c = 0; c < kTN; ++c) r += dP_s[tid * LDT + c];
                    R_s[tid] = r;
                    const int64_t row = row0 + tid;
                    if (a.hyper && row < a.B) {
                        float* dst = a.Rpart + chain * a.r_cs + (int64_t)cs * a.B + row;
                        *dst = first_ct ? r : *dst + r;
                    }
                }
                __syncthreads();
            }

            // ---- T = dP z^T in passes of 64 input columns ----
            for (int q0 = 0; q0 < dq; q0 += kTN) {
                if (!z_once) {
                    for (int e = tid; e < kTN * kTN; e += kThreads) {
                        const int qq = e / kTN, c = e % kTN;
                        const int q = q0 + qq, col = c0 + c;
                        z_s[qq * LDZ + c] = (q < a.d && col < a.M) ? __ldg(z + (int64_t)q * a.M + col) : 0.f;
                    }
                    __syncthreads();
                }
                float t[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) t[i][j] = 0.f;
#pragma unroll 4
                for (int c = 0; c < kTN; ++c) {
                    float dv[4], zv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) dv[i] = dP_s[(ty * 4 + i) * LDT + c];
#pragma unroll
                    for (int j = 0; j < 4; ++j) zv[j] = z_s[(tx + 16 * j) * LDZ + c];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[i][j] = fmaf(dv[i], zv[j], t[i][j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = ty * 4 + i;
                    const int64_t row = row0 + r;
                    if (row >= a.B) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int q = q0 + tx + 16 * j;
                        if (q < a.d_prev && a.Dpart != nullptr) {
                            float v = s_all[q] * t[i][j];
                            if (a.has_mean) v = fmaf(m_all[q], R_s[r], v);
                            float* dst = a.Dpart + chain * a.d_cs + ((int64_t)cs * a.B + row) * a.d_prev + q;
                            *dst = first_ct ? v : *dst + v;
                        }
                        if (a.hyper && q < a.d) {
                            float* dst = a.Tpart + chain * a.t_cs + ((int64_t)cs * a.B + row) * a.d + q;
                            *dst = first_ct ? t[i][j] : *dst + t[i][j];
                        }
                    }
                }
                __syncthreads();
            }
        }

        // ---- flush this column tile's gW rows into row-split slab rs ----
        if (rbf || gf < kTN) {
            const int col = c0 + (gf < kTN ? gf : gf - kTN);
            if (col < a.M) {
                const int64_t frow = (gf < kTN ? 0 : a.M) + col;
                float* dst = a.gWpart + chain * a.gw_cs + (int64_t)rs * a.gw_ss + frow * a.g;
#pragma unroll
                for (int j = 0; j < GH; ++j) {
                    const int jj = jh * GH + j;
                    if (jj < a.g) dst[jj] = accg[j];
                }
            }
        }
    }
}

static size_t bwd_smem_bytes(int GP, int d) {
    return sizeof(float) * (size_t)(kTM * GP + 2 * GP * kTN + 3 * kTM * (kTN + 4) + kTN * (kTN + 1) + kTM + 2 * d);
}

template <int GP>
static int launch_bwd(const BwdArgs& a, int n_chains, cudaStream_t st) {
    const size_t smem = bwd_smem_bytes(GP, a.d);
    { const int rc_s = dgprf_ensure_smem((const void*)k2_bwd_simt<GP>, (size_t)smem); if (rc_s) return rc_s; }
    dim3 grid(a.RS, a.CS, n_chains);
    { ProfScope _ps("k2_bwd_simt", st); k2_bwd_simt<GP><<<grid, kThreads, smem, st>>>(a); }      // (no programmatic launch: measured 12 % slower at configs[4] scale in fp32)
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

int dgprf_launch_bwd_simt(const BwdArgs& a, int n_chains, cudaStream_t st) {
    DGPRF_REQUIRE(a.d <= 8192, "RF layer input width %d > 8192 unsupported", a.d);
    switch (pad_g(a.g)) {
        case 4:  return launch_bwd<4>(a, n_chains, st);
        case 16: return launch_bwd<16>(a, n_chains, st);
        case 32: return launch_bwd<32>(a, n_chains, st);
        case 64: return launch_bwd<64>(a, n_chains, st);
        default: dgprf_set_error("n_gp=%d > 64 unsupported", a.g); return DGPRF_EINVAL;
    }
}
