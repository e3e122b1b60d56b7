// Small kernels around the hot path: slab sums, gradient finalisation, hyper-parameter
// gradient reduction (K8 tail), log-prior (K4), Adam, Welford/mass (K6), stand-alone GP matmul.
#include "common.cuh"
#include "philox.cuh"
#include "kernels.cuh"

// ---- dense = sum of slabs ------------------------------------------------------------------
__global__ void k_sum_slabs(SlabMat m, int B, int ncol, float* out, int64_t out_cs) {
    dgprf_pdl_sync();
    const int chain = blockIdx.y;
    const int64_t n = (int64_t)B * ncol;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        out[chain * out_cs + e] = slab_load(m, chain, e / ncol, (int)(e % ncol));
}

// dense rows (ld == ncol): a slab is one contiguous array, so no index arithmetic and 128-bit accesses when aligned;
// slabs are added in slab order (the order of slab_load), four to eight loads in flight per thread
template <typename T>
__global__ void __launch_bounds__(256) k_sum_slabs_dense(const float* __restrict__ base, int64_t cs, int64_t ss, int n_slabs, int64_t n,
                                                         float* __restrict__ out, int64_t out_cs) {
    dgprf_pdl_sync();
    const T* p = reinterpret_cast<const T*>(base + blockIdx.y * cs);
    T* o = reinterpret_cast<T*>(out + blockIdx.y * out_cs);
    const int64_t sst = ss / (int64_t)(sizeof(T) / sizeof(float));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        T acc = __ldg(p + e);
        int sl = 1;
        for (; sl + 7 < n_slabs; sl += 8) {            // many slabs (gW: one per row split): eight loads in flight, adds in slab order
            T t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldg(p + (sl + u) * sst + e);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if constexpr (sizeof(T) == 16) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
                else acc += t[u];
            }
        }
        for (; sl + 3 < n_slabs; sl += 4) {
            const T t0 = __ldg(p + sl * sst + e), t1 = __ldg(p + (sl + 1) * sst + e), t2 = __ldg(p + (sl + 2) * sst + e),
                    t3 = __ldg(p + (sl + 3) * sst + e);
            if constexpr (sizeof(T) == 16) {
                acc.x += t0.x; acc.y += t0.y; acc.z += t0.z; acc.w += t0.w;
                acc.x += t1.x; acc.y += t1.y; acc.z += t1.z; acc.w += t1.w;
                acc.x += t2.x; acc.y += t2.y; acc.z += t2.z; acc.w += t2.w;
                acc.x += t3.x; acc.y += t3.y; acc.z += t3.z; acc.w += t3.w;
            } else {
                acc += t0; acc += t1; acc += t2; acc += t3;
            }
        }
        for (; sl < n_slabs; ++sl) {
            const T t = __ldg(p + sl * sst + e);
            if constexpr (sizeof(T) == 16) { acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
            else acc += t;
        }
        o[e] = acc;
    }
}

int dgprf_launch_sum_slabs(const SlabMat& m, int B, int ncol, float* out, int64_t out_cs, int n_chains, cudaStream_t st) {
    const int64_t n = (int64_t)B * ncol;
    ProfScope _ps("k_sum_slabs", st);
    if (m.ld == ncol) {
        const bool v4 = (n % 4) == 0 && (m.ss % 4) == 0 && (m.cs % 4) == 0 && (out_cs % 4) == 0 &&
                        ((reinterpret_cast<uintptr_t>(m.ptr) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
        const int64_t nv = v4 ? n / 4 : n;
        int blocks = (int)((nv + 255) / 256);
        if (blocks > 2368) blocks = 2368;
        if (v4) DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_sum_slabs_dense<float4>, dim3(blocks, n_chains), dim3(256), 0, st, m.ptr, m.cs, m.ss, m.n_slabs, nv, out, out_cs));
        else DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_sum_slabs_dense<float>, dim3(blocks, n_chains), dim3(256), 0, st, m.ptr, m.cs, m.ss, m.n_slabs, nv, out, out_cs));
    } else {
        int blocks = ceil_div(n, 256);
        if (blocks > 1184) blocks = 1184;
        DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_sum_slabs, dim3(blocks, n_chains), dim3(256), 0, st, m, B, ncol, out, out_cs));
    }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- out[c] = sum_i in[c][i] (fixed order; one CTA per chain) ---------------------------------
__global__ void __launch_bounds__(256) k_sum_rows(const float* in, int64_t in_cs, int n, float* out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += in[blockIdx.x * in_cs + i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

int dgprf_launch_sum_rows(const float* in, int64_t in_cs, int n, float* out, int n_chains, cudaStream_t st) {
    k_sum_rows<<<n_chains, 256, 0, st>>>(in, in_cs, n, out);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- gradient finalisation: out = sum_p part_p (+ theta * inv_N) ---------------------------
__global__ void k_grad_finalize(const float* part, int64_t part_cs, int64_t part_ss, int n_part,
                                const float* theta, int64_t theta_cs, float inv_N,
                                float* out, int64_t out_cs, int64_t n) {
    dgprf_pdl_sync();
    const int chain = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.f;
        for (int p = 0; p < n_part; ++p) g += __ldg(part + chain * part_cs + p * part_ss + i);
        if (inv_N != 0.f) g = fmaf(__ldg(theta + chain * theta_cs + i), inv_N, g);
        out[chain * out_cs + i] = g;
    }
}

int dgprf_launch_grad_finalize(const float* part, int64_t part_cs, int64_t part_ss, int n_part,
                               const float* theta, int64_t theta_cs, float inv_N,
                               float* out, int64_t out_cs, int64_t n, int n_chains, cudaStream_t st) {
    int blocks = ceil_div(n, 256);
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    { ProfScope _ps("k_grad_finalize", st);
      DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_grad_finalize, dim3(blocks, n_chains), dim3(256), 0, st, part, part_cs, part_ss, n_part, theta, theta_cs, inv_N, out, out_cs, n)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- K8 tail: hyper-parameter gradients of one layer from T, R, dF, F ----------------------
//   d log_inv_ls[q] = exp(log_inv_ls[q]) * sum_i in[i,q] * T[i,q]
//   d mean[q]       = sum_i in[i,q] * R[i]
//   d log_amp       = sum_ij dF[i,j] * F[i,j]
// Two launches so that large minibatches use the whole GPU and the result stays deterministic: row block rb of
// `rows_per` rows writes part[rb] = [ls partial (d) | mean partial (d) | log_amp partial], then one CTA per chain adds the
// row blocks in order.  The partial slabs (T, R, F, dF) are summed slab by slab in slab order.
__device__ __forceinline__ float slab_sum_rt(const SlabMat& m, int chain, int64_t row, int col) {
    const float* p = m.ptr + chain * m.cs + row * m.ld + col;
    float v = __ldg(p);
    for (int s = 1; s < m.n_slabs; ++s) v += __ldg(p + s * m.ss);
    return v;
}
__global__ void __launch_bounds__(256) k_hyper_partial(const HypArgs a) {
    dgprf_pdl_sync();
    __shared__ float red_a[8][33];
    __shared__ float red_b[8][33];
    __shared__ float red[32];
    const int chain = blockIdx.y, rb = blockIdx.x;
    const int64_t r0 = (int64_t)rb * a.rows_per, r1 = min((int64_t)a.B, r0 + a.rows_per);
    float* part = a.part + chain * a.part_cs + (int64_t)rb * (2 * a.d + 1);
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const float* X = a.X + chain * a.x_cs;
    for (int q0 = 0; q0 < a.d; q0 += 32) {
        const int q = q0 + lane;
        float acc_ls = 0.f, acc_mu = 0.f;
        if (q < a.d) {
#pragma unroll 4
            for (int64_t row = r0 + rg; row < r1; row += 8) {
                const float in = q < a.d_prev ? slab_sum_rt(a.Fprev, chain, row, q) : __ldg(X + row * a.ldx + (q - a.d_prev));
                acc_ls = fmaf(in, slab_sum_rt(a.T, chain, row, q), acc_ls);
                if (a.has_mean) acc_mu = fmaf(in, slab_sum_rt(a.R, chain, row, 0), acc_mu);
            }
        }
        red_a[rg][lane] = acc_ls;
        red_b[rg][lane] = acc_mu;
        __syncthreads();
        if (rg == 0 && q < a.d) {
            float s = 0.f, m = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) { s += red_a[r][lane]; m += red_b[r][lane]; }
            part[q] = s;
            part[a.d + q] = m;
        }
        __syncthreads();
    }
    float acc = 0.f;
    const int64_t n = (r1 - r0) * a.g;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
        const int64_t row = r0 + e / a.g; const int j = (int)(e % a.g);
        acc += slab_sum_rt(a.dF, chain, row, j) * slab_sum_rt(a.Fcur, chain, row, j);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[2 * a.d] = acc;
}
__global__ void __launch_bounds__(256) k_hyper_final(const HypArgs a) {
    dgprf_pdl_sync();
    const int chain = blockIdx.x;
    const float* part = a.part + chain * a.part_cs;
    float* gH = a.gH + chain * a.gh_cs;
    const int w = 2 * a.d + 1;
    for (int i = threadIdx.x; i < w; i += blockDim.x) {
        float s = 0.f;
        for (int rb = 0; rb < a.n_rb; ++rb) s += part[(int64_t)rb * w + i];
        if (i < a.d) gH[a.off_log_inv_ls + i] = expf(__ldg(a.log_inv_ls + chain * a.h_cs + i)) * s;
        else if (i < 2 * a.d) { if (a.has_mean) gH[a.off_mean + (i - a.d)] = s; }
        else gH[a.off_log_amp] = s;
    }
}

// row blocks of the hyper reduction for a minibatch of B rows: <= 2 per SM, >= 16 rows each.  (128 rows per block left a
// 1000-row minibatch on 8 CTAs whose threads walked 16 rows x 8 slabs one dependent round trip after the other: 30-47 us
// per layer, 60 % of a full-Bayesian step at the configs[1] shape.)
int dgprf_hyper_row_blocks(int B) {
    int n = ceil_div(B, 16);
    return n < 1 ? 1 : (n > 296 ? 296 : n);
}

int dgprf_launch_hyper_reduce(const HypArgs& a0, int n_chains, cudaStream_t st) {
    HypArgs a = a0;
    a.n_rb = dgprf_hyper_row_blocks(a.B);
    a.rows_per = ceil_div(a.B, a.n_rb);
    a.n_rb = ceil_div(a.B, a.rows_per);
    {
        ProfScope _ps("k8_hyper_reduce", st);
        DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_hyper_partial, dim3(a.n_rb, n_chains), dim3(256), 0, st, a));
        DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_hyper_final, dim3(n_chains), dim3(256), 0, st, a));
    }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- K4: sum log N(x;0,1) -------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k4_log_prior(const float* x, int64_t cs, int64_t n, float* out) {
    __shared__ float red[32];
    const int chain = blockIdx.x;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = __ldg(x + chain * cs + i);
        acc = fmaf(v, v, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[chain] = -0.5f * ((float)n * DGPRF_LOG_2PI + acc);
}

extern "C" int dgprf_log_prior(const float* x, int64_t cs, int64_t n, int n_chains, float* out, void* stream) {
    DGPRF_REQUIRE(x && out && n >= 0 && n_chains >= 1, "log_prior: bad arguments");
    k4_log_prior<<<n_chains, 1024, 0, (cudaStream_t)stream>>>(x, cs, n, out);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- Adam (keras semantics: lr_t = lr sqrt(1-b2^t)/(1-b1^t); theta -= lr_t m/(sqrt(v)+eps)) --
__global__ void k_adam(float* theta, const float* grad, float* m, float* v, int64_t n,
                       float lr_t, float b1, float b2, float eps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i];
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi; v[i] = vi;
        theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

extern "C" int dgprf_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n,
                               float lr, float beta1, float beta2, float eps, int t, void* stream) {
    DGPRF_REQUIRE(theta && grad && m && v && n >= 0 && t >= 1, "adam_step: bad arguments");
    if (n == 0) return DGPRF_OK;
    const float lr_t = lr * sqrtf(1.f - powf(beta2, (float)t)) / (1.f - powf(beta1, (float)t));
    int blocks = ceil_div(n, 256); if (blocks > 1184) blocks = 1184;
    k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, lr_t, beta1, beta2, eps);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- K6: Welford update and per-tensor mass (models/dgp.py:259-288) -------------------------
__global__ void k_welford(const float* grad, float* mean, float* m2, int64_t n, float inv_k) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i];
        const float delta = g - mean[i];
        const float mu = mean[i] + delta * inv_k;
        mean[i] = mu;
        m2[i] += delta * (g - mu);
    }
}

extern "C" int dgprf_welford_update(const float* grad, float* mean, float* m2, int64_t n, int k, void* stream) {
    DGPRF_REQUIRE(grad && mean && m2 && n >= 0 && k >= 1, "welford_update: bad arguments");
    if (n == 0) return DGPRF_OK;
    int blocks = ceil_div(n, 256); if (blocks > 1184) blocks = 1184;
    k_welford<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad, mean, m2, n, 1.f / (float)k);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

struct MassTable { int64_t offset[DGPRF_MAX_SEGMENTS]; int64_t length[DGPRF_MAX_SEGMENTS]; };

__global__ void __launch_bounds__(1024) k_mass_estimate(const float* mean, const float* m2, const __grid_constant__ MassTable tab,
                                                        float denom, int centered, float* mass_out) {
    __shared__ float red[32];
    const int s = blockIdx.x;
    const int64_t off = tab.offset[s], len = tab.length[s];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        const float var = m2[off + i] / denom;
        acc += centered ? var : fmaf(mean[off + i], mean[off + i], var);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) mass_out[s] = sqrtf(acc / (float)len + 1.0e-7f);
}

extern "C" int dgprf_mass_estimate(const float* mean, const float* m2, const dgprf_segment* segs, int n_seg,
                                   int K, int centered, float* mass_out, void* stream) {
    DGPRF_REQUIRE(mean && m2 && segs && mass_out && n_seg >= 1 && n_seg <= DGPRF_MAX_SEGMENTS, "mass_estimate: bad arguments");
    DGPRF_REQUIRE(K >= (centered ? 2 : 1), "mass_estimate: K_batches too small");
    MassTable tab;
    for (int s = 0; s < DGPRF_MAX_SEGMENTS; ++s) {
        tab.offset[s] = s < n_seg ? segs[s].offset : 0;
        tab.length[s] = s < n_seg ? segs[s].length : 0;
    }
    k_mass_estimate<<<n_seg, 1024, 0, (cudaStream_t)stream>>>(mean, m2, tab, centered ? (float)(K - 1) : (float)K,
                                                              centered, mass_out);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- stand-alone GPLayer matmul: out[B,g] = Phi[B,F] @ W[F,g]; one warp per row -------------
template <int GP>
__global__ void __launch_bounds__(256) k_gp_matmul(const float* Phi, const float* W, int B, int F, int g, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    float acc[GP];
#pragma unroll
    for (int j = 0; j < GP; ++j) acc[j] = 0.f;
    for (int k = lane; k < F; k += 32) {
        const float a = __ldg(Phi + row * F + k);
#pragma unroll
        for (int j = 0; j < GP; ++j)
            if (j < g) acc[j] = fmaf(a, __ldg(W + (int64_t)k * g + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < GP; ++j) {
        const float v = warp_sum(acc[j]);
        if (lane == 0 && j < g) out[row * g + j] = v;
    }
}

extern "C" int dgprf_gp_matmul(const float* Phi, const float* W, int B, int F, int g, float* out, void* stream) {
    DGPRF_REQUIRE(Phi && W && out && B >= 0 && F >= 1 && g >= 1, "gp_matmul: bad arguments");
    if (B == 0) return DGPRF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = ceil_div(B, 8);
    switch (pad_g(g)) {
        case 4:  k_gp_matmul<4><<<blocks, 256, 0, st>>>(Phi, W, B, F, g, out); break;
        case 16: k_gp_matmul<16><<<blocks, 256, 0, st>>>(Phi, W, B, F, g, out); break;
        case 32: k_gp_matmul<32><<<blocks, 256, 0, st>>>(Phi, W, B, F, g, out); break;
        case 64: k_gp_matmul<64><<<blocks, 256, 0, st>>>(Phi, W, B, F, g, out); break;
        default: dgprf_set_error("n_gp=%d > 64 unsupported", g); return DGPRF_EINVAL;
    }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

// ---- debug/test hook: fill a buffer with the update kernel's Philox normals -----------------
__global__ void k_philox_fill(float* out, int64_t n4, uint64_t seed, uint64_t chain, uint64_t step, uint32_t stream_id) {
    for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x)
        *reinterpret_cast<float4*>(out + 4 * i4) = philox_normal4(seed, chain, (uint64_t)i4, step, stream_id);
}

extern "C" int dgprf_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t chain, uint64_t step,
                                   int stream_id, void* stream) {
    DGPRF_REQUIRE(out && n >= 0 && (n & 3) == 0, "philox_normal: n must be a multiple of 4");
    if (n == 0) return DGPRF_OK;
    int blocks = ceil_div(n >> 2, 256); if (blocks > 1184) blocks = 1184;
    k_philox_fill<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, n >> 2, seed, chain, step, (uint32_t)stream_id);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
