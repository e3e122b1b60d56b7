// K7: posterior-predictive averaging over stored samples
// (experiments/utils_training.py:79-85, 160-166, 242-247, 323-328):
//   out[0] = mean_n( logsumexp_s log_p[s,n] - log S_total )
//   out[1] = sqrt(mean_{s,n} se[s,n])   |   mean_{s,n} acc
// Column-wise streaming: a thread owns a test point and walks the S samples with an online
// (max, sum exp) pair -- coalesced across the warp, 4*S*N bytes read once (x2 with aux).
// Two fixed-order stages (per-block partials, then one block) keep it deterministic.
#include "common.cuh"

// VEC columns per thread (128-bit loads when VEC = 4) and 8 samples' loads in flight before the (sequential, per-column)
// online updates: the recurrence on (max, sum) would otherwise serialise one L2 / HBM round trip per sample.
// 64 registers -> 4 blocks per SM -> ~128 KB of loads in flight per SM (8 x 16 B per thread without the aux matrix,
// 4 x 2 x 16 B with it), which is what a 6.5 TB/s stream needs at ~2 us of loaded latency.
template <int VEC, bool AUX>
__global__ void __launch_bounds__(256, 4)
k7_predictive_cols(const float* __restrict__ log_p, const float* __restrict__ aux, int S, int64_t N,
                   int64_t ld, float log_S, float* __restrict__ lse_cols, float* __restrict__ part) {
    __shared__ float red[32];
    constexpr int SB = AUX ? 4 : 8;
    const int64_t n0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    float lp = 0.f, ax = 0.f;
    if (n0 < N) {
        float mx[VEC], sum[VEC], axv[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) { mx[c] = -INFINITY; sum[c] = 0.f; axv[c] = 0.f; }
        for (int s0 = 0; s0 < S; s0 += SB) {
            float v[SB][VEC], w[AUX ? SB : 1][VEC];
#pragma unroll
            for (int u = 0; u < SB; ++u) {
                const bool ok = s0 + u < S;
                if (VEC == 4) {
                    const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(log_p + (int64_t)(s0 + u) * ld + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w;
                    if (AUX) {
                        const float4 q = ok ? __ldg(reinterpret_cast<const float4*>(aux + (int64_t)(s0 + u) * ld + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        w[u][0] = q.x; w[u][1 % VEC] = q.y; w[u][2 % VEC] = q.z; w[u][3 % VEC] = q.w;
                    }
                } else {
                    v[u][0] = ok ? __ldg(log_p + (int64_t)(s0 + u) * ld + n0) : 0.f;
                    if (AUX) w[u][0] = ok ? __ldg(aux + (int64_t)(s0 + u) * ld + n0) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < SB; ++u) {
                if (s0 + u < S) {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) {
                        const float x = v[u][c];
                        if (x > mx[c]) { sum[c] = sum[c] * __expf(mx[c] - x) + 1.f; mx[c] = x; }   // exp(-inf)=0 on the first sample
                        else sum[c] += __expf(x - mx[c]);
                        if (AUX) axv[c] += w[AUX ? u : 0][c];
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            if (n0 + c < N) {
                const float lse = mx[c] + logf(sum[c]);
                if (lse_cols) lse_cols[n0 + c] = lse;
                lp += lse - log_S;
                ax += axv[c];
            }
        }
    }
    const float a = block_sum(lp, red);
    const float b = block_sum(ax, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
}

__global__ void __launch_bounds__(256)
k7_predictive_final(const float* __restrict__ part, int n_part, int S, int64_t N, int aux_is_se,
                    int has_aux, float* __restrict__ out) {
    __shared__ float red[32];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) { a += part[2 * i]; b += part[2 * i + 1]; }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
        out[0] = a / (float)N;
        const float m = b / ((float)N * (float)S);
        out[1] = has_aux ? (aux_is_se ? sqrtf(m) : m) : 0.f;
    }
}

extern "C" int dgprf_predictive_reduce(const float* log_p, const float* aux, int S, int64_t N, int64_t ld,
                                       float log_S_total, int aux_is_se, float* lse_cols, float* out,
                                       float* scratch, void* stream) {
    DGPRF_REQUIRE(log_p && out && scratch && S >= 1 && N >= 1 && ld >= N, "predictive_reduce: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (N % 4) == 0 && (ld % 4) == 0 && ((reinterpret_cast<uintptr_t>(log_p) | reinterpret_cast<uintptr_t>(aux)) & 15) == 0;
    const int nb = vec4 ? ceil_div(N, 1024) : ceil_div(N, 256);
    if (vec4 && aux) k7_predictive_cols<4, true><<<nb, 256, 0, st>>>(log_p, aux, S, N, ld, log_S_total, lse_cols, scratch);
    else if (vec4) k7_predictive_cols<4, false><<<nb, 256, 0, st>>>(log_p, aux, S, N, ld, log_S_total, lse_cols, scratch);
    else if (aux) k7_predictive_cols<1, true><<<nb, 256, 0, st>>>(log_p, aux, S, N, ld, log_S_total, lse_cols, scratch);
    else k7_predictive_cols<1, false><<<nb, 256, 0, st>>>(log_p, aux, S, N, ld, log_S_total, lse_cols, scratch);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    k7_predictive_final<<<1, 256, 0, st>>>(scratch, nb, S, N, aux_is_se, aux != nullptr, out);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
