// K7: posterior-predictive averaging over stored samples
// (experiments/utils_training.py:79-85, 160-166, 242-247, 323-328):
//   out[0] = mean_n( logsumexp_s log_p[s,n] - log S_total )
//   out[1] = sqrt(mean_{s,n} se[s,n])   |   mean_{s,n} acc
// Column-wise streaming: a thread owns a test point and walks the S samples with an online
// (max, sum exp) pair -- coalesced across the warp, 4*S*N bytes read once (x2 with aux).
// Two fixed-order stages (per-block partials, then one block) keep it deterministic.
#include "common.cuh"

__global__ void __launch_bounds__(256)
k7_predictive_cols(const float* __restrict__ log_p, const float* __restrict__ aux, int S, int64_t N,
                   int64_t ld, float log_S, float* __restrict__ lse_cols, float* __restrict__ part) {
    __shared__ float red[32];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float lp = 0.f, ax = 0.f;
    if (n < N) {
        float mx = -INFINITY, sum = 0.f;
        for (int s = 0; s < S; ++s) {
            const float v = __ldg(log_p + s * ld + n);
            if (v > mx) { sum = sum * expf(mx - v) + 1.f; mx = v; }   // exp(-inf)=0 on the first sample
            else sum += expf(v - mx);
            if (aux) ax += __ldg(aux + s * ld + n);
        }
        const float lse = mx + logf(sum);
        if (lse_cols) lse_cols[n] = lse;
        lp = lse - log_S;
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
    const int nb = ceil_div(N, 256);
    k7_predictive_cols<<<nb, 256, 0, st>>>(log_p, aux, S, N, ld, log_S_total, lse_cols, scratch);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    k7_predictive_final<<<1, 256, 0, st>>>(scratch, nb, S, N, aux_is_se, aux != nullptr, out);
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
