// K3: likelihood reductions (warp-shuffle), emitting the backward seed dU/dF directly.
//   Gaussian  likelihoods/gaussian.py:18-25 + utils.py:46-47
//   Softmax   likelihoods/softmax.py:8-22 (labels = int(Y[:,0]))
// plus what the evaluation methods derive from the same pass
//   models/regression_model.py:46   se = mean_D (y-f)^2
//   models/classification_model.py:21-24   argmax(softmax(F)) == label
// Small batches: one CTA per chain.  Large batches: up to 64 CTAs per chain write per-CTA partial sums that a
// second tiny kernel adds in a fixed order -> deterministic either way.
#include "kernels.cuh"

constexpr int kLikThreads = 1024;
constexpr int kMaxClasses = 64;

__global__ void __launch_bounds__(kLikThreads) k3_loglik(const LikArgs a) {
    dgprf_pdl_sync();
    __shared__ float red[32];
    const int chain = blockIdx.y;
    const float* Y = a.Y + chain * a.y_cs;
    const int64_t row_begin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, row_step = (int64_t)gridDim.x * blockDim.x;
    float ll_acc = 0.f, g_acc = 0.f;

    if (a.likelihood == DGPRF_LIK_GAUSSIAN) {
        const float llv = __ldg(a.lik_log_var + chain * a.h_cs);
        const float inv_var = expf(-llv);
        for (int64_t row = row_begin; row < a.B; row += row_step) {
            float ll = 0.f, se = 0.f;
            for (int j = 0; j < a.D; ++j) {
                const float f = slab_load_rt(a.F, chain, row, j);
                const float r = __ldg(Y + row * a.D + j) - f;
                const float q = r * r * inv_var;
                ll += -0.5f * (DGPRF_LOG_2PI + llv + q);
                se += r * r;
                g_acc += 0.5f * (1.f - q);
                if (a.dF) a.dF[chain * a.df_cs + row * a.D + j] = -(r * inv_var) * a.inv_B;
            }
            if (a.ll_rows) a.ll_rows[(int64_t)chain * a.B + row] = ll;
            if (a.aux_rows) a.aux_rows[(int64_t)chain * a.B + row] = se / (float)a.D;
            ll_acc += ll;
        }
    } else {
        for (int64_t row = row_begin; row < a.B; row += row_step) {
            float f[kMaxClasses];
            float mx = -INFINITY;
            int arg = 0;
#pragma unroll 1
            for (int j = 0; j < a.D; ++j) {
                f[j] = slab_load_rt(a.F, chain, row, j);
                if (f[j] > mx) { mx = f[j]; arg = j; }     // first maximum, like tf.argmax
            }
            float se = 0.f;
            for (int j = 0; j < a.D; ++j) se += expf(f[j] - mx);
            const float lse = mx + logf(se);
            const int label = a.Y ? (int)__ldg(Y + row) : -1;         // tf.cast(Y, int32)[:, 0]; Y is [B,1]
            const float ll = ((label >= 0 && label < a.D) ? f[label] : NAN) - lse;
            if (a.dF || a.probs) {
                for (int j = 0; j < a.D; ++j) {
                    const float p = expf(f[j] - lse);
                    if (a.probs) a.probs[((int64_t)chain * a.B + row) * a.D + j] = p;
                    if (a.dF) a.dF[chain * a.df_cs + row * a.D + j] = (p - (j == label ? 1.f : 0.f)) * a.inv_B;
                }
            }
            if (a.ll_rows) a.ll_rows[(int64_t)chain * a.B + row] = ll;
            if (a.aux_rows) a.aux_rows[(int64_t)chain * a.B + row] = (arg == label) ? 1.f : 0.f;
            ll_acc += ll;
        }
    }
    const float tot = block_sum(ll_acc, red);
    const bool multi = gridDim.x > 1;
    if (threadIdx.x == 0) {
        if (multi) a.part[((int64_t)chain * gridDim.x + blockIdx.x) * 2] = tot;
        else if (a.ll_sum) a.ll_sum[chain] = tot;
    }
    if (a.g_lik_log_var != nullptr) {      // uniform branch
        const float g = block_sum(g_acc, red);
        if (threadIdx.x == 0) {
            if (multi) a.part[((int64_t)chain * gridDim.x + blockIdx.x) * 2 + 1] = g;
            else a.g_lik_log_var[chain * a.g_cs] = g * a.inv_B;
        }
    }
}

__global__ void __launch_bounds__(64) k3_loglik_final(const LikArgs a, int nblk) {
    dgprf_pdl_sync();
    const int chain = blockIdx.x;
    if (threadIdx.x == 0) {
        float t = 0.f, g = 0.f;
        for (int i = 0; i < nblk; ++i) {
            t += a.part[((int64_t)chain * nblk + i) * 2];
            g += a.part[((int64_t)chain * nblk + i) * 2 + 1];
        }
        if (a.ll_sum) a.ll_sum[chain] = t;
        if (a.g_lik_log_var) a.g_lik_log_var[chain * a.g_cs] = g * a.inv_B;
    }
}

int dgprf_launch_loglik(const LikArgs& a, int n_chains, cudaStream_t st) {
    DGPRF_REQUIRE(a.likelihood == DGPRF_LIK_GAUSSIAN || a.D <= kMaxClasses,
                  "softmax with %d classes > %d unsupported", a.D, kMaxClasses);
    int nblk = 1, threads = kLikThreads;
    if (a.part != nullptr && a.B > 256) {                  // one row per thread, up to 64 CTAs per chain
        threads = 256;
        nblk = ceil_div(a.B, threads);
        if (nblk > 64) nblk = 64;
    }
    { ProfScope _ps("k3_loglik", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k3_loglik, dim3(nblk, n_chains), dim3(threads), 0, st, a)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    if (nblk > 1 && (a.ll_sum || a.g_lik_log_var)) {
        DGPRF_CHECK_CUDA(dgprf_launch_pdl(k3_loglik_final, dim3(n_chains), dim3(64), 0, st, a, nblk));
        DGPRF_CHECK_CUDA(cudaGetLastError());
    }
    return DGPRF_OK;
}
