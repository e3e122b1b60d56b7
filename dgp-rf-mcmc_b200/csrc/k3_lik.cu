// K3: likelihood reductions (warp-shuffle), emitting the backward seed dU/dF directly.
//   Gaussian  likelihoods/gaussian.py:18-25 + utils.py:46-47
//   Softmax   likelihoods/softmax.py:8-22 (labels = int(Y[:,0]))
// plus what the evaluation methods derive from the same pass
//   models/regression_model.py:46   se = mean_D (y-f)^2
//   models/classification_model.py:21-24   argmax(softmax(F)) == label
// One CTA per chain walks the rows in a fixed order -> deterministic sums.
#include "kernels.cuh"

constexpr int kLikThreads = 1024;
constexpr int kMaxClasses = 64;

__global__ void __launch_bounds__(kLikThreads) k3_loglik(const LikArgs a) {
    __shared__ float red[32];
    const int chain = blockIdx.x;
    const float* Y = a.Y + chain * a.y_cs;
    float ll_acc = 0.f, g_acc = 0.f;

    if (a.likelihood == DGPRF_LIK_GAUSSIAN) {
        const float llv = __ldg(a.lik_log_var + chain * a.h_cs);
        const float inv_var = expf(-llv);
        for (int64_t row = threadIdx.x; row < a.B; row += kLikThreads) {
            float ll = 0.f, se = 0.f;
            for (int j = 0; j < a.D; ++j) {
                const float f = slab_load(a.F, chain, row, j);
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
        for (int64_t row = threadIdx.x; row < a.B; row += kLikThreads) {
            float f[kMaxClasses];
            float mx = -INFINITY;
            int arg = 0;
#pragma unroll 1
            for (int j = 0; j < a.D; ++j) {
                f[j] = slab_load(a.F, chain, row, j);
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
    if (threadIdx.x == 0 && a.ll_sum) a.ll_sum[chain] = tot;
    if (a.g_lik_log_var != nullptr) {      // uniform branch
        const float g = block_sum(g_acc, red);
        if (threadIdx.x == 0) a.g_lik_log_var[chain * a.g_cs] = g * a.inv_B;
    }
}

int dgprf_launch_loglik(const LikArgs& a, int n_chains, cudaStream_t st) {
    DGPRF_REQUIRE(a.likelihood == DGPRF_LIK_GAUSSIAN || a.D <= kMaxClasses,
                  "softmax with %d classes > %d unsupported", a.D, kMaxClasses);
    { ProfScope _ps("k3_loglik", st); k3_loglik<<<n_chains, kLikThreads, 0, st>>>(a); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
