// Operand prep of the pipelined tensor-core kernels for ALL layers of a step in one launch.  Everything here depends only
// on the parameters (z, log_inv_ls, mean, W), not on activations, so it can run before the first layer:
//   zt [2][M][128]   z^T split into tf32 hi / lo, K-major, zero padded            (forward, input width <= 128)
//   ot [2][M][Kp]    Omega^T = (exp(log_inv_ls) z + mean)^T, tf32 hi / lo          (forward, WIDE variant)
//   wt [NG][F]       W^T rounded to tf32                                           (forward GEMM #2)
//   wp [F][32]       W rows zero-padded to 32 columns, tf32                        (backward MMA-1; TRAIN / HYPER modes)
// grid (tasks, layers, chains): a block decodes its task from blockIdx.x against the layer's task counts.
#include "kernels.cuh"
#include "tc_common.cuh"

__global__ void __launch_bounds__(256) k_prep_layers(const __grid_constant__ PrepArgs a) {
    dgprf_pdl_sync();
    __shared__ float tile[32][33];
    const PrepLayer& y = a.L[blockIdx.y];
    const int chain = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    int b = blockIdx.x;
    if (b < y.n_zt) {                                                 // ---- zt tile: 32 features x 32 K ----
        if (y.z_cs == 0 && chain > 0) return;                         // shared spectral draws: one copy
        const int m0 = (b >> 2) * 32, k0 = (b & 3) * 32;
        const float* zz = y.z + chain * y.z_cs;
        for (int i = ty; i < 32; i += 8)
            tile[i][tx] = (k0 + i < y.d && m0 + tx < y.M) ? __ldg(zz + (int64_t)(k0 + i) * y.M + m0 + tx) : 0.f;
        __syncthreads();
        float* hi = y.zt + (int64_t)chain * 2 * y.M * 128;
        float* lo = hi + (int64_t)y.M * 128;
        for (int i = ty; i < 32; i += 8)
            if (m0 + i < y.M) {
                const float v = tile[tx][i];
                const float h = tc::to_tf32(v);
                hi[(int64_t)(m0 + i) * 128 + k0 + tx] = h;
                lo[(int64_t)(m0 + i) * 128 + k0 + tx] = tc::to_tf32(v - h);
            }
        return;
    }
    b -= y.n_zt;
    if (b < y.n_ot) {                                                 // ---- Omega^T tile (layers/rf_layers.py: Omega) ----
        const int nk = y.Kp / 32;
        const int m0 = (b / nk) * 32, k0 = (b % nk) * 32;
        const float* zz = y.z + chain * y.z_cs;
        const float* ls = y.log_inv_ls + chain * y.h_cs;
        const float* mean = y.has_mean ? y.mean + chain * y.h_cs : nullptr;
        for (int i = ty; i < 32; i += 8) {
            const int q = k0 + i;
            float v = 0.f;
            if (q < y.d && m0 + tx < y.M) {
                v = expf(__ldg(ls + q)) * __ldg(zz + (int64_t)q * y.M + m0 + tx);
                if (mean != nullptr) v += __ldg(mean + q);
            }
            tile[i][tx] = v;
        }
        __syncthreads();
        float* hi = y.ot + (int64_t)chain * 2 * y.M * y.Kp;
        float* lo = hi + (int64_t)y.M * y.Kp;
        for (int i = ty; i < 32; i += 8)
            if (m0 + i < y.M) {
                const float v = tile[tx][i];
                const float h = tc::to_tf32(v);
                hi[(int64_t)(m0 + i) * y.Kp + k0 + tx] = h;
                lo[(int64_t)(m0 + i) * y.Kp + k0 + tx] = tc::to_tf32(v - h);
            }
        return;
    }
    b -= y.n_ot;
    if (b < y.n_wt) {                                                 // ---- W^T tile ----
        const int nj = (y.NG + 31) / 32;
        const int f0 = (b / nj) * 32, j0 = (b % nj) * 32;
        const float* WW = y.W + chain * y.w_cs;
        for (int i = ty; i < 32; i += 8)
            tile[i][tx] = (f0 + i < y.F && j0 + tx < y.g) ? __ldg(WW + (int64_t)(f0 + i) * y.g + j0 + tx) : 0.f;
        __syncthreads();
        float* o = y.wt + (int64_t)chain * y.NG * y.F;
        for (int i = ty; i < 32; i += 8)
            if (j0 + i < y.NG && f0 + tx < y.F) o[(int64_t)(j0 + i) * y.F + f0 + tx] = tc::to_tf32(tile[tx][i]);
        return;
    }
    b -= y.n_wt;
    if (b < y.n_wp) {                                                 // ---- padded W rows ----
        // four padded columns per thread, one 128-bit store (the buffer is 256-byte aligned, rows are 32 floats)
        const int64_t i = ((int64_t)b * 256 + threadIdx.x) * 4;
        if (i >= (int64_t)y.F * 32) return;
        const int64_t f = i >> 5;
        const int j = (int)(i & 31);
        const float* w = y.W + chain * y.w_cs + f * y.g;
        float4 o;
        o.x = j + 0 < y.g ? tc::to_tf32(__ldg(w + j + 0)) : 0.f;
        o.y = j + 1 < y.g ? tc::to_tf32(__ldg(w + j + 1)) : 0.f;
        o.z = j + 2 < y.g ? tc::to_tf32(__ldg(w + j + 2)) : 0.f;
        o.w = j + 3 < y.g ? tc::to_tf32(__ldg(w + j + 3)) : 0.f;
        *reinterpret_cast<float4*>(y.wp + (int64_t)chain * y.F * 32 + i) = o;
    }
}

int dgprf_launch_prep_layers(const PrepArgs& a, int n_chains, cudaStream_t st) {
    int max_tasks = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        const int t = a.L[l].n_zt + a.L[l].n_ot + a.L[l].n_wt + a.L[l].n_wp;
        if (t > max_tasks) max_tasks = t;
    }
    if (max_tasks == 0) return DGPRF_OK;
    { ProfScope _ps("k_prep_layers", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k_prep_layers, dim3(max_tasks, a.n_layers, n_chains), dim3(256), 0, st, a)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
