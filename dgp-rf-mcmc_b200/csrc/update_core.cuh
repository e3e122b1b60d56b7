// The per-vector SGHMC / SGLD update shared by K5 (stand-alone update kernel) and K9 (row-fused step with
// the update behind a grid barrier).  models/dgp.py:206-216.
#pragma once
#include "kernels.cuh"
#include "philox.cuh"

struct SegTable {
    int64_t offset[DGPRF_MAX_SEGMENTS];
    int64_t end[DGPRF_MAX_SEGMENTS];
    float sqrt_mass[DGPRF_MAX_SEGMENTS];
    float inv_mass[DGPRF_MAX_SEGMENTS];
    int32_t flags[DGPRF_MAX_SEGMENTS];
};

// Sum over the gradient slabs p = sub, sub+LPV, ... of vector i (LPV lanes cooperate; the caller combines
// the lanes with shuffle_sum_lpv).  All loads of a batch are in flight before the fixed-order adds.
template <int LPV>
__device__ __forceinline__ float4 slab_sum_lane(const float* __restrict__ grad, int64_t part_stride, int n_part, int sub, int64_t i) {
    if (LPV == 1 && n_part == 1) return __ldcg(reinterpret_cast<const float4*>(grad + i));     // dense gradient
    float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NB = LPV == 8 ? 16 : 8;                 // loads in flight per lane (8 lanes x 16 covers 128 slabs in one batch)
    for (int p0 = sub; p0 < n_part; p0 += NB * LPV) {
        float4 gp[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u)
            gp[u] = (p0 + u * LPV) < n_part ? __ldcg(reinterpret_cast<const float4*>(grad + (int64_t)(p0 + u * LPV) * part_stride + i))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < NB; ++u) { gsum.x += gp[u].x; gsum.y += gp[u].y; gsum.z += gp[u].z; gsum.w += gp[u].w; }
    }
    return gsum;
}
template <int LPV>
__device__ __forceinline__ float4 shuffle_sum_lpv(float4 g) {
#pragma unroll
    for (int o = LPV / 2; o > 0; o >>= 1) {
        g.x += __shfl_xor_sync(0xffffffffu, g.x, o); g.y += __shfl_xor_sync(0xffffffffu, g.y, o);
        g.z += __shfl_xor_sync(0xffffffffu, g.z, o); g.w += __shfl_xor_sync(0xffffffffu, g.w, o);
    }
    return g;
}

// The random inputs of the update of vector i4 (independent of the gradient, so a fused kernel can draw them while it
// waits at its grid barrier): e = the injected / Philox noise (zero at T = 0), m = the resampled momentum if asked.
__device__ __forceinline__ void sgmcmc_draw_vec(const UpdArgs& a, int chain, int64_t i4, float4& e, float4& m) {
    const int64_t i = i4 << 2;
    const uint64_t step = a.step + (a.step_dev ? (uint64_t)__ldg(a.step_dev) : 0ull);
    if (a.resample) {
        if (a.mom_inject) m = __ldg(reinterpret_cast<const float4*>(a.mom_inject + chain * a.cs + i));
        else m = philox_normal4(a.seed, (uint64_t)chain, (uint64_t)i4, step, a.stream_base + 1u);
    }
    e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.noise_scale != 0.f) {
        if (a.eps_inject) e = __ldg(reinterpret_cast<const float4*>(a.eps_inject + chain * a.cs + i));
        else e = philox_normal4(a.seed, (uint64_t)chain, (uint64_t)i4, step, a.stream_base);
    }
}

// Update the four parameters at flat offset i = 4*i4 of `chain` given their summed data gradient g, the current
// values th / m (loaded by the caller together with the gradient so all three are in flight at once; m already
// resampled if asked) and the noise e of sgmcmc_draw_vec.
__device__ __forceinline__ void sgmcmc_apply_vec(const UpdArgs& a, const SegTable& tab, int chain, int64_t i4, float4 g,
                                                 float4 th, float4 m, float4 e) {
    const int64_t i = i4 << 2;
    float* theta = a.theta + chain * a.cs;
    float* mom = a.mom + chain * a.cs;
    int lo = 0, hi = a.n_seg - 1;                         // segment lookup (table sorted by offset)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab.offset[mid] <= i) lo = mid; else hi = mid - 1;
    }
    if (i < tab.offset[lo] || i >= tab.end[lo]) return;   // alignment padding
    const int64_t live = tab.end[lo] - i;                 // 1..4 live lanes, the rest is padding
    const float sqrt_mass = tab.sqrt_mass[lo], inv_mass = tab.inv_mass[lo];
    const bool prior = tab.flags[lo] & 1;

    if (prior) {
        g.x = fmaf(th.x, a.inv_N, g.x); g.y = fmaf(th.y, a.inv_N, g.y);
        g.z = fmaf(th.z, a.inv_N, g.z); g.w = fmaf(th.w, a.inv_N, g.w);
    }
    const float ns = a.noise_scale * sqrt_mass, hm = a.h * inv_mass;
    m.x = fmaf(ns, e.x, fmaf(a.beta, m.x, -a.hN * g.x));
    m.y = fmaf(ns, e.y, fmaf(a.beta, m.y, -a.hN * g.y));
    m.z = fmaf(ns, e.z, fmaf(a.beta, m.z, -a.hN * g.z));
    m.w = fmaf(ns, e.w, fmaf(a.beta, m.w, -a.hN * g.w));
    th.x = fmaf(hm, m.x, th.x); th.y = fmaf(hm, m.y, th.y);
    th.z = fmaf(hm, m.z, th.z); th.w = fmaf(hm, m.w, th.w);
    if (live < 4) {          // keep the alignment padding at exactly zero
        if (live < 2) { th.y = 0.f; m.y = 0.f; }
        if (live < 3) { th.z = 0.f; m.z = 0.f; }
        th.w = 0.f; m.w = 0.f;
    }
    *reinterpret_cast<float4*>(theta + i) = th;
    *reinterpret_cast<float4*>(mom + i) = m;
}

__device__ __forceinline__ void sgmcmc_update_vec(const UpdArgs& a, const SegTable& tab, int chain, int64_t i4, float4 g,
                                                  float4 th, float4 m) {
    float4 e;
    sgmcmc_draw_vec(a, chain, i4, e, m);
    sgmcmc_apply_vec(a, tab, chain, i4, g, th, m, e);
}

// Host: validate the caller's segment list and build the kernel-parameter table.
int dgprf_build_segtable(const dgprf_segment* segs, int n_seg, int64_t n, SegTable* tab);
