// K5: the SGHMC / SGLD update of models/dgp.py:206-216 as ONE fused, vectorised, HBM-bound
// elementwise kernel over the flattened parameter buffer (all tensors, all chains):
//
//   h = sqrt(lr/N);  [resample: m <- N(0,1)]
//   g = sum_p grad_p (+ theta/N for segments carrying the N(0,1) prior, dgp.py:171,178-180)
//   m <- beta m - h N g + sqrt(2 (1-beta) T M_seg) eps ;  theta <- theta + (h / M_seg) m
//
// beta = 0 is SGLD.  M_seg is the per-tensor scalar mass (dgp.py:237,295), looked up in a
// small segment table.  Noise is generated in-kernel (Philox4x32-10 + Box-Muller, four
// normals per 128-bit lane) or read from injected buffers for parity tests.
// Algorithmic traffic: 20 B/parameter (read theta, m, g; write theta, m) with n_part == 1.
#include "kernels.cuh"
#include "philox.cuh"

struct SegTable {
    int64_t offset[DGPRF_MAX_SEGMENTS];
    int64_t end[DGPRF_MAX_SEGMENTS];
    float sqrt_mass[DGPRF_MAX_SEGMENTS];
    float inv_mass[DGPRF_MAX_SEGMENTS];
    int32_t flags[DGPRF_MAX_SEGMENTS];
};

__global__ void __launch_bounds__(256) k5_sgmcmc_update(const UpdArgs a, const __grid_constant__ SegTable tab) {
    const int chain = blockIdx.y;
    float* theta = a.theta + chain * a.cs;
    float* mom = a.mom + chain * a.cs;
    const float* grad = a.grad + chain * a.grad_cs;
    const int64_t n4 = a.n >> 2;
    for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i4 << 2;
        // segment lookup (table sorted by offset; segments start on multiples of 4)
        int lo = 0, hi = a.n_seg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tab.offset[mid] <= i) lo = mid; else hi = mid - 1;
        }
        if (i < tab.offset[lo] || i >= tab.end[lo]) continue;       // alignment padding
        const int64_t live = tab.end[lo] - i;                         // 1..4 live lanes, rest is padding
        const float sqrt_mass = tab.sqrt_mass[lo], inv_mass = tab.inv_mass[lo];
        const bool prior = tab.flags[lo] & 1;

        float4 th = *reinterpret_cast<const float4*>(theta + i);
        float4 m = *reinterpret_cast<const float4*>(mom + i);
        float4 g = __ldg(reinterpret_cast<const float4*>(grad + i));
        for (int p0 = 1; p0 < a.n_part; p0 += 8) {        // 8 partial slabs in flight, fixed-order adds
            float4 gp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                gp[u] = (p0 + u) < a.n_part ? __ldg(reinterpret_cast<const float4*>(grad + (p0 + u) * a.part_stride + i))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { g.x += gp[u].x; g.y += gp[u].y; g.z += gp[u].z; g.w += gp[u].w; }
        }
        if (prior) {
            g.x = fmaf(th.x, a.inv_N, g.x); g.y = fmaf(th.y, a.inv_N, g.y);
            g.z = fmaf(th.z, a.inv_N, g.z); g.w = fmaf(th.w, a.inv_N, g.w);
        }
        if (a.resample) {
            if (a.mom_inject) m = __ldg(reinterpret_cast<const float4*>(a.mom_inject + chain * a.cs + i));
            else m = philox_normal4(a.seed, (uint64_t)chain, (uint64_t)i4, a.step, a.stream_base + 1u);
        }
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.noise_scale != 0.f) {
            if (a.eps_inject) e = __ldg(reinterpret_cast<const float4*>(a.eps_inject + chain * a.cs + i));
            else e = philox_normal4(a.seed, (uint64_t)chain, (uint64_t)i4, a.step, a.stream_base);
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
}

int dgprf_launch_update(const UpdArgs& a, const dgprf_segment* segs, int n_seg, int n_chains, cudaStream_t st) {
    DGPRF_REQUIRE(n_seg >= 1 && n_seg <= DGPRF_MAX_SEGMENTS, "n_seg=%d out of range", n_seg);
    DGPRF_REQUIRE((a.n & 3) == 0 && (a.cs & 3) == 0 && (a.grad_cs & 3) == 0 && (a.part_stride & 3) == 0,
                  "flat buffers must be padded to multiples of 4 floats");
    SegTable tab;
    for (int s = 0; s < DGPRF_MAX_SEGMENTS; ++s) {
        const int t = s < n_seg ? s : n_seg - 1;
        DGPRF_REQUIRE((segs[t].offset & 3) == 0, "segment %d offset not a multiple of 4", t);
        DGPRF_REQUIRE(s == 0 || s >= n_seg || segs[s].offset >= segs[s - 1].offset + segs[s - 1].length,
                      "segments must be sorted and disjoint");
        DGPRF_REQUIRE(segs[t].mass > 0.f, "segment %d mass must be positive", t);
        tab.offset[s] = segs[t].offset;
        tab.end[s] = segs[t].offset + segs[t].length;
        tab.sqrt_mass[s] = sqrtf(segs[t].mass);
        tab.inv_mass[s] = 1.f / segs[t].mass;
        tab.flags[s] = segs[t].flags;
    }
    DGPRF_REQUIRE(tab.end[n_seg - 1] <= a.n, "segments exceed the buffer");
    const int64_t n4 = a.n >> 2;
    int blocks = ceil_div(n4, 256);
    const int cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, n_chains);
    { ProfScope _ps("k5_sgmcmc_update", st); k5_sgmcmc_update<<<grid, 256, 0, st>>>(a, tab); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
