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
#include "update_core.cuh"

// LPV = lanes per 128-bit vector: 1 (few gradient slabs) or 8 (many slabs, e.g. the row-fused step writes
// one slab per 8 batch rows): the 8 lanes of a group each sum every 8th slab, then combine with a fixed
// shuffle tree, so all slab loads of a vector are in flight at once and the result stays deterministic.
template <int LPV>
__global__ void __launch_bounds__(256) k5_sgmcmc_update(const UpdArgs a, const __grid_constant__ SegTable tab) {
    dgprf_pdl_sync();
    const int chain = blockIdx.y;
    const float* grad = a.grad + chain * a.grad_cs;
    const int64_t n4 = a.n >> 2;
    const int sub = LPV == 1 ? 0 : (threadIdx.x & (LPV - 1));
    const int64_t nv = ((n4 + 31) / 32) * 32;          // whole warps stay together for the shuffles
    for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPV; v < nv; v += (int64_t)gridDim.x * blockDim.x / LPV) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), th = g, m = g;
        if (sub == 0 && v < n4) {
            th = *reinterpret_cast<const float4*>(a.theta + chain * a.cs + (v << 2));
            m = *reinterpret_cast<const float4*>(a.mom + chain * a.cs + (v << 2));
        }
        if (v < n4) g = slab_sum_lane<LPV>(grad, a.part_stride, a.n_part, sub, v << 2);
        g = shuffle_sum_lpv<LPV>(g);
        if (sub == 0 && v < n4) sgmcmc_update_vec(a, tab, chain, v, g, th, m);
    }
}

int dgprf_build_segtable(const dgprf_segment* segs, int n_seg, int64_t n, SegTable* tab) {
    DGPRF_REQUIRE(n_seg >= 1 && n_seg <= DGPRF_MAX_SEGMENTS, "n_seg=%d out of range", n_seg);
    for (int s = 0; s < DGPRF_MAX_SEGMENTS; ++s) {
        const int t = s < n_seg ? s : n_seg - 1;
        DGPRF_REQUIRE((segs[t].offset & 3) == 0, "segment %d offset not a multiple of 4", t);
        DGPRF_REQUIRE(s == 0 || s >= n_seg || segs[s].offset >= segs[s - 1].offset + segs[s - 1].length,
                      "segments must be sorted and disjoint");
        DGPRF_REQUIRE(segs[t].mass > 0.f, "segment %d mass must be positive", t);
        tab->offset[s] = segs[t].offset;
        tab->end[s] = segs[t].offset + segs[t].length;
        tab->sqrt_mass[s] = sqrtf(segs[t].mass);
        tab->inv_mass[s] = 1.f / segs[t].mass;
        tab->flags[s] = segs[t].flags;
    }
    DGPRF_REQUIRE(tab->end[n_seg - 1] <= n, "segments exceed the buffer");
    return DGPRF_OK;
}

// 8 lanes per vector only pay off when there are many slabs AND too few vectors to fill the GPU with one thread each
int dgprf_update_lpv(int n_part, int64_t n4, int n_chains) {
    return (n_part > 16 && n4 * n_chains < (int64_t)148 * 2048) ? 8 : 1;
}

int dgprf_launch_update(const UpdArgs& a, const dgprf_segment* segs, int n_seg, int n_chains, cudaStream_t st) {
    DGPRF_REQUIRE((a.n & 3) == 0 && (a.cs & 3) == 0 && (a.grad_cs & 3) == 0 && (a.part_stride & 3) == 0,
                  "flat buffers must be padded to multiples of 4 floats");
    SegTable tab;
    const int rc = dgprf_build_segtable(segs, n_seg, a.n, &tab);
    if (rc) return rc;
    const int64_t n4 = a.n >> 2;
    const int lpv = dgprf_update_lpv(a.n_part, n4, n_chains);
    int blocks = ceil_div(n4 * lpv, 256);
    const int cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, n_chains);
    {
        ProfScope _ps("k5_sgmcmc_update", st);
        if (lpv == 8) DGPRF_CHECK_CUDA(dgprf_launch_pdl(k5_sgmcmc_update<8>, grid, dim3(256), 0, st, a, tab));
        else DGPRF_CHECK_CUDA(dgprf_launch_pdl(k5_sgmcmc_update<1>, grid, dim3(256), 0, st, a, tab));
    }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}
