// Per-SM throughput of TMA loads global -> shared for a 64 KB tile, L2-resident source, 2-stage ring, one issuing thread per CTA:
//   tensor-tiled boxes [32 floats x 128 rows] (SWIZZLE_128B, the kernels' pattern; four boxes per tile, blocked layout = each box
//   is 16 KB contiguous in global) against 1-D bulk copies (cp.async.bulk, no tensor map) of the same bytes.
// Question behind it: the pipelined backward waits ~4 k cycles for every 64 KB saved-feature tile (18 B / clk / SM) and neither
// an L2 prefetch nor the contiguous layout changed that -- is that the tensor-tiled TMA path itself?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/tma_load_rate scripts/tma_load_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* m, uint32_t dst, uint64_t* b, int c, int r) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(s32(b)), "r"(c), "r"(r) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

// mode 0: 4 tensor boxes per tile; 1: one 64 KB bulk copy; 2: four 16 KB bulk copies.  STAGES-deep ring.
template <int STAGES>
__global__ void k(const __grid_constant__ CUtensorMap map, const char* src, int tiles_total, int iters, int mode, long long* cyc) {
    extern __shared__ __align__(1024) char sm[];
    __shared__ uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) mbar_init(bar + i);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    int issued = 0, waited = 0;
    auto issue = [&](int i) {
        const int st = i % STAGES;
        const int tile = (blockIdx.x + (long long)i * gridDim.x) % tiles_total;
        mbar_expect(bar + st, 65536);
        const uint32_t dst = s32(sm) + st * 65536;
        if (mode == 0) for (int b = 0; b < 4; ++b) tma_2d(&map, dst + b * 16384, bar + st, 0, (tile * 4 + b) * 128);
        else if (mode == 1) bulk_1d(dst, src + (size_t)tile * 65536, 65536, bar + st);
        else for (int b = 0; b < 4; ++b) bulk_1d(dst + b * 16384, src + (size_t)tile * 65536 + b * 16384, 16384, bar + st);
    };
    for (; issued < STAGES && issued < iters; ++issued) issue(issued);
    for (; waited < iters; ++waited) {
        mbar_wait(bar + waited % STAGES, (waited / STAGES) & 1);
        if (issued < iters) { issue(issued); ++issued; }     // the stage just consumed is refilled at once
    }
    if (cyc) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t bytes = 48u << 20;                            // 48 MB source: L2 resident after the first pass
    char* p; cudaMalloc(&p, bytes); cudaMemset(p, 1, bytes);
    long long* cyc; cudaMalloc(&cyc, 148 * sizeof(long long));
    const int tiles_total = (int)(bytes / 65536);
    for (int swz = 0; swz < 2; ++swz) {
        CUtensorMap map;
        cuuint64_t dims[2] = {32, (cuuint64_t)(bytes / 128)}, strides[1] = {128};
        cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            swz ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
        cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024);
        cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 1024);
        for (int grid : {8, 148})
            for (int mode = 0; mode < 3; ++mode) {
                if (swz == 1 && mode != 0) continue;
                for (int stages : {2, 3}) {
                    const int iters = 400;
                    auto run = [&] { if (stages == 2) k<2><<<grid, 32, 2 * 65536 + 1024>>>(map, p, tiles_total, iters, mode, cyc);
                                     else k<3><<<grid, 32, 3 * 65536 + 1024>>>(map, p, tiles_total, iters, mode, cyc); };
                    run(); cudaDeviceSynchronize();
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    cudaEventRecord(a); run(); cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    long long h[148]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
                    const char* names[3] = {"tensor boxes 4 x [32 x 128]", "bulk 1 x 64 KB", "bulk 4 x 16 KB"};
                    printf("%-28s %s  CTAs %3d  stages %d: %7.0f cycles / 64 KB tile / SM = %5.1f B/clk/SM, aggregate %6.0f GB/s  (%s)\n", names[mode],
                           mode == 0 ? (swz ? "ATOM_32B" : "SW128   ") : "        ", grid, stages, avg / iters, 65536.0 * iters / avg,
                           (double)grid * iters * 65536 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
                }
            }
    }
}
