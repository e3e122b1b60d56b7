// Why does the saved-feature store of the pipelined forward run at 3.5 TB/s when bulk stores reach 6.4?  This replays ITS
// store pattern with no compute: every CTA owns a 128-row block of a [B, F] fp32 matrix and walks column tiles; per tile it
// TMA-stores four [128 x 32] SWIZZLE_128B boxes (cos b0, b1 | sin b0, b1 at +M columns) from one shared tile, waits until
// the tile has been read (single-buffered, as in the kernel) or keeps two tiles in flight (DEPTH 2).  Variants: row pitch
// F vs F + pad (a 32 KB power-of-two pitch may camp on few HBM channels), tile order.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/tma_store_pattern scripts/tma_store_pattern.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t s, int c, int r) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(s), "r"(c), "r"(r) : "memory");
}
// BLOCKED layout: the matrix is kept as [row block][column tile][cos b0 | cos b1 | sin b0 | sin b1][128 rows][32 floats], so the
// four boxes of a tile are one contiguous 64 KB run in HBM (the tensor map is a plain [n_blocks * 128, 32] matrix)
template <int DEPTH>
__global__ void kb(const __grid_constant__ CUtensorMap map, int B, int M, int CS, int order) {
    extern __shared__ __align__(1024) char sm[];
    for (int i = threadIdx.x; i < DEPTH * 65536 / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = make_float4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int n_rb = B / 128, n_ct = M / 64, per = n_ct / CS;
    for (int w = blockIdx.x; w < n_rb * CS; w += gridDim.x) {
        const int rb = order == 0 ? w / CS : w % n_rb, cs = order == 0 ? w % CS : w / n_rb;
        for (int t = 0; t < per; ++t) {
            const int ct = cs * per + t;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm) + (t % DEPTH) * 65536;
            for (int b = 0; b < 4; ++b) tma_store_2d(&map, s + b * 16384, 0, ((rb * n_ct + ct) * 4 + b) * 128);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int DEPTH>
__global__ void k(const __grid_constant__ CUtensorMap map, int B, int M, int CS, int order) {
    extern __shared__ __align__(1024) char sm[];
    for (int i = threadIdx.x; i < DEPTH * 65536 / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = make_float4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int n_rb = B / 128, n_ct = M / 64, per = n_ct / CS;
    // CTA id -> (row block, column split): persistent loop over work items
    for (int w = blockIdx.x; w < n_rb * CS; w += gridDim.x) {
        const int rb = order == 0 ? w / CS : w % n_rb, cs = order == 0 ? w % CS : w / n_rb;
        for (int t = 0; t < per; ++t) {
            const int c0 = (cs * per + t) * 64;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm) + (t % DEPTH) * 65536;
            for (int b = 0; b < 2; ++b) {
                tma_store_2d(&map, s + b * 16384, c0 + 32 * b, rb * 128);
                tma_store_2d(&map, s + (2 + b) * 16384, M + c0 + 32 * b, rb * 128);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
    const int B = 65536, M = 4096, F = 2 * M;
    {
        float* p; cudaMalloc(&p, (size_t)B * F * 4);
        CUtensorMap map;
        cuuint64_t dims[2] = {32, (cuuint64_t)B * F / 32}, strides[1] = {128};
        cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
        cudaFuncSetAttribute(kb<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
        cudaFuncSetAttribute(kb<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096);
        for (int CS : {2, 8})
            for (int order : {0, 1})
                for (int depth : {1, 2}) {
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    auto run = [&] { if (depth == 1) kb<1><<<148, 128, 66560>>>(map, B, M, CS, order); else kb<2><<<148, 128, 132096>>>(map, B, M, CS, order); };
                    run(); cudaDeviceSynchronize();
                    cudaEventRecord(a); for (int i = 0; i < 3; ++i) run(); cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
                    printf("BLOCKED (64 KB contiguous per tile) CS %2d order %d depth %d: %.3f ms  %.0f GB/s  (%s)\n", CS, order, depth, ms, (double)B * F * 4 / ms / 1e6,
                           cudaGetErrorString(cudaGetLastError()));
                }
        cudaFree(p);
    }
    for (int pad : {0, 32, 64, 256}) {
        const size_t pitch = F + pad;
        float* p; cudaMalloc(&p, (size_t)B * pitch * 4);
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)B}, strides[1] = {pitch * 4};
        cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
        cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
        cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132096);
        for (int CS : {2, 8, 64})
            for (int order : {0, 1})
                for (int depth : {1, 2}) {
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    auto run = [&] { if (depth == 1) k<1><<<148, 128, 66560>>>(map, B, M, CS, order); else k<2><<<148, 128, 132096>>>(map, B, M, CS, order); };
                    run(); cudaDeviceSynchronize();
                    cudaEventRecord(a); for (int i = 0; i < 3; ++i) run(); cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
                    printf("pitch F+%-3d CS %2d order %d depth %d: %.3f ms  %.0f GB/s  (%s)\n", pad, CS, order, depth, ms, (double)B * F * 4 / ms / 1e6,
                           cudaGetErrorString(cudaGetLastError()));
                }
        cudaFree(p);
    }
}
