// What is the WRITE-only ceiling of HBM3e on this B200?  Hand-written store kernels (no library fill): plain 128-bit
// stores, streaming (.cs) stores, no-allocate (.L1::no_allocate / evict_first) stores, and TMA bulk stores
// shared -> global (cp.async.bulk.global.shared::cta), each over a grid sweep, on a 2 GiB buffer (>> 126 MB L2).
// The saved-feature store of the pipelined forward (k1_fwd_tc2) is a pure write stream: this is its roofline.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/hbm_store_ceiling scripts/hbm_store_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_st(float4* __restrict__ p, size_t n4, int mode) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (mode == 0) p[i] = v;
        else if (mode == 1) __stcs(p + i, v);
        else if (mode == 2) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else if ((i & 1) == 0 && i + 1 < n4) {          // 256-bit stores (sm_100+), L2 evict-first
            const unsigned u = 0x3f800000u;
            asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + i), "r"(u) : "memory");
        }
    }
}
// every CTA owns contiguous 32 KB chunks: fill a shared tile once, then stream it out with bulk stores, D deep
template <int DEPTH>
__global__ void k_tma(char* __restrict__ p, size_t bytes, int chunk) {
    extern __shared__ __align__(128) char tile[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<float4*>(tile)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t nchunks = bytes / chunk;
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(tile);
        int inflight = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + c * (size_t)chunk), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= DEPTH) { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory"); --inflight; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
template <typename F> static double time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize(); cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 10;
}
int main() {
    const size_t bytes = (size_t)2 << 30;
    char* p; cudaMalloc(&p, bytes);
    char* q; cudaMalloc(&q, bytes);
    printf("kernel, config, ms, GB/s written\n");
    double best = 0;
    { double ms = time_ms([&] { cudaMemsetAsync(p, 1, bytes); }); printf("cudaMemsetAsync, -, %.3f, %.1f\n", ms, bytes / ms / 1e6); }
    { double ms = time_ms([&] { cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D (read+write), -, %.3f, %.1f (x2 = %.1f total)\n", ms, bytes / ms / 1e6, 2 * bytes / ms / 1e6); }
    const char* names[4] = {"st.global.v4", "st.global.cs.v4", "st.global.L1::no_allocate.v4", "st.global.L2::evict_first.v8 (256-bit)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int cps : {1, 2, 4, 8, 16})
            for (int th : {256, 1024}) {
                double ms = time_ms([&] { k_st<<<148 * cps, th>>>((float4*)p, bytes / 16, mode); });
                double g = bytes / ms / 1e6; if (g > best) best = g;
                printf("%s, %d CTAs/SM x %d thr, %.3f, %.1f\n", names[mode], cps, th, ms, g);
            }
    cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int chunk : {8192, 32768, 65536})
        for (int cps : {1, 2, 3}) {
            double ms = time_ms([&] { k_tma<2><<<148 * cps, 128, chunk>>>(p, bytes, chunk); });
            double g = bytes / ms / 1e6; if (g > best) best = g;
            printf("TMA bulk store depth 2, %d B chunks %d CTAs/SM, %.3f, %.1f\n", chunk, cps, ms, g);
            ms = time_ms([&] { k_tma<8><<<148 * cps, 128, chunk>>>(p, bytes, chunk); });
            g = bytes / ms / 1e6; if (g > best) best = g;
            printf("TMA bulk store depth 8, %d B chunks %d CTAs/SM, %.3f, %.1f\n", chunk, cps, ms, g);
        }
    printf("best hand-written write-only rate: %.1f GB/s   (%s)\n", best, cudaGetErrorString(cudaDeviceSynchronize()));
}
