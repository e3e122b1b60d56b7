// How fast does an SM run STRAIGHT-LINE code it has never executed?  K9 / K10 execute ~100 KB of mostly non-repeating
// SASS per launch (one pass over unrolled GEMM chains per layer); ncu's top stall reason for them is "no instruction".
// This kernel executes N independent FFMAs laid out as straight-line code (no loop), once, with 16 warps per SM, and
// reports cycles per instruction for the first (cold) and later (warm, same launch) passes over the same code.
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct Unroll {
    static __device__ __forceinline__ void run(float (&a)[8], float b) {
        Unroll<N / 2>::run(a, b);
        Unroll<N - N / 2>::run(a, b);
    }
};
template <> struct Unroll<1> {
    static __device__ __forceinline__ void run(float (&a)[8], float b) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
    }
};
template <int N>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int passes) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    const float b = 1.0001f;
    long long t[4];
    __syncthreads();
    t[0] = clock64();
    for (int p = 0; p < passes; ++p) {
        Unroll<N>::run(a, b);
        __syncthreads();
        if (p < 3) t[p + 1] = clock64();
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) for (int p = 0; p < 3; ++p) cyc[p] = t[p + 1] - t[p];
}
template <int N> void run(float* out, long long* cyc) {
    long long h[3];
    for (int rep = 0; rep < 2; ++rep) {
        k<N><<<148, 512>>>(out, cyc, 3);
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    }
    const double ninst = 8.0 * N;          // per warp
    printf("%6d FFMA (%4d KB of SASS): pass 1 %7lld cyc (%.2f cyc/inst/warp), pass 2 %7lld (%.2f), pass 3 %7lld (%.2f)   [16 warps/SM: issue floor 4.00]\n",
           8 * N, 8 * N * 16 / 1024, h[0], h[0] / ninst, h[1], h[1] / ninst, h[2], h[2] / ninst);
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 64);
    run<64>(out, cyc); run<128>(out, cyc); run<256>(out, cyc); run<512>(out, cyc); run<1024>(out, cyc); run<2048>(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
