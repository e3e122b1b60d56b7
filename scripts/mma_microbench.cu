// Latency / throughput of the legacy mma.sync tf32 path on sm_100a (what K10's GEMM chains are made of).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int ILP>
__global__ void k(float* out, long long* cyc, int iters) {
    float c[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a = threadIdx.x, b = threadIdx.x * 3;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma(c[i], a, a + 1, a + 2, a + 3, b, b + 1);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16, 32}) {
        long long h;
#define RUN(ILP) k<ILP><<<148, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("warps/SM %2d ILP %d: %.2f cycles per MMA per warp, %.3f MMA/clk/SM\n", warps, ILP, (double)h / (iters * ILP), (double)iters * ILP * warps / h);
        RUN(1) RUN(2) RUN(4) RUN(8)
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
