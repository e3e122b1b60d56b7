// K11: two-shot all-reduce (sum) of the flat gradient of the data-parallel step over NVLink PEER MEMORY -- no library
// collective on the step.  The data term of dU/dtheta is a mean over the minibatch (models/dgp.py:174), so a minibatch
// split by rows over GPUs needs the sum of the per-GPU gradients before the update of models/dgp.py:206-216.
//
// Every rank owns one symmetric allocation [2][n_pad] floats that all peers have mapped (torch.distributed._symmetric_memory
// supplies the mapping and a zero-initialised signal pad; this file only sees raw pointers):
//   half 0: the rank's own gradient (written by the backward / slab-sum kernels of the step),
//   half 1: the reduced gradient (written by the owners of its slices).
// Rank r owns the contiguous slice [r, r + 1) * n_pad / world.  Three small launches, chained with programmatic stream
// serialization like every other kernel of the layered step:
//   k11_signal_wait (phase 0)  tell every peer "my gradient of step `epoch` is complete", wait until all of theirs are;
//   k11_reduce_push            read the own slice from every rank's half 0 IN RANK ORDER (so the sum is the same whoever
//                              forms it), add, and store the result into half 1 of EVERY rank (128-bit peer loads / stores);
//   k11_signal_wait (phase 1)  tell every peer "my pushes are done", wait for all of theirs: half 1 is complete.
// Signals are monotonic epochs (never reset): word [phase * world + src] of a rank's pad holds the last epoch `src` has
// signalled.  Single-buffered and safe: a rank overwrites half 0 only after phase 1 of the previous step, i.e. after every
// peer has finished reading it; a peer pushes into half 1 only after phase 0 of the next step, which this rank signals
// after its update kernel has consumed half 1 (stream order).
// A wait gives up after ~2 s of polling (a peer that died must not hang the GPU): it raises the sticky error word the
// host reads back (and clears) with dgprf_peer_allreduce_status.
// ~4 MB at 8 GPUs: 0.5 MB slices, 7 peer reads + 7 peer writes of 0.5 MB per rank: the cost is three launch latencies and
// two signal round trips, not bandwidth (NCCL's ring / NVLS all-reduce of the same buffer: 35-55 us).
#include "kernels.cuh"

constexpr int kMaxPeers = 16;
struct PeerArgs {
    float* buf[kMaxPeers];                 // symmetric buffers, indexed by rank
    unsigned int* sig[kMaxPeers];          // signal pads, indexed by rank
    int32_t rank, world;
    int64_t n_pad;                         // floats per half, multiple of 4 * world
    uint32_t epoch;
    uint32_t sig_off;                      // first signal word used (uint32 index)
};

__device__ unsigned int g_k11_error = 0;   // sticky: (phase + 1) << 8 | peer that never signalled

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer_v4(float* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp: lane r talks to rank r
__global__ void __launch_bounds__(32) k11_signal_wait(const __grid_constant__ PeerArgs a, const int phase) {
    dgprf_pdl_sync();
    const int r = threadIdx.x;
    if (r >= a.world) return;
    __threadfence_system();                // everything this GPU wrote before (previous kernels) is ordered before the signal
    const uint32_t word = a.sig_off + (uint32_t)(phase * a.world);
    st_release_sys(a.sig[r] + word + a.rank, a.epoch);
    const unsigned int* mine = a.sig[a.rank] + word + r;
    const long long t0 = clock64();
    // epochs are monotonic; the signed difference keeps the comparison right across a wrap
    while ((int)(ld_acquire_sys(mine) - a.epoch) < 0) {
        if (clock64() - t0 > 4000000000LL) {           // ~2 s
            atomicExch(&g_k11_error, ((unsigned int)(phase + 1) << 8) | (unsigned int)r);
            break;
        }
        __nanosleep(64);
    }
}

__global__ void __launch_bounds__(256) k11_reduce_push(const __grid_constant__ PeerArgs a) {
    dgprf_pdl_sync();
    const int64_t slice4 = a.n_pad / a.world / 4;      // float4 per slice
    const int64_t base = (int64_t)a.rank * slice4 * 4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < slice4; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = base + 4 * e;
        float4 v[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) v[r] = ld_peer_v4(a.buf[r] + off);       // all loads in flight, then the adds in rank order
        float4 acc = v[0];
#pragma unroll
        for (int r = 1; r < kMaxPeers; ++r)
            if (r < a.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) st_peer_v4(a.buf[r] + a.n_pad + off, acc);
    }
    // no fence here: the phase-1 signal is raised by the NEXT kernel of the stream, after a system-scope fence of its own, and a
    // kernel boundary orders this grid's stores before it (a fence.sys per thread made the 2-GPU case 30 -> 49 us)
}

extern "C" int dgprf_peer_allreduce(void* const* bufs, void* const* sigs, int rank, int world, int64_t n_pad,
                                    unsigned int epoch, unsigned int sig_word_offset, void* stream) {
    DGPRF_REQUIRE(bufs != nullptr && sigs != nullptr, "bufs / sigs is NULL");
    DGPRF_REQUIRE(world >= 1 && world <= kMaxPeers, "world=%d out of range [1, %d]", world, kMaxPeers);
    DGPRF_REQUIRE(rank >= 0 && rank < world, "rank=%d out of range [0, %d)", rank, world);
    DGPRF_REQUIRE(n_pad > 0 && n_pad % (4 * (int64_t)world) == 0, "n_pad=%lld is not a multiple of 4 * world", (long long)n_pad);
    DGPRF_REQUIRE(epoch != 0, "epoch 0 is the initial state of the signal pads");
    PeerArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < world; ++r) {
        DGPRF_REQUIRE(bufs[r] != nullptr && sigs[r] != nullptr, "rank %d: NULL buffer / signal pad", r);
        DGPRF_REQUIRE((reinterpret_cast<uintptr_t>(bufs[r]) & 15) == 0, "rank %d: buffer not 16-byte aligned", r);
        a.buf[r] = static_cast<float*>(bufs[r]);
        a.sig[r] = static_cast<unsigned int*>(sigs[r]);
    }
    a.rank = rank; a.world = world; a.n_pad = n_pad; a.epoch = epoch; a.sig_off = sig_word_offset;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t slice4 = n_pad / world / 4;
    int blocks = ceil_div(slice4, 256);
    if (blocks > 4 * 148) blocks = 4 * 148;          // a peer load is a ~2-3 us round trip: a few vectors per thread, all peers in flight
    { ProfScope _ps("k11_signal_wait", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k11_signal_wait, dim3(1), dim3(32), 0, st, a, 0)); }
    { ProfScope _ps("k11_reduce_push", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k11_reduce_push, dim3(blocks), dim3(256), 0, st, a)); }
    { ProfScope _ps("k11_signal_wait", st); DGPRF_CHECK_CUDA(dgprf_launch_pdl(k11_signal_wait, dim3(1), dim3(32), 0, st, a, 1)); }
    DGPRF_CHECK_CUDA(cudaGetLastError());
    return DGPRF_OK;
}

extern "C" int dgprf_peer_allreduce_status(unsigned int* status) {
    DGPRF_REQUIRE(status != nullptr, "status is NULL");
    DGPRF_CHECK_CUDA(cudaMemcpyFromSymbol(status, g_k11_error, sizeof(unsigned int)));      // (synchronises the device)
    if (*status != 0) {                                // read and clear: the word reports the waits since the last query
        const unsigned int zero = 0;
        DGPRF_CHECK_CUDA(cudaMemcpyToSymbol(g_k11_error, &zero, sizeof(unsigned int)));
    }
    return DGPRF_OK;
}
